#!/bin/bash
O=gpurun_out/r2_6; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6 | tee $O/smoke.txt
LAMEGPU_TIMING=1 BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err | cut -c1-300; cat $O/bench.json | cut -c1-1500
LAMEGPU_GROUP_NW=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_nw0.json 2> $O/bench_nw0.err; cat $O/bench_nw0.json | cut -c1-700
