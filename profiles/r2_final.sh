#!/bin/bash
# what the driver runs at round end, in its order; then the handles' throughput record
O=gpurun_out/r2_final; mkdir -p $O; rm -f $O/handles.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -1 $O/bench_ref.json | cut -c1-200
timeout 900 python bench.py > $O/bench_ours.json 2> $O/bench_ours.err; tail -1 $O/bench_ours.json | cut -c1-200
for t in "512 32 1152" "512 256 1152" "512 128 2304" "64 64 1152" "1 256 1152"; do LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt; done
