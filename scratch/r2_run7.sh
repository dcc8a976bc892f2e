#!/bin/bash
O=gpurun_out/r2_7; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/pytest.txt
for t in "512 32 1152" "512 8 1152" "64 64 1152" "1 256 1152"; do LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt; done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels_ms_per_step'])"
