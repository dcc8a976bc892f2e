#!/bin/bash
O=gpurun_out/r2_8; mkdir -p $O
for t in "512 32 1152" "512 32 1152" "512 8 1152" "64 64 1152" "512 32 4608"; do LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt; done
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -20 | tee $O/pytest.txt
