#!/bin/bash
O=gpurun_out/r2_23; mkdir -p $O
for t in "512 256 1152" "512 256 1152" "512 96 1152"; do
LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|differs" | cut -c1-300 | tee -a $O/handles.txt
done
timeout 600 python -m pytest tests/test_frontend_dropin.py tests/test_shared_handles.py -m gpu -x -q 2>&1 | tail -3
