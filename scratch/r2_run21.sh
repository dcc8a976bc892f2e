#!/bin/bash
O=gpurun_out/r2_21; mkdir -p $O
for t in "512 32 1152" "512 32 1152" "512 32 2304" "512 96 1152" "512 256 1152" "64 64 1152" "200 40 700" "1 256 1152"; do
LAMEGPU_TIMING=1 LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|shared engine|differs" | cut -c1-300 | tee -a $O/handles5.txt
done
LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt 512 24 1152 2 oracle/_ref/libmp3lame_ref.so 4 2 2>&1 | tail -1 | tee -a $O/handles5.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
