#!/bin/bash
O=gpurun_out/r2_21; mkdir -p $O
for t in "512 32 1152" "512 32 1152" "512 32 2304" "64 64 1152"; do
LAMEGPU_TIMING=1 LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|shared engine" | cut -c1-300 | tee -a $O/handles3.txt
done
