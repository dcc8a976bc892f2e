#!/bin/bash
O=gpurun_out/r2_9; mkdir -p $O
for t in "512 32 1152" "512 32 1152" "512 64 9216" "512 64 2304" "64 64 1152" "1 256 1152"; do LAMEGPU_TIMING=1 LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|shared engine" | tee -a $O/handles.txt; done
