#!/bin/bash
# long streams through the handles in every rate mode, every stream compared with the reference library
O=gpurun_out/r2_25; mkdir -p $O
run() { echo "== $*" | tee -a $O/long.txt; LAMEGPU_LANES=512 timeout 600 tests/c/bin/handles_mt "$@" 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|differs|failed" | cut -c1-200 | tee -a $O/long.txt; }
run 512 150 1152 128 oracle/_ref/libmp3lame_ref.so 4 2 1024
run 512 150 1152 128 oracle/_ref/libmp3lame_ref.so 3 4 1536
run 512 150 1152 128 oracle/_ref/libmp3lame_ref.so 2 4 2048
run 512 150 1152 320 oracle/_ref/libmp3lame_ref.so 0 4 2560
run 512 150 1152 64 oracle/_ref/libmp3lame_ref.so 0 4 3072
run 512 150 1000 192 oracle/_ref/libmp3lame_ref.so 0 4 3584
