#!/bin/bash
O=gpurun_out/r2_21; mkdir -p $O
for n in 512 256 128 64; do
echo "lanes $n" | tee -a $O/handles.txt
LAMEGPU_LANES=$n timeout 300 tests/c/bin/handles_mt 512 32 1152 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt
LAMEGPU_LANES=$n timeout 300 tests/c/bin/handles_mt 512 32 1152 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt
done
