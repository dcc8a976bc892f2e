#!/bin/bash
# handles: frames per lane per launch (LAMEGPU_HANDLE_FRAMES) x calls outstanding (LAMEGPU_HANDLE_DEPTH), rate only
O=gpurun_out/r2_depth; mkdir -p $O
REF=oracle/_ref/libmp3lame_ref.so
for f in 1 2 3 8; do for d in 3 5; do
  echo "frames per launch $f depth $d" | tee -a $O/handles_frames.txt
  HANDLES_MT_RATE_ONLY=1 LAMEGPU_TIMING=1 LAMEGPU_HANDLE_FRAMES=$f LAMEGPU_HANDLE_DEPTH=$d LAMEGPU_LANES=512 timeout 100 tests/c/bin/handles_mt 512 256 1152 128 $REF 2>&1 | grep -E "shared engine closed|UNCHECKED" | tee -a $O/handles_frames.txt
done; done
