#!/bin/bash
# handles: calls outstanding per lane (LAMEGPU_HANDLE_DEPTH) against the aggregate rate; checked runs short, rate runs long
O=gpurun_out/r2_depth; mkdir -p $O; rm -f $O/*
REF=oracle/_ref/libmp3lame_ref.so
for d in 1 2 3 4; do
  echo "depth $d" | tee -a $O/handles_depth.txt
  LAMEGPU_HANDLE_DEPTH=$d LAMEGPU_LANES=512 timeout 200 tests/c/bin/handles_mt 512 48 1152 128 $REF 2>&1 | tail -1 | tee -a $O/handles_depth.txt
  if [ $d = 3 ]; then L=("512 256 1152" "512 128 2304" "64 128 1152" "1 256 1152"); else L=("512 256 1152"); fi
  for t in "${L[@]}"; do
    HANDLES_MT_RATE_ONLY=1 LAMEGPU_HANDLE_DEPTH=$d LAMEGPU_LANES=512 timeout 100 tests/c/bin/handles_mt $t 128 $REF 2>&1 | tail -1 | tee -a $O/handles_depth.txt
  done
done
for d in 1 3; do
  echo "timing depth $d" | tee -a $O/handles_depth.txt
  HANDLES_MT_RATE_ONLY=1 LAMEGPU_TIMING=1 LAMEGPU_HANDLE_DEPTH=$d LAMEGPU_LANES=512 timeout 100 tests/c/bin/handles_mt 512 256 1152 128 $REF 2>&1 | grep -E "shared engine closed|UNCHECKED" | tee -a $O/handles_depth.txt
done
# other rate modes at depth 3, checked
for m in "4 2" "3 4" "2 4"; do
  LAMEGPU_HANDLE_DEPTH=3 LAMEGPU_LANES=512 timeout 200 tests/c/bin/handles_mt 256 40 1152 128 $REF $m 2>&1 | tail -1 | tee -a $O/handles_depth.txt
done
