#!/bin/bash
# handles: where the crowd cap should start, deeper settings, more lanes than 512 (rate only)
O=gpurun_out/r2_depth; mkdir -p $O; rm -f $O/handles_more.txt
REF=oracle/_ref/libmp3lame_ref.so
run() { echo "$1" | tee -a $O/handles_more.txt; shift; env HANDLES_MT_RATE_ONLY=1 "$@" 2>&1 | tail -1 | tee -a $O/handles_more.txt; }
H=tests/c/bin/handles_mt
run "64 threads, default (no cap below 128 lanes)" LAMEGPU_LANES=512 timeout 60 $H 64 128 1152 128 $REF
run "64 threads, cap 2 from 32 lanes" LAMEGPU_LANES=512 LAMEGPU_HANDLE_CROWD_LANES=32 timeout 60 $H 64 128 1152 128 $REF
run "128 threads, default (cap 2)" LAMEGPU_LANES=512 timeout 60 $H 128 128 1152 128 $REF
run "128 threads, no cap" LAMEGPU_LANES=512 LAMEGPU_HANDLE_CROWD_CAP=0 timeout 60 $H 128 128 1152 128 $REF
run "256 threads, default (cap 2)" LAMEGPU_LANES=512 timeout 60 $H 256 128 1152 128 $REF
run "256 threads, no cap" LAMEGPU_LANES=512 LAMEGPU_HANDLE_CROWD_CAP=0 timeout 60 $H 256 128 1152 128 $REF
run "512 threads, depth 7 cap 3" LAMEGPU_LANES=512 LAMEGPU_HANDLE_DEPTH=7 LAMEGPU_HANDLE_CROWD_CAP=3 timeout 60 $H 512 256 1152 128 $REF
run "512 threads, depth 7 cap 2" LAMEGPU_LANES=512 LAMEGPU_HANDLE_DEPTH=7 timeout 60 $H 512 256 1152 128 $REF
run "1024 threads, 1024 lanes, default" LAMEGPU_LANES=1024 timeout 60 $H 1024 128 1152 128 $REF
run "1024 threads, 1024 lanes, depth 9 cap 4" LAMEGPU_LANES=1024 LAMEGPU_HANDLE_DEPTH=9 LAMEGPU_HANDLE_CROWD_CAP=4 timeout 60 $H 1024 128 1152 128 $REF
