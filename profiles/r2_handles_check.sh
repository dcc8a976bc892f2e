#!/bin/bash
O=gpurun_out/r2_depth; mkdir -p $O; rm -f $O/handles_check.txt
REF=oracle/_ref/libmp3lame_ref.so
for t in "64 128 1152" "200 48 1152" "40 64 700"; do LAMEGPU_LANES=512 timeout 60 tests/c/bin/handles_mt $t 128 $REF 2>&1 | tail -1 | tee -a $O/handles_check.txt; done
timeout 90 python -m pytest tests/test_shared_handles.py -x -q -m gpu 2>&1 | tail -2 | tee -a $O/handles_check.txt
