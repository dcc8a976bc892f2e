#!/bin/bash
O=gpurun_out/r2_24; mkdir -p $O
LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt 512 256 1152 128 oracle/_ref/libmp3lame_ref.so 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|differs" | cut -c1-300 | tee -a $O/handles.txt
LAMEGPU_LANES=512 timeout 400 tests/c/bin/handles_mt 512 200 1152 128 oracle/_ref/libmp3lame_ref.so 0 4 512 2>&1 | grep -E "IDENTICAL|DIFFERENT|FAILED|differs" | cut -c1-300 | tee -a $O/handles.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest.txt
