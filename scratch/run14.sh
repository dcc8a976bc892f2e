#!/bin/bash
export LAMEGPU_PIECES=1
for v in base restrict; do for cfg in "4096 8 5" "2048 32 3" "512 8 10"; do echo "$v $cfg: $(python tools/kbench.py scratch/variants/lib_$v.so $cfg 2>&1 | tail -1 | cut -c1-200)"; done; done
