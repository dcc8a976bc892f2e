#!/bin/bash
export LAMEGPU_PIECES=1
cd scratch/oldtree
for cfg in "4096 8 5" "512 8 10"; do echo "old $cfg: $(python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so $cfg 2>&1 | tail -1 | cut -c1-200)"; done
cd ../..
for cfg in "4096 8 5" "512 8 10"; do echo "new $cfg: $(python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so $cfg 2>&1 | tail -1 | cut -c1-200)"; done
