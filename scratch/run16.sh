#!/bin/bash
export LAMEGPU_PIECES=1 LAMEGPU_DEBUG_OCC=1
echo "carveout: $(python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 4096 8 5 2>&1 | grep -v '^$' | tail -3 | cut -c1-180 | tr '\n' '|')"
export LAMEGPU_NO_CARVEOUT=1
echo "nocarve: $(python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 4096 8 5 2>&1 | tail -3 | cut -c1-180 | tr '\n' '|')"
