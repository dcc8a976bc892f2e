#!/bin/bash
O=gpurun_out/r2_12; mkdir -p $O
for lib in deprecated-lame-mirror_b200/liblamegpu.so scratch/variants/lib_r122.so; do
timeout 300 python tools/kbench.py $lib 512 8 10 2>&1 | tail -1 | tee -a $O/kbench.txt
LAMEGPU_DEFAULT_CARVEOUT=1 timeout 300 python tools/kbench.py $lib 512 8 10 2>&1 | tail -1 | sed 's/^/defaultcarve /' | tee -a $O/kbench.txt
LAMEGPU_GROUP_NW=0 timeout 300 python tools/kbench.py $lib 512 8 10 2>&1 | tail -1 | sed 's/^/onewarp /' | tee -a $O/kbench.txt
done
