#!/bin/bash
O=gpurun_out/r2_18; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for c in 57 50 44; do
echo "carveout $c" | tee -a $O/kbench.txt
LAMEGPU_CARVEOUT=$c timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
done
