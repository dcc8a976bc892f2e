#!/bin/bash
O=gpurun_out/r2_2; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for nw in 3 1; do
LAMEGPU_GROUP_NW=$nw timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quantg -s 3 -c 1 -f -o $O/quantg_nw$nw python tools/kbench.py $L 512 8 2 > $O/ncu_nw$nw.log 2>&1
done
ls -la $O
