#!/bin/bash
O=gpurun_out/r2_4; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
LAMEGPU_GROUP_NW=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quantg -s 3 -c 1 -f -o $O/quantg_nw2 python tools/kbench.py $L 512 8 2 > $O/ncu_nw2.log 2>&1
