#!/bin/bash
O=gpurun_out/r2_20; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
ncu --set full --clock-control none --import-source on -k regex:"lg_kernel_scan|lg_kernel_analysis" -s 24 -c 2 -f -o $O/r2_ab python tools/kbench.py $L 512 8 2 > /dev/null 2>&1
ls -la $O
