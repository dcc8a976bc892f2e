#!/bin/bash
export LAMEGPU_PIECES=1
O=$PWD/gpurun_out/r17; mkdir -p $O
cd scratch/oldtree
timeout 300 ncu --set full --clock-control none -k regex:lg_kernel_quant -s 3 -c 1 -f -o $O/old_quant python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 4096 8 2 > $O/old.log 2>&1
cd ../..
timeout 300 ncu --set full --clock-control none -k regex:lg_kernel_quant -s 3 -c 1 -f -o $O/new_quant python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 4096 8 2 > $O/new.log 2>&1
ls -la $O
