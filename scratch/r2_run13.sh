#!/bin/bash
O=gpurun_out/r2_13; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for S in 148 296 444 512 592; do timeout 300 python tools/kbench.py $L $S 8 6 2>&1 | tail -1 | cut -c1-200 | tee -a $O/kbench.txt; done
