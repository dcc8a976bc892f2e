#!/bin/bash
O=gpurun_out/r2_18; mkdir -p $O
for L in scratch/variants/lib_u2.so scratch/variants/lib_u3.so; do
timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
done
