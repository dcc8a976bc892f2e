#!/bin/bash
O=gpurun_out/r2_5; mkdir -p $O
L=scratch/variants/lib_gsm.so
for nw in 3 2 1; do
  LAMEGPU_GROUP_NW=$nw timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | sed "s/^/NW=$nw /" | tee -a $O/kbench.txt
done
LAMEGPU_GROUP_NW=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quantg -s 3 -c 1 -f -o $O/gsm_nw3 python tools/kbench.py $L 512 8 2 > $O/ncu_nw3.log 2>&1
LAMEGPU_GROUP_NW=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quantg -s 3 -c 1 -f -o $O/gsm_nw2 python tools/kbench.py $L 512 8 2 > $O/ncu_nw2.log 2>&1
