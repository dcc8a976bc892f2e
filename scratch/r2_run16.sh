#!/bin/bash
O=gpurun_out/r2_16; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for g in 0 1; do
echo "gate $g" | tee -a $O/kbench.txt
LAMEGPU_GATE=$g timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
done
LAMEGPU_GATE=1 timeout 300 python tools/kbench.py $L 592 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
LAMEGPU_GATE=1 timeout 300 python tools/kbench.py $L 4096 8 4 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $O/bench1.json
