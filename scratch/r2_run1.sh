#!/bin/bash
# round 2, run 1: group-form kernel D, NW = 0 (old kernel) / 1 / 2 / 3, timing + hash, then the GPU suite with the default
O=gpurun_out/r2_1; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for nw in 0 1 2 3; do
  LAMEGPU_GROUP_NW=$nw timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | sed "s/^/NW=$nw /" | tee -a $O/kbench.txt
done
for nw in 0 1 2 3; do
  LAMEGPU_GROUP_NW=$nw timeout 300 python tools/kbench.py $L 4096 8 3 2>&1 | tail -1 | sed "s/^/NW=$nw /" | tee -a $O/kbench.txt
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.txt
