set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for v in old u11 u31 u33 u99; do timeout 300 python tools/kbench.py scratch/variants/lib_$v.so 512 8 10; done 2>&1 | grep -v "^+" | tee gpurun_out/kbench1.txt
for v in old u33; do timeout 300 python tools/kbench.py scratch/variants/lib_$v.so 4096 8 5; done 2>&1 | grep -v "^+" | tee -a gpurun_out/kbench1.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest1.txt
