#!/bin/bash
O=gpurun_out/r2_22; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
for sp in 1 2 4; do
echo "split $sp" | tee -a $O/kbench.txt
LAMEGPU_ANA_SPLIT=$sp timeout 300 python tools/kbench.py $L 512 8 10 2>&1 | tail -1 | cut -c1-260 | tee -a $O/kbench.txt
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > $O/bench1.json; python -c "
import json; d=json.loads(open('$O/bench1.json').read()); print('value %.4g ms %.3f  e2e %.4g ms %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']), d['kernels_ms_per_step']['alone'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest.txt
