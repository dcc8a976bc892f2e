#!/bin/bash
O=gpurun_out/r7; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for P in 1 2 4 8; do
  LAMEGPU_PIECES=$P timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pieces $P', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else '') for k,v in d['kernels_ms_per_step'].items()}, d['gpu_launches'])"
done
LAMEGPU_PIECES=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 4096 --frames 8 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('4096x8 pieces 4', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3))"
LAMEGPU_PIECES=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams 4096 --frames 8 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('4096x8 pieces 1', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3))"
