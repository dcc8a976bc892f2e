#!/bin/bash
for P in 1 2 4; do
  LAMEGPU_PIECES=$P timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pieces $P', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else '') for k,v in d['kernels_ms_per_step'].items()}, d['gpu_launches'])"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
