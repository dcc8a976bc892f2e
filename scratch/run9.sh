#!/bin/bash
for V in "LAMEGPU_PIECES=4" "LAMEGPU_PIECES=4 LAMEGPU_NO_PRIO=1" "LAMEGPU_PIECES=8" "LAMEGPU_PIECES=2" "LAMEGPU_PIECES=3"; do
  env $V timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$V', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else '') for k,v in d['kernels_ms_per_step'].items()})"
done
