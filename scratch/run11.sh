#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "--steps 20" "--signal sine --brate 320 --streams 2048 --frames 32 --steps 5" "--streams 4096 --frames 8 --steps 5" "--streams 1024 --frames 8 --steps 10"; do
  timeout 300 python bench.py --warmup 3 --no-cpu-baseline $cfg > /tmp/b.json 2> /tmp/b.err
  python -c "
import json
try:
    d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['kernels_ms_per_step']['quant'],3))
except Exception as e:
    print('$cfg', 'FAILED', e); print(open('/tmp/b.err').read()[-1500:])"
done
