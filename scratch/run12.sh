#!/bin/bash
for cfg in "--steps 20" "--steps 20" "--streams 640 --frames 8 --steps 10" "--streams 740 --frames 8 --steps 10" "--streams 512 --frames 32 --steps 5" "--streams 100 --frames 16 --steps 10"; do
  timeout 120 python bench.py --warmup 3 --no-cpu-baseline $cfg > /tmp/b.json 2> /tmp/b.err
  python -c "
import json
try:
    d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['kernels_ms_per_step']['quant'],3))
except Exception as e:
    print("$cfg", "FAILED", open("/tmp/b.err").read()[-160:].replace(chr(10)," "))"
done

