#!/bin/bash
# final: both arms, launch list under ncu (the engine falls back to one piece there), env check
O=gpurun_out/r10; mkdir -p $O
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 30 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log | cut -c1-400
env | grep -i "inject\|nsight\|COMPUTE_PROF" | head
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3 python -c "import os; print({k:v for k,v in os.environ.items() if 'INJECT' in k or 'NV_' in k or 'NSIGHT' in k})" 2>&1 | tail -3
for cfg in "--vbr 2 --brate 2 --signal sine" "--vbr 4 --brate 2 --signal sine" "--vbr 3 --brate 128" "--signal sine --brate 320 --streams 2048 --frames 32 --steps 5" "--streams 512 --frames 32 --steps 5" ; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $cfg | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else '') for k,v in d['kernels_ms_per_step'].items()})" >> $O/modes.txt 2>&1
done
cat $O/modes.txt
python -c "
import json
for f in ('bench_ours','bench_ref'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks'), d.get('gpu_launches'))"
