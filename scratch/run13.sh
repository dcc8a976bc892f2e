#!/bin/bash
O=gpurun_out/r13; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 30 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
python scratch/sweep.py scratch/cfgs13.txt | tee $O/modes.txt
python -c "
import json
for f in ('bench_ours','bench_ref'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks'), d.get('gpu_launches'))"
