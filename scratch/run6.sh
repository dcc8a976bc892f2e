#!/bin/bash
# r1 final measurement: smoke, both bench arms, launch list, ncu --set full of all kernels, e2e breakdown, new-mode timings
O=gpurun_out/r6; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err
LAMEGPU_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/e2e_timing.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 30 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quant -s 3 -c 1 -f -o $O/r1_quant python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 2 > $O/ncu_quant.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lg_kernel_analysis|lg_kernel_scan|lg_kernel_mdct|lg_kernel_pack" -s 12 -c 4 -f -o $O/r1_others python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 2 > $O/ncu_others.log 2>&1
for cfg in "--vbr 2 --brate 2 --signal sine" "--vbr 4 --brate 2 --signal sine" "--quality 0" "--quality 2" "--quality 5" "--quality 7"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $cfg | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['e2e']['value'], d['kernels_ms_per_step'])" >> $O/modes.txt 2>&1
done
ls -la $O; cat $O/smoke.log | tail -5; cat $O/modes.txt; tail -3 $O/e2e_timing.txt
python -c "
import json
for f in ('bench_ours','bench_ref'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('e2e',{}).get('value'), d.get('clocks'))"
