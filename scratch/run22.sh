#!/bin/bash
# final numbers of round 1 for profiles/: both arms, launch list (one piece under the profiler), modes
O=gpurun_out/r22; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 30 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quant -s 3 -c 1 -f -o $O/r1_quant python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 2 > $O/ncu_quant.log 2>&1
cat > /tmp/cfgs.txt <<EOT
--vbr 3 --brate 128 --steps 10
--vbr 4 --brate 2 --signal sine --steps 10
--vbr 2 --brate 2 --signal sine --steps 5
--quality 0 --steps 5
--quality 2 --steps 10
--quality 5 --steps 10
--quality 7 --steps 10
--streams 1 --frames 4096 --steps 3
--streams 4096 --frames 1 --steps 10
EOT
python scratch/sweep.py /tmp/cfgs.txt | tee $O/modes.txt
tail -3 $O/smoke.log
python -c "
import json
for f in ('bench_ours','bench_ref'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks'), d.get('gpu_launches'))"
