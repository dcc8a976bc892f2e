#!/bin/bash
# r1 measurement pass 2: smoke (CBR/ABR/VBR), BASELINE configs 2-4 + the two extremes, ncu of the VBR kernel
mkdir -p gpurun_out/r5
O=gpurun_out/r5
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 300 python bench.py --steps 10 --warmup 3 --vbr 4 --brate 2 --signal sine > $O/bench_vbr_512x8.json 2> $O/bench_vbr_512x8.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 --vbr 4 --brate 2 --signal sine > $O/ref_vbr_512x8.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --vbr 3 --brate 128 --no-cpu-baseline > $O/bench_abr_512x8.json 2> $O/bench_abr.err
timeout 400 python bench.py --steps 5 --warmup 3 --signal sine --brate 320 --streams 2048 --frames 32 --no-cpu-baseline > $O/bench_cfg2_2048x32.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --steps 3 --warmup 3 --signal sine --vbr 4 --brate 2 --streams 4096 --frames 64 --no-cpu-baseline > $O/bench_cfg3_4096x64.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --steps 3 --warmup 3 --streams 1 --frames 4096 --no-cpu-baseline > $O/bench_1x4096.json 2> $O/bench_1x4096.err
timeout 300 python bench.py --steps 10 --warmup 3 --streams 4096 --frames 1 --no-cpu-baseline > $O/bench_4096x1.json 2> $O/bench_4096x1.err
timeout 300 python bench.py --steps 5 --warmup 3 --streams 4096 --frames 8 --no-cpu-baseline > $O/bench_4096x8.json 2> $O/bench_4096x8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_kernel_vbr -c 1 -s 3 -o $O/vbr_full -f python bench.py --steps 1 --warmup 3 --vbr 4 --brate 2 --signal sine --no-cpu-baseline > $O/ncu_vbr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_vbr.csv python bench.py --steps 2 --warmup 3 --vbr 4 --brate 2 --signal sine --no-cpu-baseline > $O/ncu_list.log 2>&1
ls -la $O
tail -c 600 $O/smoke.log
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['kernels_ms_per_step'], d['clocks'])
except Exception as e: print('ERR', e)
"; done
