#!/bin/bash
# round-2 measurement pass on one B200 (the files named in profiles/README.md come from here)
O=gpurun_out/r2_measure; mkdir -p $O
L=deprecated-lame-mirror_b200/liblamegpu.so
python bench.py --steps 20 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err; tail -1 $O/bench_ours.json | cut -c1-200
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -1 $O/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
grep -c lg_kernel $O/launches.csv
ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quantg -s 6 -c 1 -f -o $O/r2_quantg python tools/kbench.py $L 512 8 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lg_kernel_analysis|lg_kernel_scan|lg_kernel_mdct|lg_kernel_pack" -s 24 -c 4 -f -o $O/r2_others python tools/kbench.py $L 512 8 2 > /dev/null 2>&1
ls -la $O/*.ncu-rep
for t in "512 32 1152" "512 32 2304" "64 64 1152" "1 256 1152"; do LAMEGPU_LANES=512 timeout 300 tests/c/bin/handles_mt $t 128 oracle/_ref/libmp3lame_ref.so 2>&1 | tail -1 | tee -a $O/handles.txt; done
if [ -n "$FULL_SWEEP" ]; then LAMEGPU_FULL_SWEEP=1 timeout 1200 python -m pytest tests/test_frontend_dropin.py -m gpu -x -q 2>&1 | tail -3 | tee $O/full_sweep.txt; fi
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
