mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -1 gpurun_out/bench_ours.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 215 -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:lg_kernel_quant -s 3 -c 1 -f -o gpurun_out/r1_quant python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lg_kernel_analysis|lg_kernel_scan|lg_kernel_mdct|lg_kernel_pack" -s 12 -c 4 -f -o gpurun_out/r1_others python tools/kbench.py deprecated-lame-mirror_b200/liblamegpu.so 512 8 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
