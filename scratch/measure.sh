mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -1 gpurun_out/bench_ours.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -32 gpurun_out/launches.csv | cut -d, -f5,12- | head -40
