#!/bin/bash
# scaling pass on one 8-GPU box: bench.py as the driver launches it (torchrun) at N = 8, 4, 2, 1, the single-process form at N = 8,
# and BASELINE configs[4] (8M frames over 8 GPUs: 4096 streams x 32 frames per step and GPU, 8 timed steps = 1.05M frames per GPU)
O=gpurun_out/r2_scale; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt; nproc | tee -a $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n8.json 2> $O/n8.err
timeout 900 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --streams 4096 --frames 32 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_n8_config4_8Mframes.json 2> $O/n8c4.err
timeout 600 python bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n8_one_process.json 2> $O/n8p.err
timeout 600 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n4.json 2> $O/n4.err
timeout 600 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/bench_ref_n1.json 2> $O/ref.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_scale/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], d["n_gpus"], "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], "%.3f"%d["ms_per_step"], "%.3f"%d["e2e"].get("ms_per_step",0))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $O/*.err | cut -c1-300
