#!/bin/bash
O=gpurun_out/r2_11; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_visible or side_by_side" 2>&1 | tail -4 | tee $O/pytest.txt
timeout 600 python bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1proc_2gpu.json 2> $O/bench_1proc.err; tail -2 $O/bench_1proc.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_torchrun_2gpu.json 2> $O/bench_torchrun.err; tail -2 $O/bench_torchrun.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1gpu.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_1gpu","bench_1proc_2gpu","bench_torchrun_2gpu"):
    try:
        d=json.loads(open("gpurun_out/r2_11/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "%.3g"%d["value"], "%.3g"%d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
