import json, subprocess, sys
cfgs = [l for l in open(sys.argv[1]).read().splitlines() if l.strip()]
for cfg in cfgs:
    env = {}
    args = cfg.split()
    while args and "=" in args[0] and not args[0].startswith("--"):
        k, v = args.pop(0).split("=", 1); env[k] = v
    import os
    r = subprocess.run([sys.executable, "bench.py", "--warmup", "3", "--no-cpu-baseline"] + args, capture_output=True, text=True, timeout=400, env=dict(os.environ, **env))
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(cfg, "| value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "quant", round(d["kernels_ms_per_step"]["quant"], 3), flush=True)
    except Exception:
        print(cfg, "| FAILED", r.stderr[-200:].replace("\n", " "), flush=True)
