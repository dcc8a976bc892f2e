"""CPU suite: the multi-GPU plan is "shard independent streams, no data-path collective" (SURVEY.md section 8e).
This checks the host-side sharding/gather logic of bench.py with a real 2-process gloo group."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_shard_plan_partitions_streams():
    sys.path.insert(0, ROOT)
    import bench
    for n, w in [(512, 1), (512, 2), (4096, 8), (10, 4), (3, 8)]:
        seen = []
        for r in range(w):
            lo, hi = bench.shard_range(n, r, w)
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import bench
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lo, hi = bench.shard_range(10, r, w)
        frames, ms = bench.reduce_over_ranks(float(hi - lo), 10.0 + r, device="cpu")
        if r == 0:
            assert frames == 10.0 and ms == 11.0, (frames, ms)
            print("OK")
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
