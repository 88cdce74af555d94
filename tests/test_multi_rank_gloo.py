"""N>1 host logic on CPU: world_size-2 gloo run of the bench's sharding / reduction plumbing (pairs are
independent units sharded over ranks with no data-path collective; only the shared base pair is broadcast and the
per-rank times are max-reduced)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
import bench

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# shared base pair: rank 0 generates, everybody receives the same bytes (bench.py uses the same call on NCCL)
base = torch.zeros((2, 40, 48, 4), dtype=torch.uint8)
if rank == 0:
    base.copy_(torch.from_numpy(np.random.default_rng(0).integers(0, 256, (2, 40, 48, 4), dtype=np.uint8)))
dist.broadcast(base, src=0)
B = 3
shifts = bench.pair_shifts(rank, B)
assert len(set(shifts)) == B
allshifts = [None] * world
dist.all_gather_object(allshifts, shifts)
flat = [s for ss in allshifts for s in ss]
assert len(set(flat)) == world * B, "ranks must work on disjoint pairs"
# max-over-ranks timing
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 10.0 + world - 1
chk = torch.tensor([float(base.sum())], dtype=torch.float64)
lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
assert all(x.item() == lst[0].item() for x in lst)
# whole-job throughput = units of all ranks / max time
assert abs(bench.aggregate_mpix(world, B, 4000, 2000, 100.0) - world * B * 8.0 / 0.1) < 1e-6
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_reference_arm_only_rank0_prints(tmp_path):
    """--impl reference under torchrun: rank 0 alone runs and prints, the other ranks exit 0 without work."""
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--rows", "256", "--cols", "128", "--cpu-threads", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
