"""Host <-> device copy bandwidth of the box with 1, 2, 4 ... N GPUs copying at the same time (no compute).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe.py

What bench.py's `e2e` moves per step and GPU is 1.02 GB host -> device (16 BGRA pairs) and 2.05 GB device -> host (32 flow fields)
from / to pinned host memory.  This probe moves exactly those byte counts per round -- H2D alone, D2H alone, both at once on two
streams -- with the first n ranks active and the others idle, and prints the aggregate GB/s per n.  It answers whether the e2e
curve of `bench.py --gpus 8` is bounded by the platform (PCIe switches / host memory) or by the engine's host path.
"""
import json
import os
import time

import torch
import torch.distributed as dist

H2D_BYTES = 16 * 2 * 4000 * 2000 * 4
D2H_BYTES = 16 * 2 * 4000 * 2000 * 8


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(0)
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def round_(mode):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_dn):
                h_out.copy_(d_out, non_blocking=True)

    results = []
    n = 1
    sizes = []
    while n <= world:
        sizes.append(n)
        n *= 2
    if sizes[-1] != world:
        sizes.append(world)
    reps = 4
    for n in sizes:
        for mode in ("h2d", "d2h", "both"):
            active = rank < n
            if active:
                round_(mode)           # warm-up
            barrier()
            t0 = time.perf_counter()
            if active:
                for _ in range(reps):
                    round_(mode)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            nbytes = reps * n * ((H2D_BYTES if mode != "d2h" else 0) + (D2H_BYTES if mode != "h2d" else 0))
            results.append({"gpus_copying": n, "mode": mode, "aggregate_gb_s": nbytes / dt / 1e9, "per_gpu_gb_s": nbytes / dt / 1e9 / n,
                            "ms_per_round": dt / reps * 1e3})
            barrier()
    if rank == 0:
        print(json.dumps({"h2d_bytes_per_round_per_gpu": H2D_BYTES, "d2h_bytes_per_round_per_gpu": D2H_BYTES, "nproc": os.cpu_count(),
                          "results": results}, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
