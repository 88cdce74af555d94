#!/usr/bin/env python
"""bench.py -- Mpix/s of bidirectional PixFlow (NovelViewGeneratorAsymmetricFlow::prepare semantics) on synthetic
4000 x 2000 (rows x cols) overlap pairs, BASELINE.json's metric, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of B independent pairs per GPU (weak scaling: every rank
processes its own B pairs, no data-path collective; NCCL only broadcasts the shared base pair during set-up).
`value` is measured with inputs and outputs resident in HBM; `e2e` through the same C-ABI call with pinned HOST
buffers (H2D of the images and D2H of both flow fields inside the timed region).
--impl reference times the CPU restatement of the reference (oracle/, the only other place allowed to run it)
on the host cores of the box.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# One hardware work queue per stream of the engine (3 streams per pair in flight): without this, streams share
# queues and kernels of independent pairs serialise behind each other.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: NCCL's banner / debug lines (NCCL_DEBUG=VERSION|INFO) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mpix/s bidirectional flow (2000x4000 pair)"
UNIT = "Mpix/s"
SWEEP_BYTES_PER_PX = 48          # alpha0, alpha1, I0x, I0y, I1x, I1y, blurred(8) read + flow r/w(16)
PAIR_BYTES_PER_PX = 762.15       # whole pair, both directions + warp/blend
SWEEP_DRAM_BYTES_PER_LAUNCH = 13.18e6  # measured once with ncu (see roofline.traffic_source); algorithmic avg is 15.0e6


def pyramid_levels(rows, cols, pad):
    """CPU/PixFlow.hpp:80-81, :137-151 (fp32 size arithmetic)"""
    f32 = np.float32
    w = int(f32(cols + 2 * pad) * f32(0.5))
    h = int(f32(rows) * f32(0.5))
    out = [(w, h)]
    while True:
        nw, nh = int(f32(out[-1][0]) * f32(0.9) + f32(0.5)), int(f32(out[-1][1]) * f32(0.9) + f32(0.5))
        if nw <= 24 or nh <= 24:
            break
        out.append((nw, nh))
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def pair_shifts(rank, batch):
    """Row shifts that derive this rank's `batch` pairs from the shared base pair: disjoint across ranks."""
    return [37 * (rank * batch + i) for i in range(batch)]


def aggregate_mpix(world, batch, rows, cols, ms_per_step):
    """Whole-job throughput: units processed by ALL ranks per step / (max-over-ranks) step time."""
    return world * batch * rows * cols / 1e6 / (ms_per_step / 1e3)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_step(pairs, pct, threads):
    """Each thread runs one full prepare_bidirectional of the CPU oracle (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda p: orc.prepare_bidirectional(p[0], p[1], pct), pairs))
    return time.perf_counter() - t0


def cpu_sample_pairs(rows, cols, scale, threads):
    from panorama_opticalflow_b200 import synth
    r, c = max(64, rows // scale), max(64, cols // scale)
    base = synth.make_pair(r, c, seed=1, amplitude=c / 12.0 + 1.0)
    return [(np.roll(base[0], 37 * i, axis=0), np.roll(base[1], 37 * i, axis=0)) for i in range(threads)], r, c


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    threads = min(os.cpu_count() or 1, args.cpu_threads)
    pairs, r, c = cpu_sample_pairs(args.rows, args.cols, args.ref_scale, threads)
    for _ in range(min(args.warmup, 1)):
        cpu_step(pairs, 20, threads)
    secs = [cpu_step(pairs, 20, threads) for _ in range(args.steps)]
    tot = sum(secs)
    val = threads * r * c * args.steps / tot / 1e6
    sample = "%d pairs of %d x %d (rows x cols; 1/%d-scale copies of the workload) per step, one per host thread" % (threads, r, c, args.ref_scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pixflow_search_20 bidirectional flow (prepare semantics), CPU restatement of the reference (oracle/pixflow_oracle.c; the reference itself needs OpenCV C++ and cannot be built here)",
                   "rows": r, "cols": c, "pairs_per_step": threads},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import panorama_opticalflow_b200 as pf
    from panorama_opticalflow_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows, cols, B = args.rows, args.cols, args.batch
    amp = cols / 12.0 + 1.0
    # shared base pair: generated on rank 0, NCCL-broadcast to the other ranks (north_star: "NCCL broadcast of
    # the shared base image only"); every rank then derives its own B pairs by rolling rows.
    base = torch.empty((2, rows, cols, 4), dtype=torch.uint8, device="cuda")
    if rank == 0:
        L, R = synth.make_pair(rows, cols, seed=1, amplitude=amp)
        base[0].copy_(torch.from_numpy(L)); base[1].copy_(torch.from_numpy(R))
    if dist is not None:
        dist.broadcast(base, src=0)
    shifts = pair_shifts(rank, B)
    dL = [torch.roll(base[0], s, dims=0).contiguous() for s in shifts]
    dR = [torch.roll(base[1], s, dims=0).contiguous() for s in shifts]
    oLR = [torch.empty((rows, cols, 2), dtype=torch.float32, device="cuda") for _ in range(B)]
    oRL = [torch.empty((rows, cols, 2), dtype=torch.float32, device="cuda") for _ in range(B)]
    torch.cuda.synchronize()

    eng = pf.makeOpticalFlowByName(args.preset, device=local_rank)
    lib = pf._lib.load()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        n0 = lib.pf_kernel_launch_count()
        eng.timerStart()
        for _ in range(steps):
            fn()
        ms = eng.timerStop()
        n1 = lib.pf_kernel_launch_count()
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n1 - n0

    step_dev = lambda: eng.prepareBidirectionalBatch(dL, dR, oLR, oRL)
    for _ in range(args.warmup):
        step_dev()
    # latency of ONE pair (BASELINE configs[1] read literally), for the record next to the batched throughput
    step_one = lambda: eng.prepareBidirectionalBatch(dL[:1], dR[:1], oLR[:1], oRL[:1])
    step_one()
    one_ms, _ = timed(step_one, 2)
    one_ms /= 2
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    mpix_step = world * B * rows * cols / 1e6
    value = aggregate_mpix(world, B, rows, cols, ms_step)

    # ---- end to end through the same C-ABI call with pinned HOST buffers ----
    hL = [torch.empty((rows, cols, 4), dtype=torch.uint8).pin_memory() for _ in range(B)]
    hR = [torch.empty((rows, cols, 4), dtype=torch.uint8).pin_memory() for _ in range(B)]
    for i in range(B):
        hL[i].copy_(dL[i]); hR[i].copy_(dR[i])
    hLR = [torch.empty((rows, cols, 2), dtype=torch.float32).pin_memory() for _ in range(B)]
    hRL = [torch.empty((rows, cols, 2), dtype=torch.float32).pin_memory() for _ in range(B)]
    nL, nR = [t.numpy() for t in hL], [t.numpy() for t in hR]
    nLR, nRL = [t.numpy() for t in hLR], [t.numpy() for t in hRL]
    step_e2e = lambda: eng.prepareBidirectionalBatch(nL, nR, nLR, nRL)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps)
    e2e_value = mpix_step / (e2e_ms / args.steps / 1e3)
    same = all(np.array_equal(nLR[i], oLR[i].cpu().numpy()) for i in range(B))

    # ---- roofline of the dominant kernel (the wavefront sweep), CUDA events on the launching streams ----
    eng.setSweepTiming(True)
    step_dev()
    step_dev()
    sweep_ms, sweep_launches = eng.lastSweepMs()
    eng.setSweepTiming(False)
    levels = pyramid_levels(rows, cols, cols // 20)
    px_sum = sum(w * h for w, h in levels)
    sweep_bytes = B * 2 * 2 * SWEEP_BYTES_PER_PX * px_sum          # per step: B pairs x 2 directions x 2 sweeps
    peak, peak_src = load_peaks()
    achieved = sweep_bytes / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_sweep (wavefront Gauss-Seidel)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": SWEEP_DRAM_BYTES_PER_LAUNCH, "peak_source": peak_src,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 148 sweep launches of one 4000x2000 pair (ncu, profiles/r1_launches_head_4000x2000.csv)",
                "algorithmic_bytes_per_launch_avg": sweep_bytes / max(1, sweep_launches),
                "avg_launch_ms": sweep_ms / max(1, sweep_launches), "launches_per_step": sweep_launches,
                "note": "sum of per-launch durations (launches of different directions/pairs overlap in time)",
                "whole_step_hbm_frac": (PAIR_BYTES_PER_PX * B * rows * cols / (ms_step / 1e3) / 1e9) / peak}

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- the whole stitching iteration around the path (CPU/main.cpp:72-95: Stitchtools::prepare -> flow -> blend -> Gather)
    # on one canvas of the workload's size, device-resident; reported next to the headline, not part of it ----
    stitch = None
    if world == 1 and not args.no_stitch:
        x = torch.arange(cols, device="cuda")[None, :, None]
        cL, cR = dL[0].clone(), dR[0].clone()
        cL[..., 3:4] = torch.where(x < int(0.7 * cols), 255, 0).to(torch.uint8).expand(rows, cols, 1)
        cR[..., 3:4] = torch.where(x > int(0.3 * cols), 255, 0).to(torch.uint8).expand(rows, cols, 1)
        cL *= (cL[..., 3:4] > 0)
        cR *= (cR[..., 3:4] > 0)
        final = torch.empty_like(cL)
        step_st = lambda: pf.stitch_iteration(eng, cL, cR, out=final)
        step_st()
        st_ms, st_launches = timed(step_st, 2)
        stitch = {"canvas": "%d x %d, overlap 40 %% of the columns" % (rows, cols), "ms_per_iteration": st_ms / 2,
                  "mpix_s": rows * cols / 1e6 / (st_ms / 2 / 1e3), "kernel_launches": int(st_launches // 2)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        orc.build()
        threads = min(os.cpu_count() or 1, args.cpu_threads)
        pairs, r, c = cpu_sample_pairs(rows, cols, args.ref_scale, threads)
        secs = cpu_step(pairs, 20, threads)
        cpu_baseline = {"value": threads * r * c / secs / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d pairs of %d x %d (1/%d-scale copies of the workload), one per host thread, %.1f s wall" % (threads, r, c, args.ref_scale, secs)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s bidirectional flow (NovelViewGeneratorAsymmetricFlow::prepare semantics) on synthetic %d x %d (rows x cols) BGRA overlap pairs, disparity amplitude %.0f px" % (args.preset, rows, cols, amp),
                   "rows": rows, "cols": cols, "pairs_per_step_per_gpu": B, "single_pair_latency_ms": one_ms,
                   "single_pair_mpix_s": rows * cols / 1e6 / (one_ms / 1e3), "sweep_lanes_per_row": int(os.environ.get("PF_SWEEP_LANES", "2")),
                   "parallelism": "replicas x%d (pairs are independent; NCCL broadcast of the base pair at set-up only)" % world,
                   "l2": "working set ~0.7 GB per pair >> 126 MB L2, no flush needed", "timing": "CUDA events (pf_timer_*), barrier+sync both sides, max over ranks",
                   "e2e_outputs_match_device_run": bool(same), "stitch_iteration": stitch},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * B * 2 * rows * cols * 4,
                "d2h_bytes_per_step": world * B * 2 * rows * cols * 8, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches) * world, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="independent pairs in flight per GPU per step")
    ap.add_argument("--rows", type=int, default=4000)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--preset", default="pixflow_search_20")
    ap.add_argument("--ref-scale", type=int, default=2, help="CPU arm: linear down-scale of the sample pairs")
    ap.add_argument("--cpu-threads", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stitch", action="store_true", help="skip the stitching-iteration measurement (config.stitch_iteration)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
