#!/usr/bin/env python
"""bench.py -- Mpix/s of bidirectional PixFlow (NovelViewGeneratorAsymmetricFlow::prepare semantics) on synthetic
4000 x 2000 (rows x cols) overlap pairs, BASELINE.json's metric, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --workload stitch5|four_input        # BASELINE configs 3 / 5 on the reference's own inputs (data/)

A step = one pass of the hot path over one batch of B independent pairs per GPU (weak scaling: every rank processes its own B
pairs, no data-path collective; NCCL only broadcasts the shared base pair during set-up).
  value   inputs and outputs resident in HBM, the synchronous batch call (pf_prepare_bidirectional_batch)
  e2e     the same B pairs per step from pinned HOST buffers to pinned host buffers through the asynchronous C-ABI
          (pf_prepare_bidirectional_batch_async / pf_wait, two slots): the H2D of the images and the D2H of both flow fields of
          EVERY step are inside the timed region; consecutive steps overlap their copies with each other's compute
  single_pair   BASELINE configs[1] read literally: one pair alone on the device (latency-bound)
--impl reference times the CPU restatement of the reference (oracle/, the only other place allowed to run it) on the host
cores of the box, on the SAME configuration: B full-size pairs per step, spread over all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# One hardware work queue per stream of the engine (2 streams per pair in flight): without this, streams share
# queues and kernels of independent pairs serialise behind each other.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: NCCL's banner / debug lines (NCCL_DEBUG=VERSION|INFO) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout must carry exactly ONE JSON line.  Libraries write banners straight to file descriptor 1 (NCCL's "NCCL version ..."
# ignores NCCL_DEBUG_FILE), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

METRIC = "Mpix/s bidirectional flow (2000x4000 pair)"
UNIT = "Mpix/s"
SWEEP_BYTES_PER_PX = 48          # alpha0, alpha1, I0x, I0y, I1x, I1y, blurred(8) read + flow r/w(16)
PAIR_BYTES_PER_PX = 762.15       # whole pair, both directions + warp/blend
BLEND_BYTES_PER_PX = 32.0        # ... of which the warp/blend (combineNovelViews), which the timed flow call does not run


def workload_config(args, world):
    """The workload description shared VERBATIM by both arms (--impl b200 and --impl reference)."""
    amp = args.cols / 12.0 + 1.0
    return {"workload": "%s bidirectional flow (NovelViewGeneratorAsymmetricFlow::prepare semantics) on synthetic %d x %d (rows x cols) "
                        "BGRA overlap pairs, disparity amplitude %.0f px" % (args.preset, args.rows, args.cols, amp),
            "rows": args.rows, "cols": args.cols, "pairs_per_step_per_gpu": args.batch, "preset": args.preset,
            "parallelism": "replicas x%d (pairs are independent; NCCL broadcast of the base pair at set-up only)" % world}


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


def load_sweep_traffic():
    """DRAM bytes per sweep launch from the committed ncu capture of this kernel (profiles/sweep_traffic.json, written by
    tools/summarise_launches.py from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` launch list); None when
    there is no capture for the current kernel."""
    p = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), d.get("source")


def pair_shifts(rank, batch):
    """Row shifts that derive this rank's `batch` pairs from the shared base pair: disjoint across ranks."""
    return [37 * (rank * batch + i) for i in range(batch)]


def aggregate_mpix(world, batch, rows, cols, ms_per_step):
    """Whole-job throughput: units processed by ALL ranks per step / (max-over-ranks) step time."""
    return world * batch * rows * cols / 1e6 / (ms_per_step / 1e3)


class ClockSampler:
    """nvidia-smi samples of THIS rank's GPU only (selected by UUID), started right before and stopped right after a timed
    region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_uuid=None):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.sel = ["-i", gpu_uuid] if gpu_uuid else []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi"] + self.sel + ["--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def discard_so_far(self):
        """call right before the timed region: samples taken while nvidia-smi was starting up do not count"""
        self.f.flush()
        self.skip = os.path.getsize(self.f.name)

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(getattr(self, "skip", 0))
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
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "scope": "this rank's GPU, sampled every 100 ms inside the timed region of `value`"}


def gpu_uuid(local_rank):
    try:
        import torch
        u = str(torch.cuda.get_device_properties(local_rank).uuid)
        return u if u.startswith("GPU-") else "GPU-" + u
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and, by first touch, its pinned host buffers) to the CPUs of the GPU's NUMA node, when the box has more
    than one.  Returns a short description for the bench line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return "single NUMA node (no binding needed)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "bound to NUMA node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # noqa: BLE001 -- best effort, never fatal
        return "not bound (%s)" % type(e).__name__
    return "not bound"


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores, SAME configuration (full-size pairs)
# ---------------------------------------------------------------------------------------------------------
def cpu_pairs(args, n):
    from panorama_opticalflow_b200 import synth
    base = synth.make_pair(args.rows, args.cols, seed=1, amplitude=args.cols / 12.0 + 1.0)
    return [(np.roll(base[0], s, axis=0), np.roll(base[1], s, axis=0)) for s in pair_shifts(0, n)]


def cpu_step(pairs, pct, threads):
    """One step of the CPU arm: every pair of the batch through the oracle's prepare_bidirectional, spread over `threads` host
    threads (ctypes releases the GIL).  Returns (seconds, results)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda p: orc.prepare_bidirectional(p[0], p[1], pct), pairs))
    return time.perf_counter() - t0, res


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    nproc = os.cpu_count() or 1
    threads = max(1, min(nproc, args.cpu_threads, args.batch))
    pct = 20 if args.preset == "pixflow_search_20" else 0
    pairs = cpu_pairs(args, args.batch)
    t1, _ = cpu_step(pairs[:1], pct, 1)            # the reference's own loops are single-threaded: one pair on one thread
    for _ in range(max(0, args.warmup - 1)):       # (counts as the first warm-up step)
        cpu_step(pairs, pct, threads)
    secs = [cpu_step(pairs, pct, threads)[0] for _ in range(args.steps)]
    tot = sum(secs)
    px = args.batch * args.rows * args.cols
    val = px * args.steps / tot / 1e6
    sample = ("%d full-size pairs of %d x %d (rows x cols) per step -- the B200 arm's step -- over %d host threads (nproc %d); "
              "kind 'port': oracle/pixflow_oracle.c, the C restatement of the reference CPU path (the reference itself needs OpenCV C++, "
              "glog and gflags and cannot be built in this image)" % (args.batch, args.rows, args.cols, threads, nproc))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "nproc": nproc, "kind": "port", "sample": sample,
                         "single_thread_value": args.rows * args.cols / t1 / 1e6, "single_thread_s_per_pair": t1},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs 3 and 5: the reference's drivers on the reference's own inputs, canvas resident in HBM
# ---------------------------------------------------------------------------------------------------------
def compare_canvas(result, shipped):
    d = np.abs(result[..., :3].astype(np.int16) - shipped[..., :3].astype(np.int16))
    mse = float(np.mean(d.astype(np.float64) ** 2))
    return {"alpha_identical": bool(np.array_equal(result[..., 3], shipped[..., 3])),
            "psnr_db": float(10 * np.log10(255.0 ** 2 / mse)) if mse > 0 else float("inf"),
            "pixels_bit_equal": float(np.all(d == 0, axis=2).mean()), "pixels_within_1lsb": float(np.all(d <= 1, axis=2).mean())}


def run_stitch_workload(args):
    """--workload stitch5: CPU/main.cpp:55-105 (top + 1..5.tif, five sequential iterations, FinalResult fed back as
    colorImageR without leaving HBM).  --workload four_input: CPU_4Input/main.cpp:54-113 (one pass; --crop95 enables the
    row crop of :82-83 with which the shipped FinalResult.png was produced)."""
    import torch
    import panorama_opticalflow_b200 as pf
    from panorama_opticalflow_b200 import testdata
    set_name = "Test_data_1" if args.workload == "stitch5" else "Test_data_4Input"
    if not testdata.available(set_name):
        raise SystemExit("bench.py: data/%s is missing -- run tools/pack_test_data.py in the build container" % set_name)
    torch.cuda.set_device(0)
    eng = pf.makeOpticalFlowByName(args.preset, device=0)
    lib = pf._lib.load()
    names = ["top", "1", "2", "3", "4", "5"] if args.workload == "stitch5" else ["1", "2", "3", "4"]
    t0 = time.perf_counter()
    host_t = {n: torch.from_numpy(testdata.load(set_name, n)).pin_memory() for n in names}      # pinned host memory
    host = {n: t.numpy() for n, t in host_t.items()}
    load_s = time.perf_counter() - t0
    rows, cols = host[names[0]].shape[:2]

    canvas = [torch.empty((rows, cols, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()

    def one_run(collect=False):
        """-> (FinalResult on the device, [per-iteration ms]) with the H2D of every input inside the timing"""
        per = []
        if args.workload == "stitch5":
            # host pointers go straight into the C-ABI call, which stages them on its own stream; the canvas stays in HBM
            R = host["top"]
            for i in range(1, 6):
                eng.timerStart()
                out = canvas[i % 2]
                pf.stitch_iteration(eng, host[str(i)], R, out=out)
                per.append(eng.timerStop())
                R = out
            return R, per
        eng.timerStart()
        dev = [host_t[n].cuda() for n in names]
        torch.cuda.synchronize()                      # torch's stream is not the engine's
        L, R = pf.four_input_frontend(eng, *dev, device_out=True)
        if args.crop95:
            n95 = int(0.95 * rows)
            L, R = L[:n95].contiguous(), R[:n95].contiguous()
        out = torch.empty_like(L)
        pf.stitch_iteration(eng, L, R, out=out)
        per.append(eng.timerStop())
        return out, per

    for _ in range(max(1, args.warmup)):
        final, _ = one_run()
    torch.cuda.synchronize()
    n0 = lib.pf_kernel_launch_count()
    runs = []
    for _ in range(args.steps):
        final, per = one_run()
        runs.append(per)
    launches = (lib.pf_kernel_launch_count() - n0) // max(1, args.steps)
    per_iter = [float(np.median([r[k] for r in runs])) for k in range(len(runs[0]))]
    total_ms = float(np.median([sum(r) for r in runs]))
    result = final.cpu().numpy()
    shipped = testdata.final_result(set_name)
    cmp_shipped = None
    if shipped is not None and shipped.shape == result.shape:
        cmp_shipped = compare_canvas(result, shipped)
    if args.save_result:
        import cv2
        cv2.imwrite(args.save_result, result)
    line = {"metric": "s per stitch (%s)" % ("Test_data/1: top + 1..5, five iterations" if args.workload == "stitch5" else "Test_data_4Input: single pass"),
            "value": total_ms / 1e3, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": max(1, args.warmup),
            "ms_per_step": total_ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "reference test inputs (data/%s, packed losslessly by tools/pack_test_data.py)" % set_name,
            "config": {"workload": args.workload, "canvas": [int(result.shape[0]), int(result.shape[1])], "preset": args.preset,
                       "crop95": bool(args.crop95), "per_iteration_ms": per_iter,
                       "timed": "every input's H2D from pinned host memory + Stitchtools::prepare + both flows + blend + Gather, canvas resident in HBM between iterations; PNG decode/encode outside (%.1f s to decode the packed inputs on the host)" % load_s,
                       "reference_published": "README.md:10-12: 'less than 30 s' for this stitch on an unspecified GPU"},
            "vs_shipped_final_result": cmp_shipped, "gpu_launches": int(launches)}
    emit(line)
    eng.close()


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
    numa = bind_to_gpu_numa_node(local_rank)
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

    def timed(fn, steps, finish=None):
        barrier()
        n0 = lib.pf_kernel_launch_count()
        eng.timerStart()
        for k in range(steps):
            fn(k)
        if finish is not None:
            finish()
        ms = eng.timerStop()
        n1 = lib.pf_kernel_launch_count()
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n1 - n0

    step_dev = lambda k=0: eng.prepareBidirectionalBatch(dL, dR, oLR, oRL)
    for _ in range(args.warmup):
        step_dev()
    # latency of ONE pair (BASELINE configs[1] read literally), next to the batched throughput
    step_one = lambda k=0: eng.prepareBidirectionalBatch(dL[:1], dR[:1], oLR[:1], oRL[:1])
    step_one()
    one_ms, _ = timed(step_one, 3)
    one_ms /= 3
    sampler = ClockSampler(gpu_uuid(local_rank))
    if rank == 0:
        sampler.start()
    for _ in range(4):                   # the GPU stays under the same load while nvidia-smi starts up (~0.3 s)
        step_dev()
    if rank == 0:
        sampler.discard_so_far()
    ms_total, launches = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    mpix_step = world * B * rows * cols / 1e6
    value = aggregate_mpix(world, B, rows, cols, ms_step)

    # ---- end to end: pinned HOST buffers through the asynchronous C-ABI, two slots (double buffering) ----
    def pinned(shape, dtype):
        return [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(B)]
    hL, hR = [pinned((rows, cols, 4), torch.uint8) for _ in range(2)], [pinned((rows, cols, 4), torch.uint8) for _ in range(2)]
    hLR, hRL = [pinned((rows, cols, 2), torch.float32) for _ in range(2)], [pinned((rows, cols, 2), torch.float32) for _ in range(2)]
    for s in range(2):
        for i in range(B):
            hL[s][i].copy_(dL[i]); hR[s][i].copy_(dR[i])
    nL, nR = [[t.numpy() for t in hL[s]] for s in range(2)], [[t.numpy() for t in hR[s]] for s in range(2)]
    nLR, nRL = [[t.numpy() for t in hLR[s]] for s in range(2)], [[t.numpy() for t in hRL[s]] for s in range(2)]

    def step_e2e(k):
        s = k & 1
        eng.prepareBidirectionalBatchAsync(s, nL[s], nR[s], nLR[s], nRL[s])      # waits for the slot's previous batch first
        if k > 0:
            eng.wait(1 - s)                                                     # results of step k-1 are on the host now

    def finish_e2e():
        eng.wait(0); eng.wait(1)

    for k in range(max(2, min(args.warmup, 3))):
        step_e2e(k)
    finish_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps, finish_e2e)
    e2e_value = mpix_step / (e2e_ms / args.steps / 1e3)
    same = all(np.array_equal(nLR[s][i], oLR[i].cpu().numpy()) and np.array_equal(nRL[s][i], oRL[i].cpu().numpy())
               for i in range(B) for s in range(2))
    # the synchronous call with host buffers (what a caller without the async API sees)
    sync_ms, _ = timed(lambda k: eng.prepareBidirectionalBatch(nL[0], nR[0], nLR[0], nRL[0]), max(2, args.steps // 2))
    sync_ms /= max(2, args.steps // 2)

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
    traffic, traffic_src = load_sweep_traffic()
    flow_bytes_px = PAIR_BYTES_PER_PX - BLEND_BYTES_PER_PX
    roofline = {"bound": "hbm", "kernel": "k_sweep (wavefront Gauss-Seidel)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch_avg": sweep_bytes / max(1, sweep_launches),
                "avg_launch_ms": sweep_ms / max(1, sweep_launches), "launches_per_step": sweep_launches,
                "note": "achieved = algorithmic bytes / sum of per-launch durations, CUDA events on the launching streams, plain stream launches "
                        "(launches of different directions / pairs overlap in time); the sweep is bound by its dependent chain, not by HBM",
                "whole_step_hbm_frac": (flow_bytes_px * B * rows * cols / (ms_step / 1e3) / 1e9) / peak,
                "whole_step_bytes_per_px": flow_bytes_px}
    single = {"ms": one_ms, "value": rows * cols / 1e6 / (one_ms / 1e3), "unit": UNIT,
              "hbm_frac": (flow_bytes_px * rows * cols / (one_ms / 1e3) / 1e9) / peak,
              "note": "BASELINE configs[1] read literally: one 4000 x 2000 pair alone on the device, inputs and outputs resident in HBM"}

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
        step_st = lambda k=0: pf.stitch_iteration(eng, cL, cR, out=final)
        step_st()
        st_ms, st_launches = timed(step_st, 2)
        stitch = {"canvas": "%d x %d, overlap 40 %% of the columns" % (rows, cols), "ms_per_iteration": st_ms / 2,
                  "mpix_s": rows * cols / 1e6 / (st_ms / 2 / 1e3), "kernel_launches": int(st_launches // 2)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        orc.build()
        nproc = os.cpu_count() or 1
        threads = max(1, min(nproc, args.cpu_threads, B))
        pairs = [(dL[i].cpu().numpy(), dR[i].cpu().numpy()) for i in range(B)]
        secs, res = cpu_step(pairs, 20 if args.preset == "pixflow_search_20" else 0, threads)
        parity = all(np.array_equal(res[i][0], oLR[i].cpu().numpy()) and np.array_equal(res[i][1], oRL[i].cpu().numpy()) for i in range(B))
        cpu_baseline = {"value": B * rows * cols / secs / 1e6, "unit": UNIT, "cores": threads, "nproc": nproc, "kind": "port",
                        "sample": "the %d full-size pairs of one step (%d x %d each), over %d host threads, %.1f s wall" % (B, rows, cols, threads, secs),
                        "parity_full_size": bool(parity),
                        "parity_note": "flowLtoR and flowRtoL of all %d pairs of the timed step, GPU == CPU oracle bit for bit" % B}

    cfg = workload_config(args, world)
    cfg.update({"sweep_lanes_per_row": int(os.environ.get("PF_SWEEP_LANES", "2")),
                "l2": "working set ~0.7 GB per pair >> 126 MB L2, no flush needed",
                "timing": "CUDA events (pf_timer_*), barrier+sync both sides, max over ranks",
                "e2e_outputs_match_device_run": bool(same), "host_binding": numa, "stitch_iteration": stitch})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "single_pair": single,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * B * 2 * rows * cols * 4,
                "d2h_bytes_per_step": world * B * 2 * rows * cols * 8, "ms_per_step": e2e_ms / args.steps,
                "api": "pf_prepare_bidirectional_batch_async + pf_wait, two slots: every step uploads its %d image pairs from pinned host memory "
                       "and downloads its %d flow fields; copies of consecutive steps overlap compute" % (B, 2 * B),
                "sync_api_value": mpix_step / (sync_ms / 1e3), "sync_api_ms_per_step": sync_ms,
                "fraction_of_device_resident_value": e2e_value / value,
                "note": "beyond one GPU the step is bounded by the box's host <-> device copy bandwidth (D2H into host memory stops scaling at "
                        "two GPUs; at 8 GPUs one step's copies alone take as long as the e2e step: profiles/r2_pcie_probe_n8.md)"},
        "gpu_launches": int(launches) * world, "clocks": clocks,
    }
    emit(line)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="flow", choices=["flow", "stitch5", "four_input"])
    ap.add_argument("--batch", type=int, default=16, help="independent pairs in flight per GPU per step")
    ap.add_argument("--rows", type=int, default=4000)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--preset", default="pixflow_search_20")
    ap.add_argument("--cpu-threads", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stitch", action="store_true", help="skip the stitching-iteration measurement (config.stitch_iteration)")
    ap.add_argument("--crop95", action="store_true", help="four_input: enable the 0.95 row crop of CPU_4Input/main.cpp:82-83")
    ap.add_argument("--save-result", default=None, help="stitch workloads: write FinalResult as PNG here")
    args = ap.parse_args()
    _claim_stdout()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload != "flow":
        if rank == 0:
            run_stitch_workload(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
