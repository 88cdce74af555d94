"""One device-resident stitching iteration (CPU/main.cpp:72-95) on a synthetic canvas, timed per stage with CUDA events
through the public API; run it under `ncu -k regex:'stitch|box|four'` for the launch list of the stitching kernels.

    python tools/stitch_profile.py [rows cols] [--iters N]
"""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import panorama_opticalflow_b200 as pf  # noqa: E402
from panorama_opticalflow_b200 import synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rows, cols = (int(args[0]), int(args[1])) if len(args) >= 2 else (4000, 2000)
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 3
    # a tile of synthetic texture repeated over the canvas (the generator is slow for 36 Mpx)
    tr, tc = min(rows, 1000), min(cols, 1000)
    L0, R0 = synth.make_pair(tr, tc, seed=1, amplitude=40.0)
    reps = ((rows + tr - 1) // tr, (cols + tc - 1) // tc, 1)
    L = torch.from_numpy(np.tile(L0, reps)[:rows, :cols].copy()).cuda()
    R = torch.from_numpy(np.tile(R0, reps)[:rows, :cols].copy()).cuda()
    x = torch.arange(cols, device="cuda")[None, :, None]
    L[..., 3:4] = torch.where(x < int(0.7 * cols), 255, 0).to(torch.uint8).expand(rows, cols, 1)
    R[..., 3:4] = torch.where(x > int(0.3 * cols), 255, 0).to(torch.uint8).expand(rows, cols, 1)
    L *= (L[..., 3:4] > 0)
    R *= (R[..., 3:4] > 0)
    eng = pf.makeOpticalFlowByName("pixflow_search_20")
    out = torch.empty_like(L)
    st = pf.Stitchtools(eng)
    res = {"rows": rows, "cols": cols}
    for name, fn in (("stitch_iteration", lambda: pf.stitch_iteration(eng, L, R, out=out)),):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        res[name + "_ms"] = (time.perf_counter() - t0) / iters * 1e3
    # stage split through the step-by-step mirror's device entry points (host arrays: includes copies, for orientation only)
    hL, hR = L.cpu().numpy(), R.cpu().numpy()
    t0 = time.perf_counter(); st.prepare(hL, hR); res["Stitchtools.prepare_host_ms"] = (time.perf_counter() - t0) * 1e3
    res["blocks_smoothed"] = int((st.MergedDis[::max(1, min(rows, cols) // 200), ::max(1, min(rows, cols) // 200)] > min(rows, cols) // 200).sum())
    print(json.dumps(res))
    eng.close()


if __name__ == "__main__":
    main()
