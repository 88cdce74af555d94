"""Runs one level-sized sweep (prep + wavefront) through the diagnostic stage entry point, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 2 -o gpurun_out/sweep python tools/profile_sweep.py [h w reps]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panorama_opticalflow_b200 import stages  # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
G0 = np.stack([np.sin(xx / 7) * 0.05, np.cos(yy / 9) * 0.05], 2).astype(np.float32)
G1 = np.stack([np.sin((xx + 3) / 7) * 0.05, np.cos((yy + 1) / 9) * 0.05], 2).astype(np.float32)
flow = np.stack([-3 + 0.3 * np.sin(yy / 50), 0.3 * np.cos(xx / 40)], 2).astype(np.float32)
flow += (rng.standard_normal((h, w, 2)) * 0.05).astype(np.float32)
one = np.full((h, w), float(os.environ.get("PF_PROFILE_ALPHA", "1")), np.float32)
for d in ([+1, -1] * reps)[:reps]:
    t = time.time()
    out = stages.sweep(one, one, G0, G1, flow, flow, d)
    print("sweep dir %+d: %.1f ms wall (incl. copies)" % (d, (time.time() - t) * 1e3), float(np.abs(out - flow).mean()))
