"""Runs the tiled stencils of one level (15x15 blur, 5x5 median) at BASELINE config 2's level-0 size through the diagnostic stage
entry points, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:"k_blur15|k_median5" -c 2 -o gpurun_out/stencils python tools/profile_stencils.py
The width is even, so interior tiles take the TMA tensor-copy path (PF_NO_TMA_TILES=1 for the per-thread-load path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panorama_opticalflow_b200 import stages  # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1100
f = (np.random.default_rng(0).standard_normal((h, w, 2)) * 3).astype(np.float32)
print("blur15", float(np.abs(stages.blur15(f)).mean()))
print("median5", float(np.abs(stages.median5(f)).mean()))
