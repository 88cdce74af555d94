#!/usr/bin/env python
"""How often does a pixel of the Gauss-Seidel sweep adopt a neighbour's proposal?  (CPU, analysis only.)

    python tools/sweep_adoption_stats.py [rows cols]

Builds the oracle with -DORC_STATS (oracle/pixflow_oracle.c records, per sweep, which pixels adopt the left / up proposal) and
runs the bench's pair at a reduced size.  Motivation: if a pixel's neighbours usually kept their own gradient step r0 -- which
the parallel prep pass knows in advance -- the sweep could evaluate the two candidates speculatively in the prep pass and the
dependent chain would shrink to a compare.  Measured (seed 1, amplitude cols/12 + 1): 61 % of the updatable pixels adopt a
proposal, and only 0.3-0.5 % of the warp-steps (16 rows on an anti-diagonal) have no adopting neighbour at all, so the
speculation would almost never hold for a whole warp.  Not built."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panorama_opticalflow_b200 import synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 500
so = os.path.join(tempfile.gettempdir(), "liborc_stats.so")
subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-DORC_STATS",
                       "-o", so, os.path.join(ROOT, "oracle", "pixflow_oracle.c"), "-lm", "-lpthread"])
lib = C.CDLL(so)
st = (C.c_double * 8).in_dll(lib, "orc_stats")
L, R = synth.make_pair(rows, cols, seed=1, amplitude=cols / 12.0 + 1.0)
fLR = np.empty((rows, cols, 2), np.float32)
fRL = np.empty((rows, cols, 2), np.float32)
lib.orc_prepare_bidirectional(L.ctypes.data_as(C.c_void_p), C.c_size_t(cols * 4), R.ctypes.data_as(C.c_void_p), C.c_size_t(cols * 4),
                              rows, cols, 20, fLR.ctypes.data_as(C.c_void_p), fRL.ctypes.data_as(C.c_void_p))
print("pair %d x %d: %.0f updatable pixel visits, %.1f %% adopt a neighbour's proposal; %.0f warp-steps with an updatable pixel, "
      "%.2f %% of them without any adopting neighbour" % (rows, cols, st[2], 100 * st[3] / st[2], st[0], 100 * st[1] / st[0]))
