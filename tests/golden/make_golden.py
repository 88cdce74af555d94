"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

Every fixture is produced by the pure-C oracle (oracle/pixflow_oracle.c) and, at generation time, cross-checked
bit-for-bit against the cv2 4.13.0 scalar-mode composition (oracle/cv2_oracle.py).  The reference itself cannot be
built in this image, so these are the frozen definition of "reference CPU path" results.
Fixtures hold the inputs (uint8), the final flows, the merged novel view, and a sha256 of every per-level
intermediate so that a divergence can be localised to (level, stage).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_oracle, orc  # noqa: E402
from panorama_opticalflow_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (rows, cols, seed, amplitude, sparse, preset, mode)
CASES = {
    # BASELINE.json configs[0]: pixflow_low on one 512x512 pair, direct computeOpticalFlow(L, R, LEFT)
    "config1_512_low": (512, 512, 0, 6.0, False, "pixflow_low", "flow"),
    # search branch live (disparity >= cols/12), odd sizes (SIMD-tail rules), through prepare()
    "search_odd_prepare": (181, 243, 1, 24.0, False, "pixflow_search_20", "prepare"),
    # sparse alpha: alpha tests, NaN/inf path of computePatchError, diffusion
    "sparse_prepare": (160, 220, 2, 20.0, True, "pixflow_search_20", "prepare"),
}
PCT = {"pixflow_low": 0, "pixflow_search_20": 20}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name, (rows, cols, seed, amp, sparse, preset, mode) in CASES.items():
        L, R = synth.make_pair(rows, cols, seed, amp, sparse)
        pct = PCT[preset]
        out = {"L": L, "R": R}
        meta = {"rows": rows, "cols": cols, "seed": seed, "amplitude": amp, "sparse": sparse, "preset": preset,
                "mode": mode, "oracle": "oracle/pixflow_oracle.c (gcc -O2 -ffp-contract=off), cross-checked vs cv2 4.13.0 scalar"}
        if mode == "flow":
            trace = {}
            flow = orc.compute_flow(L, R, pct, orc.HINT_LEFT, trace)
            ref = cv2_oracle.compute_flow(L, R, pct, orc.HINT_LEFT)
            assert np.array_equal(flow, ref), name
            out["flow"] = flow
            meta["stages"] = {"%d/%s" % k: sha(v) for k, v in sorted(trace.items())}
        else:
            fLR, fRL = orc.prepare_bidirectional(L, R, pct)
            rLR, rRL = cv2_oracle.prepare_bidirectional(L, R, pct)
            assert np.array_equal(fLR, rLR) and np.array_equal(fRL, rRL), name
            blend = synth.make_blend(rows, cols)
            out["flowLR"], out["flowRL"] = fLR, fRL
            out["merged"] = orc.combine_novel_views(L, R, fLR, fRL, blend)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(os.path.join(HERE, name + ".npz")) >> 10, "KiB")


if __name__ == "__main__":
    main()
