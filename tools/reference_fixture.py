#!/usr/bin/env python
"""Pins the CPU oracle to the only artefacts the reference itself holds for this path: the shipped FinalResult.png files.

    python tools/reference_fixture.py [--set 1|2|4input|all] [--alg pixflow_low|pixflow_search_20] [--out tests/golden/reference_fixture.json]

Runs the oracle's restatement of the reference drivers on the reference's own inputs (read from /root/reference, which only
exists in the build container):
  * CPU/main.cpp:55-105     top.tif + 1..5.tif, five sequential Stitchtools::prepare -> flow -> blend -> Gather iterations
  * CPU_4Input/main.cpp:54-113  the four-input single pass (with the 0.95 row crop of :82-83 enabled, which is how the shipped
    3405-row FinalResult.png was produced)
and compares the result with the shipped PNG: alpha equality, PSNR, share of bit-equal / within-1-LSB pixels.  The provenance
of the PNGs (CPU or GPU build, preset, OpenCV version) is not recorded by the reference, and the flow iteration amplifies
rounding differences, so they are a structural known answer, not a bit-exact golden.  The numbers
are stored in tests/golden/reference_fixture.json; tests/test_reference_fixture.py re-derives and asserts them.
TEST INFRASTRUCTURE (uses oracle/), not product code.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def imread_bgra(path):
    """imreadExceptionOnFail(path, -1) + the CV_BGR2BGRA promotion of CPU/main.cpp:57-58"""
    import cv2
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise IOError("cannot read " + path)
    if im.ndim == 3 and im.shape[2] == 3:
        im = cv2.cvtColor(im, cv2.COLOR_BGR2BGRA)
    return np.ascontiguousarray(im)


def compare(result, shipped):
    """alpha equality + PSNR / equality statistics over the B, G, R channels"""
    assert result.shape == shipped.shape, (result.shape, shipped.shape)
    d = np.abs(result[..., :3].astype(np.int16) - shipped[..., :3].astype(np.int16))
    mse = float(np.mean(d.astype(np.float64) ** 2))
    px_eq = np.all(d == 0, axis=2)
    px_1 = np.all(d <= 1, axis=2)
    return {
        "shape": list(result.shape),
        "alpha_identical": bool(np.array_equal(result[..., 3], shipped[..., 3])),
        "alpha_mismatch_px": int(np.count_nonzero(result[..., 3] != shipped[..., 3])),
        "psnr_db": float(10 * np.log10(255.0 ** 2 / mse)) if mse > 0 else float("inf"),
        "pixels_bit_equal": float(px_eq.mean()),
        "pixels_within_1lsb": float(px_1.mean()),
        "max_abs_diff": int(d.max()),
    }


def run_set5(orc, name, alg_pct, keep=None, threads=2):
    """CPU/main.cpp:55-105"""
    d = os.path.join(REF, "Test_data", name)
    final = imread_bgra(os.path.join(d, "top.tif"))
    secs = []
    for i in range(1, 6):
        t0 = time.time()
        L = imread_bgra(os.path.join(d, "%d.tif" % i))
        final, inter = orc.stitch_iteration(L, final, alg_pct, threads=threads)     # colorImageR = previous FinalResult (:64-65)
        secs.append(time.time() - t0)
        if keep is not None:
            keep.append(final.copy())
        print("  %s iteration %d: %.1f s" % (name, i, secs[-1]), flush=True)
    return final, secs


def run_4input(orc, alg_pct, crop=True, threads=2):
    """CPU_4Input/main.cpp:54-113; crop=True enables the commented-out 0.95 row crop of :82-83"""
    d = os.path.join(REF, "Test_data_4Input")
    imgs = [imread_bgra(os.path.join(d, "%d.tif" % k)) for k in range(1, 5)]
    t0 = time.time()
    L, R = orc.four_input_frontend(*imgs)
    if crop:
        n = int(0.95 * L.shape[0])                                   # Range(0, 0.95*rows): double -> int truncation
        L, R = np.ascontiguousarray(L[:n]), np.ascontiguousarray(R[:n])
    final, inter = orc.stitch_iteration(L, R, alg_pct, threads=threads)
    return final, [time.time() - t0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="all")
    ap.add_argument("--alg", default="pixflow_low")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_fixture.json"))
    ap.add_argument("--save-dir", default=None, help="also write the oracle's FinalResult as PNG here (scratch)")
    args = ap.parse_args()
    import cv2
    from oracle import orc
    orc.build()
    pct = {"pixflow_low": 0, "pixflow_search_20": 20}[args.alg]
    results = {}
    if os.path.exists(args.out):
        with open(args.out) as f:
            results = json.load(f)
    sets = ["1", "2", "4input"] if args.set == "all" else [args.set]
    for s in sets:
        print("set", s, args.alg, flush=True)
        if s == "4input":
            final, secs = run_4input(orc, pct)
            shipped = imread_bgra(os.path.join(REF, "Test_data_4Input", "FinalResult.png"))
            key = "Test_data_4Input"
        else:
            final, secs = run_set5(orc, s, pct)
            shipped = imread_bgra(os.path.join(REF, "Test_data", s, "FinalResult.png"))
            key = "Test_data/" + s
        r = compare(final, shipped)
        r["oracle_seconds"] = [round(x, 1) for x in secs]
        r["flow_alg"] = args.alg
        print(json.dumps(r), flush=True)
        results.setdefault(key, {})[args.alg] = r
        if args.save_dir:
            os.makedirs(args.save_dir, exist_ok=True)
            cv2.imwrite(os.path.join(args.save_dir, "oracle_%s_%s.png" % (s, args.alg)), final)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
