"""Loader of the packed copies of the reference's test inputs (data/, written by tools/pack_test_data.py): BASELINE configs 3
(Test_data/1: top + 1..5, 4000 x 8998) and 5 (Test_data_4Input: 1..4, 3585 x 7352).  Harness code for bench.py and the tests."""
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")


def available(set_name):
    return os.path.exists(os.path.join(DATA, set_name, "index.json"))


def load(set_name, name, verify=True, out=None):
    """-> the full-canvas BGRA uint8 array of input `name` ("top", "1", ...), exactly as cv2.imread(<name>.tif, -1) gives it"""
    import cv2
    d = os.path.join(DATA, set_name)
    with open(os.path.join(d, "index.json")) as f:
        meta = json.load(f)["images"][name]
    crop = cv2.imread(os.path.join(d, name + ".png"), cv2.IMREAD_UNCHANGED)
    if crop is None or crop.ndim != 3 or crop.shape[2] != 4:
        raise IOError("bad packed image %s/%s" % (set_name, name))
    full = np.zeros((meta["rows"], meta["cols"], 4), np.uint8) if out is None else out
    if out is not None:
        full[...] = 0
    full[meta["y0"]:meta["y0"] + crop.shape[0], meta["x0"]:meta["x0"] + crop.shape[1]] = crop
    if verify and hashlib.sha256(full.tobytes()).hexdigest() != meta["sha256"]:
        raise IOError("packed image %s/%s does not reproduce the reference TIFF" % (set_name, name))
    return full


def final_result(set_name):
    """the reference's shipped FinalResult.png (BGRA), or None when it was not packed"""
    import cv2
    p = os.path.join(DATA, set_name, "FinalResult.png")
    if not os.path.exists(p):
        return None
    return cv2.imread(p, cv2.IMREAD_UNCHANGED)
