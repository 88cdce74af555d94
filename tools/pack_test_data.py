#!/usr/bin/env python
"""Packs the reference's test inputs (BASELINE configs 3 and 5) into a compact LOSSLESS form under data/ (git-ignored, but it
travels to the GPU box with the repo snapshot; /root/reference does not exist there).

    python tools/pack_test_data.py            # Test_data/1 (top + 1..5) and Test_data_4Input (1..4)

Every input is an RGBA TIFF of the full canvas whose alpha is {0, 255} with 26-40 % coverage and whose colour is zero wherever
alpha is zero (asserted), so only the bounding box of the non-zero alpha is stored, as a 4-channel PNG, plus its offset and the
canvas size in index.json.  panorama_opticalflow_b200.testdata.load() rebuilds the exact arrays (sha256 of the full-size
array is recorded and checked).  The shipped FinalResult.png files are copied as they are (known answers for PSNR / alpha).
"""
import hashlib
import json
import os
import shutil
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SETS = {"Test_data_1": ("Test_data/1", ["top", "1", "2", "3", "4", "5"]),
        "Test_data_4Input": ("Test_data_4Input", ["1", "2", "3", "4"])}


def main():
    want_final = "--no-final" not in sys.argv
    for out_name, (src_dir, names) in SETS.items():
        out = os.path.join(ROOT, "data", out_name)
        os.makedirs(out, exist_ok=True)
        index = {"source": src_dir, "images": {}}
        for n in names:
            im = cv2.imread(os.path.join(REF, src_dir, n + ".tif"), cv2.IMREAD_UNCHANGED)
            assert im is not None and im.ndim == 3 and im.shape[2] == 4 and im.dtype == np.uint8, n
            a = im[..., 3]
            assert not np.count_nonzero(im[a == 0]), "colour under zero alpha would be lost by the crop"
            ys, xs = np.nonzero(a)
            y0, y1, x0, x1 = int(ys.min()), int(ys.max()) + 1, int(xs.min()), int(xs.max()) + 1
            crop = np.ascontiguousarray(im[y0:y1, x0:x1])
            assert cv2.imwrite(os.path.join(out, n + ".png"), crop, [cv2.IMWRITE_PNG_COMPRESSION, 6])
            index["images"][n] = {"rows": im.shape[0], "cols": im.shape[1], "y0": y0, "x0": x0,
                                  "sha256": hashlib.sha256(im.tobytes()).hexdigest()}
            print(out_name, n, im.shape, "->", os.path.getsize(os.path.join(out, n + ".png")) // 1000, "kB", flush=True)
        if want_final:
            shutil.copyfile(os.path.join(REF, src_dir, "FinalResult.png"), os.path.join(out, "FinalResult.png"))
            index["final_result"] = "FinalResult.png"
        with open(os.path.join(out, "index.json"), "w") as f:
            json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
