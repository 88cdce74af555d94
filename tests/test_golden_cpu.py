"""The CPU oracle must reproduce the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PCT = {"pixflow_low": 0, "pixflow_search_20": 20}


def load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLD, name + ".npz"))


def test_config1_golden(orc):
    meta, g = load("config1_512_low")
    trace = {}
    flow = orc.compute_flow(g["L"], g["R"], 0, orc.HINT_LEFT, trace)
    assert_bit_equal(flow, g["flow"], "config1 flow")
    for k, v in trace.items():
        assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == meta["stages"]["%d/%s" % k], k


@pytest.mark.parametrize("name", ["search_odd_prepare", "sparse_prepare"])
def test_prepare_goldens(orc, name):
    from panorama_opticalflow_b200 import synth
    meta, g = load(name)
    fLR, fRL = orc.prepare_bidirectional(g["L"], g["R"], PCT[meta["preset"]])
    assert_bit_equal(fLR, g["flowLR"], name + " flowLtoR")
    assert_bit_equal(fRL, g["flowRL"], name + " flowRtoL")
    blend = synth.make_blend(meta["rows"], meta["cols"])
    assert_bit_equal(orc.combine_novel_views(g["L"], g["R"], fLR, fRL, blend), g["merged"], name + " merged")


def test_synth_generator_is_reproducible():
    from panorama_opticalflow_b200 import synth
    meta, g = load("search_odd_prepare")
    L, R = synth.make_pair(meta["rows"], meta["cols"], meta["seed"], meta["amplitude"], meta["sparse"])
    # the generator is numpy-only; allow 1 LSB in case libm's sin/cos differ across hosts
    assert np.abs(L.astype(int) - g["L"].astype(int)).max() <= 1
    assert np.abs(R.astype(int) - g["R"].astype(int)).max() <= 1
