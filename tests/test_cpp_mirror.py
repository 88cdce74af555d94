"""The C++ host-side mirror (include/pixflow_b200.hpp): compiles and links here (CPU); on the GPU box the reference's
own call sequence (CPU/main.cpp:82-89) runs through it and is checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_exe(tmp_path):
    from panorama_opticalflow_b200 import _lib
    _lib.load()
    exe = str(tmp_path / "novel_view_main")
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "novel_view_main.cpp"), "-o", exe,
           "-L", libdir, "-lpixflow_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def test_cpp_mirror_compiles_links_and_reports_errors(tmp_path):
    exe = _build_exe(tmp_path)
    # unknown algorithm name -> VrCamException, exit code 3 (reference: throw VrCamException, CPU/PixFlow.hpp:499);
    # without a GPU the error is "no CUDA device" instead -- either way a VrCamException, never a silent fallback
    for n in ("L", "R"):
        np.zeros((8, 8, 4), np.uint8).tofile(str(tmp_path / n))
    np.zeros((8, 8), np.float32).tofile(str(tmp_path / "B"))
    r = subprocess.run([exe, "pixflow_bogus", "8", "8", str(tmp_path / "L"), str(tmp_path / "R"), str(tmp_path / "B"),
                        str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 3 and "VrCamException" in r.stderr
    assert "unrecognized flow algorithm name: pixflow_bogus" in r.stderr


@pytest.mark.gpu
def test_reference_call_sequence_in_cpp(orc, tmp_path):
    from panorama_opticalflow_b200 import synth
    exe = _build_exe(tmp_path)
    rows, cols = 96, 150
    L, R = synth.make_pair(rows, cols, 60, 12.0, True)
    blend = synth.make_blend(rows, cols)
    L.tofile(str(tmp_path / "L")); R.tofile(str(tmp_path / "R")); blend.tofile(str(tmp_path / "B"))
    out = str(tmp_path / "out")
    subprocess.check_call([exe, "pixflow_search_20", str(rows), str(cols), str(tmp_path / "L"), str(tmp_path / "R"),
                           str(tmp_path / "B"), out])
    fLR = np.fromfile(out + ".flowLR", np.float32).reshape(rows, cols, 2)
    fRL = np.fromfile(out + ".flowRL", np.float32).reshape(rows, cols, 2)
    merged = np.fromfile(out + ".merged", np.uint8).reshape(rows, cols, 4)
    flow = np.fromfile(out + ".flow", np.float32).reshape(rows, cols, 2)
    wLR, wRL = orc.prepare_bidirectional(L, R, 20)
    assert_bit_equal(fLR, wLR, "C++ getFlowLtoR")
    assert_bit_equal(fRL, wRL, "C++ getFlowRtoL")
    want = orc.combine_novel_views(L, R, wLR, wRL, blend)
    assert np.abs(merged.astype(int) - want.astype(int)).max() <= 1
    assert_bit_equal(flow, orc.compute_flow(L, R, 20, orc.HINT_LEFT), "C++ computeOpticalFlow")
