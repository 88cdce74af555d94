"""The C++ host-side mirror (include/pixflow_b200.hpp): compiles and links here (CPU); on the GPU box the reference's
own call sequence (CPU/main.cpp:82-89) runs through it and is checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import assert_bit_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_exe(tmp_path, name="novel_view_main"):
    from panorama_opticalflow_b200 import _lib
    _lib.load()
    exe = str(tmp_path / name)
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
           "-L", libdir, "-lpixflow_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("name", ["novel_view_main", "stitch_main"])
def test_cv_mat_branch_type_checks_against_a_stub_opencv(tmp_path, name):
    """OpenCV C++ is not installed in this image, so the `#ifdef PIXFLOW_B200_HAVE_OPENCV` branch of pixflow_b200.hpp (Mat =
    cv::Mat, the reference's own matrix type at the boundary: CPU/OpticalFlow.hpp:34-70) would never be compiled.  A stub
    <opencv2/core.hpp> with cv::Mat's real member signatures (tests/cpp/stub) makes the compiler check that branch, and the
    reference-style drivers link against the library through it."""
    from panorama_opticalflow_b200 import _lib
    _lib.load()
    libdir = os.path.dirname(_lib.lib_path())
    exe = str(tmp_path / (name + "_cv"))
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "pixflow_b200.hpp"\n#ifndef PIXFLOW_B200_HAVE_OPENCV\n#error cv::Mat branch not selected\n#endif\n'
                     'static_assert(std::is_same<pf::Mat, cv::Mat>::value, "pf::Mat must be cv::Mat");\n'
                     'static_assert(pf::PF_8UC4 == CV_8UC4 && pf::PF_32FC2 == CV_32FC2 && pf::PF_32FC1 == CV_32FC1 && pf::PF_8UC1 == CV_8UC1, "type codes");\n'
                     'int main() { return 0; }\n')
    inc = ["-I", os.path.join(ROOT, "tests", "cpp", "stub"), "-I", os.path.join(ROOT, "include")]
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only"] + inc + ["-include", "type_traits", str(probe)])
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall"] + inc + [os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
                           "-L", libdir, "-lpixflow_b200", "-Wl,-rpath," + libdir])


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


def test_cpp_stitch_mirror_compiles_and_links(tmp_path):
    exe = _build_exe(tmp_path, "stitch_main")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_reference_driver_loop_body_in_cpp(orc, tmp_path):
    """CPU/main.cpp:72-95 compiled against the C++ mirror, checked against the oracle's composition of the same steps."""
    from panorama_opticalflow_b200 import synth
    exe = _build_exe(tmp_path, "stitch_main")
    rows, cols = 400, 320
    L, R = synth.make_pair(rows, cols, seed=9, amplitude=16.0, sparse=False)
    x = np.mgrid[0:rows, 0:cols][1]
    L[..., 3] = np.where(x < 0.7 * cols, 255, 0)
    R[..., 3] = np.where(x > 0.3 * cols, 255, 0)
    L[L[..., 3] == 0] = 0
    R[R[..., 3] == 0] = 0
    L.tofile(str(tmp_path / "L")); R.tofile(str(tmp_path / "R"))
    out = str(tmp_path / "out")
    subprocess.check_call([exe, "pixflow_search_20", str(rows), str(cols), str(tmp_path / "L"), str(tmp_path / "R"), out])
    want, inter = orc.stitch_iteration(L, R, 20)
    assert_bit_equal(np.fromfile(out + ".map", np.uint8).reshape(rows, cols), inter["map"], "C++ getMap")
    assert_bit_equal(np.fromfile(out + ".blend", np.float32).reshape(rows, cols), inter["blend"], "C++ getBlend")
    final = np.fromfile(out + ".final", np.uint8).reshape(rows, cols, 4)
    d = np.abs(final.astype(int) - want.astype(int))
    assert d[..., :3].max() <= 1 and d[..., 3].max() == 0
    assert_bit_equal(np.fromfile(out + ".fused", np.uint8).reshape(rows, cols, 4), final, "C++ stitchIteration vs step by step")
