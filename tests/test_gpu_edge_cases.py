"""Edge cases of the flow path against the oracle: smallest pyramids, degenerate images (all-zero -> 0/0 intensity ratio,
NaN patch errors that must never win), fully transparent images, extreme aspect ratios, and argument validation."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(52, 52), (50, 64), (64, 50), (51, 53)])
def test_smallest_images_single_level_pyramid(orc, engine_search, shape):
    """half-resolution side <= 27 -> the pyramid has exactly one level (CPU/PixFlow.hpp:143)"""
    from panorama_opticalflow_b200 import synth
    L, R = synth.make_pair(shape[0], shape[1], 80, 2.0, False)
    for hint in (0, 1, 3):
        assert_bit_equal(engine_search.computeOpticalFlow(L, R, hint), orc.compute_flow(L, R, 20, hint), "tiny %s hint %d" % (shape, hint))


def test_all_zero_images_nan_intensity_ratio(orc, engine_search):
    """I0 = I1 = 0 with alpha 255: computeIntensityRatio is 0/0 = NaN, every patch error is NaN and must never win
    (CPU/PixFlow.hpp:204, :257); the gradient-constancy error is 0 everywhere."""
    L = np.zeros((90, 120, 4), np.uint8)
    L[..., 3] = 255
    want = orc.compute_flow(L, L, 20, orc.HINT_LEFT)
    got = engine_search.computeOpticalFlow(L, L, orc.HINT_LEFT)
    assert_bit_equal(got, want, "all-zero images")
    assert np.isfinite(want).all()


def test_fully_transparent_images(orc, engine_search):
    from panorama_opticalflow_b200 import synth
    L, R = synth.make_pair(100, 130, 81, 5.0, False)
    L[..., 3] = 0
    R[..., 3] = 0
    a = engine_search.prepareBidirectional(L, R)
    w = orc.prepare_bidirectional(L, R, 20)
    assert_bit_equal(a[0], w[0], "alpha=0 flowLtoR")
    assert_bit_equal(a[1], w[1], "alpha=0 flowRtoL")
    assert not a[0].any() and not a[1].any()       # nothing is ever updated: the flow stays exactly zero


def test_one_sided_alpha(orc, engine_search):
    """alpha0 > 0.9 but alpha1 = 0: adjustInitialFlow runs (tests alpha0 only, :241) with sad/0 = inf or NaN patch
    errors, the sweeps never update (both alphas are required, :317), the diffusion replaces the flow by its blur."""
    from panorama_opticalflow_b200 import synth
    L, R = synth.make_pair(120, 150, 82, 20.0, False)
    R[..., 3] = 0
    assert_bit_equal(engine_search.computeOpticalFlow(L, R, orc.HINT_LEFT), orc.compute_flow(L, R, 20, orc.HINT_LEFT), "alpha1=0")


@pytest.mark.parametrize("shape", [(60, 900), (900, 60)])
def test_extreme_aspect_ratios(orc, engine_low, shape):
    from panorama_opticalflow_b200 import synth
    L, R = synth.make_pair(shape[0], shape[1], 83, 7.0, True)
    a = engine_low.prepareBidirectional(L, R)
    w = orc.prepare_bidirectional(L, R, 0)
    assert_bit_equal(a[0], w[0], "aspect %s LR" % (shape,))
    assert_bit_equal(a[1], w[1], "aspect %s RL" % (shape,))


def test_saturated_and_noise_images(orc, engine_search):
    """white images (zero gradients: sqrt(0) and 0/x paths of the branch-free exact math) and pure noise."""
    W = np.full((80, 100, 4), 255, np.uint8)
    assert_bit_equal(engine_search.computeOpticalFlow(W, W, 3), orc.compute_flow(W, W, 20, 3), "white")
    rng = np.random.default_rng(9)
    N0 = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)
    N1 = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)
    assert_bit_equal(engine_search.computeOpticalFlow(N0, N1, 1), orc.compute_flow(N0, N1, 20, 1), "noise")


def test_argument_validation(engine_low):
    from panorama_opticalflow_b200 import _lib
    lib = _lib.load()
    img = np.zeros((64, 64, 4), np.uint8)
    flow = np.zeros((64, 64, 2), np.float32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    h = engine_low._h
    assert lib.pf_compute_flow(h, p(img), 256, p(img), 256, 64, 64, 7, p(flow), 512) == _lib.PF_ERR_INVALID_ARGUMENT   # bad hint
    assert b"unexpected direction" in lib.pf_last_error()
    assert lib.pf_compute_flow(h, p(img), 100, p(img), 256, 64, 64, 0, p(flow), 512) == _lib.PF_ERR_INVALID_ARGUMENT   # stride < cols*4
    assert lib.pf_compute_flow(h, p(img), 256, p(img), 256, 6, 6, 0, p(flow), 512) == _lib.PF_ERR_INVALID_ARGUMENT     # too small
    assert lib.pf_compute_flow(h, None, 256, p(img), 256, 64, 64, 0, p(flow), 512) == _lib.PF_ERR_INVALID_ARGUMENT
    assert lib.pf_compute_flow(h, p(img), 256, p(img), 256, 0, 64, 0, p(flow), 512) == _lib.PF_ERR_INVALID_ARGUMENT
    # the engine is still usable afterwards
    assert lib.pf_compute_flow(h, p(img), 256, p(img), 256, 64, 64, 0, p(flow), 512) == _lib.PF_OK


def test_engines_in_concurrent_threads(orc):
    """one engine per thread: engines are created, used (graph capture on first use) and destroyed
    concurrently; every call must give the oracle's bits"""
    import threading
    import panorama_opticalflow_b200 as pf
    from panorama_opticalflow_b200 import synth
    n = 3
    pairs = [synth.make_pair(100 + 8 * i, 140, 90 + i, 9.0, bool(i % 2)) for i in range(n)]
    want = [orc.prepare_bidirectional(L, R, 20) for L, R in pairs]
    errors = []

    def work(i):
        try:
            for _ in range(3):                      # create / capture / replay / destroy, three times over
                e = pf.makeOpticalFlowByName("pixflow_search_20")
                for _ in range(2):
                    got = e.prepareBidirectional(*pairs[i])
                    assert np.array_equal(got[0], want[i][0]) and np.array_equal(got[1], want[i][1])
                e.close()
        except Exception as ex:                     # noqa: BLE001
            errors.append((i, repr(ex)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
