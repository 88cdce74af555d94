"""Pins the pure-C oracle against the real OpenCV arithmetic (cv2 4.13.0, scalar mode), primitive by primitive
and for the whole pipeline.  OpenCV is the un-vendored third-party dependency the reference calls
(README.md:34), so this is the strongest pin available in an image where the reference cannot be compiled."""
import numpy as np
import pytest

from conftest import assert_bit_equal

cv2 = pytest.importorskip("cv2")
cv2.setUseOptimized(False)

from panorama_opticalflow_b200 import synth  # noqa: E402

RNG = np.random.default_rng(1234)


@pytest.mark.parametrize("shape", [(33, 47), (26, 25), (64, 101)])
@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("ks", [(5, 0.25), (3, 0.5), (3, 1.0), (15, 8.0)])
def test_gaussian_blur(orc, shape, ch, ks):
    a = RNG.standard_normal(shape if ch == 1 else shape + (ch,)).astype(np.float32)
    assert_bit_equal(orc.gaussian_blur(a, *ks), cv2.GaussianBlur(a, (ks[0], ks[0]), ks[1]), "GaussianBlur")


@pytest.mark.parametrize("ks", [(5, 0.25), (3, 0.5), (3, 1.0), (15, 8.0)])
def test_gaussian_kernel(orc, ks):
    assert_bit_equal(orc.gaussian_kernel(*ks), cv2.getGaussianKernel(ks[0], ks[1], cv2.CV_32F).ravel(), "kernel")


def test_sobel_median_gray(orc):
    a = RNG.random((40, 51)).astype(np.float32)
    kw = dict(ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    assert_bit_equal(orc.sobel(a, 1), cv2.Sobel(a, -1, 1, 0, **kw), "Sobel x")
    assert_bit_equal(orc.sobel(a, 0), cv2.Sobel(a, -1, 0, 1, **kw), "Sobel y")
    f = RNG.standard_normal((37, 45, 2)).astype(np.float32)
    assert_bit_equal(orc.median5_c2(f), cv2.medianBlur(f, 5), "medianBlur")
    u = RNG.integers(0, 256, (61, 83, 4), dtype=np.uint8)
    assert_bit_equal(orc.bgra2gray(u), cv2.cvtColor(u, cv2.COLOR_BGRA2GRAY), "cvtColor")


@pytest.mark.parametrize("shape", [(50, 60), (101, 77), (45, 25)])
@pytest.mark.parametrize("ch", [1, 2])
def test_resize_linear(orc, shape, ch):
    sh, sw = shape
    a = RNG.standard_normal(shape if ch == 1 else shape + (ch,)).astype(np.float32)
    dw = int(np.float32(sw) * np.float32(0.9) + np.float32(0.5))
    dh = int(np.float32(sh) * np.float32(0.9) + np.float32(0.5))
    for (oh, ow) in [(dh, dw), (2 * sh, 2 * sw), (2 * sh + 1, 2 * sw + 1)]:
        assert_bit_equal(orc.resize_linear(a, oh, ow), cv2.resize(a, (ow, oh), interpolation=cv2.INTER_LINEAR),
                         "INTER_LINEAR %s->%s" % (shape, (oh, ow)))


@pytest.mark.parametrize("shape", [(50, 60), (101, 77), (45, 25)])
def test_resize_cubic_f32(orc, shape):
    sh, sw = shape
    a = RNG.standard_normal(shape + (2,)).astype(np.float32)
    for (oh, ow) in [(int(sh / 0.9), int(sw / 0.9)), (int(sh / 0.9) + 1, int(sw / 0.9) + 1), (56, 28), (55, 29)]:
        assert_bit_equal(orc.resize_cubic_f32(a, oh, ow), cv2.resize(a, (ow, oh), interpolation=cv2.INTER_CUBIC),
                         "INTER_CUBIC f32 %s->%s" % (shape, (oh, ow)))


@pytest.mark.parametrize("shape", [(50, 60), (101, 77), (45, 25), (64, 66), (200, 202)])
def test_resize_cubic_u8(orc, shape):
    u = RNG.integers(0, 256, shape + (4,), dtype=np.uint8)
    dh, dw = orc.downscale_size(*shape)
    assert_bit_equal(orc.resize_cubic_u8c4(u, dh, dw), cv2.resize(u, (dw, dh), interpolation=cv2.INTER_CUBIC), "INTER_CUBIC u8")


@pytest.mark.parametrize("case", [(128, 160, 0, 6, False, 0, 3), (150, 131, 1, 14, False, 20, 3), (133, 158, 2, 14, True, 20, 1)])
def test_pipeline_matches_cv2_composition(orc, case):
    from oracle import cv2_oracle
    rows, cols, seed, amp, sparse, pct, hint = case
    L, R = synth.make_pair(rows, cols, seed, amp, sparse)
    assert_bit_equal(orc.compute_flow(L, R, pct, hint), cv2_oracle.compute_flow(L, R, pct, hint), "computeOpticalFlow")


def test_prepare_matches_cv2_composition(orc):
    from oracle import cv2_oracle
    L, R = synth.make_pair(120, 200, 3, 12, True)
    a = orc.prepare_bidirectional(L, R, 20)
    b = cv2_oracle.prepare_bidirectional(L, R, 20)
    assert_bit_equal(a[0], b[0], "flowLtoR")
    assert_bit_equal(a[1], b[1], "flowRtoL")
