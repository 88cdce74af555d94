"""Oracle pins for the rest of the stitching step: the blend smoothing of GenerateBlend
(CPU/StitchTool.cpp:133-145) against cv2.blur, and Stitchtools::Gather (:52-96) against a direct numpy restatement."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from test_stitch_prepare import _canvas_pair

cv2 = pytest.importorskip("cv2")
cv2.setUseOptimized(False)


def _adversarial(shape, seed):
    """values whose double-precision running sums are inexact (2^-35-sized values next to 1.0): any change in the order of
    the box filter's additions shows up in the fp32 result"""
    rng = np.random.default_rng(seed)
    img = (rng.random(shape) * 2.0 ** -35).astype(np.float32)
    img[rng.random(shape) < 0.02] = 1.0
    return img


@pytest.mark.parametrize("shape", [(67, 93), (40, 31)])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 10, 15, 30])
def test_box_blur_matches_cv2_blur_bit_for_bit(orc, shape, k):
    for seed, img in ((0, _adversarial(shape, k)), (1, np.random.default_rng(k).random(shape).astype(np.float32))):
        got = orc.box_blur_rect(img, 0, 0, shape[1], shape[0], k)
        assert_bit_equal(got, cv2.blur(img, (k, k)), "blur k=%d data %d" % (k, seed))


@pytest.mark.parametrize("k", [3, 6, 30])
def test_box_blur_roi_reads_the_parent(orc, k):
    """OpenCV ROI semantics: window pixels outside the rectangle come from the parent image.  On data whose sums are exact
    (multiples of 2^-10) the result is independent of the summation order, so it must equal the same rectangle of the
    whole-image blur."""
    rng = np.random.default_rng(5)
    parent = (rng.integers(0, 1025, (90, 120)) / 1024.0).astype(np.float32)
    whole = cv2.blur(parent, (k, k))
    for (x0, y0, rw, rh) in [(0, 0, 10, 10), (50, 40, 10, 10), (110, 80, 10, 10), (3, 77, 7, 5), (100, 0, 20, 9)]:
        got = orc.box_blur_rect(parent, x0, y0, rw, rh, k)
        assert_bit_equal(got, whole[y0:y0 + rh, x0:x0 + rw], "roi %r" % ((x0, y0, rw, rh),))


def _smooth_with_cv2(blend, md, rows, cols):
    """GenerateBlend :133-145 with cv2 calls.  Each block is blurred inside a crop of the current image that either ends at
    the image edge (same reflect-101 as the parent) or extends 2k past the block (further than the window reaches)."""
    b = blend.copy()
    step = cols // 200 if cols <= rows else rows // 200
    k1, k2 = rows // 130, rows // 400
    for y in range(0, rows - step, step):
        if y + step >= rows:
            break
        for x in range(0, cols, step):
            if x + step >= cols:
                break
            if md[y, x] > step:
                ya, yb = max(0, y - 2 * k1), min(rows, y + step + 2 * k1)
                xa, xb = max(0, x - 2 * k1), min(cols, x + step + 2 * k1)
                crop = cv2.blur(b[ya:yb, xa:xb].copy(), (k1, k1))
                b[y:y + step, x:x + step] = crop[y - ya:y - ya + step, x - xa:x - xa + step]
    return cv2.blur(b, (k2, k2))


@pytest.mark.parametrize("rows,cols,wrap", [(400, 250, True), (810, 420, False), (520, 1000, True)])
def test_blend_smoothing_matches_cv2_composition(orc, rows, cols, wrap):
    L, R = _canvas_pair(rows, cols, 3, wrap)
    m, _, _ = orc.stitch_match_and_mask(L, R)
    braw, md = orc.stitch_blend_raw(m)
    step = cols // 200 if cols <= rows else rows // 200
    assert (md[::step, ::step] > step).any(), "no block is smoothed: the test would be vacuous"
    got = orc.stitch_blend_smooth(braw, md)
    assert_bit_equal(got, _smooth_with_cv2(braw, md, rows, cols), "Blend")


def test_blend_smoothing_rejects_what_the_reference_cannot_run(orc):
    b = np.zeros((390, 300), np.float32)
    with pytest.raises(ValueError):
        orc.stitch_blend_smooth(b, b)


def _gather_numpy(L, R, M, map0):
    rows, cols = map0.shape
    mp = np.minimum(map0.astype(int) + np.where(M[..., 3] > 0, 75, 0), 255).astype(np.uint8)
    flat = mp.ravel()
    n = flat.size
    out = np.zeros_like(L)
    out[mp == 100] = L[mp == 100]
    out[mp == 50] = R[mp == 50]
    mm = (mp == 225) | (mp == 125) | (mp == 175)
    out[mm] = M[mm]
    for y, x in np.argwhere(mp == 150):
        for i in range(1, 100):
            idx = [y * cols + x + i, y * cols + x - i, (y + i) * cols + x, (y - i) * cols + x, (y - i) * cols + x - i,
                   (y - i) * cols + x + i, (y + i) * cols + x - i, (y + i) * cols + x + i]
            v = [int(flat[j]) if 0 <= j < n else 0 for j in idx]
            if 100 in v:
                out[y, x] = L[y, x]
                break
            if 50 in v:
                out[y, x] = R[y, x]
                break
            out[y, x] = (0, 0, 0, 255)
    return out


@pytest.mark.parametrize("rows,cols,wrap", [(240, 330, False), (400, 250, True), (420, 900, False)])
def test_gather_against_numpy(orc, rows, cols, wrap):
    L, R = _canvas_pair(rows, cols, 9, wrap)
    if cols == 900:     # a 300 px wide overlap band: its centre is further than 99 px from both images
        L[..., 3] = 0
        L[:, :600, 3] = 255
        R[..., 3] = 0
        R[:, 300:, 3] = 255
        L[:40, :200, 3] = 0
    m, _, _ = orc.stitch_match_and_mask(L, R)
    rng = np.random.default_rng(2)
    M = rng.integers(1, 256, (rows, cols, 4), dtype=np.uint8)
    # the merged view covers the overlap except for a few holes (alpha 0), some of them far from any L-only / R-only pixel
    M[..., 3] = np.where(m == 150, 255, 0)
    yy, xx = np.mgrid[0:rows, 0:cols]
    holes = ((yy // 7 + xx // 5) % 11 == 0) | ((np.abs(yy - rows // 2) < 6) & (np.abs(xx - cols // 2) < 130))
    M[..., 3] = np.where(holes, 0, M[..., 3])
    zy, zx = np.nonzero(m == 0)
    M[zy[:40], zx[:40], 3] = 255      # merged alpha outside both images: code 75, result stays 0
    got = orc.stitch_gather(L, R, M, m)
    want = _gather_numpy(L, R, M, m)
    assert_bit_equal(got, want, "FinalResult")
    mp = m.astype(int) + np.where(M[..., 3] > 0, 75, 0)
    assert (mp == 150).sum() > 100 and (mp == 75).sum() == min(40, zy.size) > 0
    if cols == 900:
        assert ((got == (0, 0, 0, 255)).all(-1) & (mp == 150)).any(), "no hole further than 99 px from both images"


def _four_inputs(rows, cols, seed):
    rng = np.random.default_rng(seed)
    imgs = []
    for k in range(4):
        a = rng.integers(0, 256, (rows, cols, 4), dtype=np.uint8)
        x = np.mgrid[0:rows, 0:cols][1]
        a[..., 3] = np.where((x + 37 * k) % 90 < 55, 255, 0)
        a[rows // 2, rng.integers(0, cols, 12), 3] = 0     # columns blanked only because of the middle row
        imgs.append(a)
    return imgs


def test_four_input_frontend_against_numpy(orc):
    rows, cols = 61, 200
    imgs = _four_inputs(rows, cols, 4)
    L, R = orc.four_input_frontend(*imgs)
    kept = [np.where((a[rows // 2, :, 3] != 0)[None, :, None], a, 0).astype(int) for a in imgs]
    assert np.array_equal(L, np.minimum(kept[0] + kept[2], 255).astype(np.uint8))
    assert np.array_equal(R, np.minimum(kept[1] + kept[3], 255).astype(np.uint8))
    assert (L == 255).any() and (kept[0] + kept[2] > 255).any()
