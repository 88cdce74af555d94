"""Stage-by-stage parity of every CUDA kernel against the CPU oracle -- bit-exact (fp32 results compared with
==), through the diagnostic pf_stage_* entry points of the C-ABI."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(77)


def _pair_planes(orc, rows, cols, seed, amp, sparse, level=0):
    """I0, I1, A0, A1 at a pyramid level from a synthetic pair, via the oracle front end."""
    from panorama_opticalflow_b200 import synth
    L, R = synth.make_pair(rows, cols, seed, amp, sparse)
    I0, A0 = orc.frontend(L)
    I1, A1 = orc.frontend(R)
    for _ in range(level):
        h, w = I0.shape
        (nw, nh) = orc.pyramid_sizes(w, h)[1]
        I0, I1, A0, A1 = (orc.resize_linear(x, nh, nw) for x in (I0, I1, A0, A1))
    return I0, I1, A0, A1


def _grad(orc, I):
    return np.stack([orc.gaussian_blur(orc.sobel(I, 1), 3, 0.5), orc.gaussian_blur(orc.sobel(I, 0), 3, 0.5)], axis=2)


@pytest.mark.parametrize("shape,pad", [((64, 80), 0), ((101, 77), 0), ((90, 121), 6), ((200, 202), 10), ((67, 64), 3)])
def test_frontend(orc, shape, pad):
    from panorama_opticalflow_b200 import stages
    u = RNG.integers(0, 256, shape + (4,), dtype=np.uint8)
    cols = shape[1]
    padded = np.ascontiguousarray(np.concatenate([u[:, cols - pad:], u, u[:, :pad]], 1)) if pad else u
    dh, dw = orc.downscale_size(padded.shape[0], padded.shape[1])
    small = orc.resize_cubic_u8c4(padded, dh, dw)
    inv = np.float32(1.0 / 255.0)
    want_g = orc.bgra2gray(small).astype(np.float32) * inv
    want_a = small[..., 3].astype(np.float32) * inv
    g, a = stages.frontend(u, pad)
    assert_bit_equal(g, want_g, "grey")
    assert_bit_equal(a, want_a, "alpha")
    assert_bit_equal(stages.gauss5(g), orc.gaussian_blur(want_g, 5, 0.25), "pre-blur")


@pytest.mark.parametrize("shape", [(50, 60), (101, 77), (45, 25), (256, 256)])
def test_pyr_down(orc, shape):
    from panorama_opticalflow_b200 import stages
    a = RNG.random(shape).astype(np.float32)
    nw = int(np.float32(shape[1]) * np.float32(0.9) + np.float32(0.5))
    nh = int(np.float32(shape[0]) * np.float32(0.9) + np.float32(0.5))
    assert_bit_equal(stages.pyr_down(a, nh, nw), orc.resize_linear(a, nh, nw), "pyr_down")


@pytest.mark.parametrize("shape", [(26, 25), (40, 51), (128, 97)])
def test_gradient(orc, shape):
    from panorama_opticalflow_b200 import stages
    a = RNG.random(shape).astype(np.float32)
    assert_bit_equal(stages.gradient(a), _grad(orc, a), "gradient")


# even widths >= 46 make interior tiles go through the TMA tensor copy, odd widths / small images through per-thread loads
@pytest.mark.parametrize("shape", [(26, 25), (45, 61), (130, 99), (150, 200), (97, 46), (46, 160)])
def test_blur15_and_diffusion(orc, shape):
    from panorama_opticalflow_b200 import stages
    f = (RNG.standard_normal(shape + (2,)) * 3).astype(np.float32)
    assert_bit_equal(stages.blur15(f), orc.gaussian_blur(f, 15, 8.0), "blur15")
    a0 = RNG.random(shape).astype(np.float32)
    a1 = (RNG.random(shape) > 0.3).astype(np.float32)
    assert_bit_equal(stages.blur15(f, a0, a1), orc.low_alpha_diffusion(a0, a1, f), "diffusion")


@pytest.mark.parametrize("shape", [(25, 26), (37, 45), (120, 131), (64, 200), (12, 36), (99, 130)])
def test_median5(orc, shape):
    from panorama_opticalflow_b200 import stages
    f = RNG.standard_normal(shape + (2,)).astype(np.float32)
    f[RNG.random(shape) < 0.2] = 0.0   # ties
    assert_bit_equal(stages.median5(f), orc.median5_c2(f), "median5")


@pytest.mark.parametrize("src,dst", [((45, 25), (50, 28)), ((50, 60), (56, 67)), ((101, 77), (112, 85)), ((90, 81), (100, 90))])
def test_upsample_cubic(orc, src, dst):
    from panorama_opticalflow_b200 import stages
    f = RNG.standard_normal(src + (2,)).astype(np.float32)
    want = orc.resize_cubic_f32(f, dst[0], dst[1]) * (np.float32(1.0) / np.float32(0.9))
    assert_bit_equal(stages.upsample_cubic(f, dst[0], dst[1]), want, "upsample")


@pytest.mark.parametrize("sh,sw,rows,cols,pad", [(40, 50, 80, 100, 0), (40, 55, 81, 100, 5), (33, 47, 67, 85, 4)])
def test_tail(orc, sh, sw, rows, cols, pad):
    from panorama_opticalflow_b200 import stages
    f = RNG.standard_normal((sh, sw, 2)).astype(np.float32)
    pc = cols + 2 * pad
    up = orc.resize_linear(f, rows, pc) * np.float32(2.0)
    want = orc.gaussian_blur(up, 3, 1.0)[:, pad:pad + cols]
    assert_bit_equal(stages.tail(f, rows, pc, pad, cols), want, "tail")


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("hint", [1, 2, 3, 4, 0])
def test_initial_flow(orc, sparse, hint):
    from panorama_opticalflow_b200 import stages
    I0, I1, A0, A1 = _pair_planes(orc, 160, 200, 5, 30.0, sparse, level=10)
    dist = orc.search_distance(20)
    want = orc.adjust_initial_flow(I0, I1, A0, A1, hint, dist) if hint else np.zeros(I0.shape + (2,), np.float32)
    got = stages.initial_flow(I0, I1, A0, A1, hint, dist)
    assert_bit_equal(got, want, "adjustInitialFlow")
    if hint in (1, 3) and not sparse:
        assert np.abs(want).max() > 0   # the search actually fires


@pytest.mark.parametrize("direction", [+1, -1])
@pytest.mark.parametrize("case", [(128, 160, 0, 6.0, False, 0), (181, 243, 1, 24.0, False, 0), (160, 220, 2, 20.0, True, 0),
                                  (400, 300, 3, 10.0, False, 0), (400, 300, 3, 10.0, True, 1), (70, 60, 4, 3.0, False, 0)])
def test_sweep(orc, case, direction):
    """One Gauss-Seidel sweep as an anti-diagonal wavefront == the reference raster-order sweep, bit for bit."""
    from panorama_opticalflow_b200 import stages
    rows, cols, seed, amp, sparse, level = case
    I0, I1, A0, A1 = _pair_planes(orc, rows, cols, seed, amp, sparse, level)
    G0, G1 = _grad(orc, I0), _grad(orc, I1)
    h, w = I0.shape
    flow = (np.random.default_rng(seed).standard_normal((h, w, 2)) * 0.7).astype(np.float32)
    flow[..., 0] -= np.float32(amp / 4)
    blurred = orc.gaussian_blur(flow, 15, 8.0)
    want = orc.sweep(A0, A1, G0[..., 0], G0[..., 1], G1[..., 0], G1[..., 1], blurred, flow, direction)
    got = stages.sweep(A0, A1, G0, G1, blurred, flow, direction)
    assert_bit_equal(got, want, "sweep dir %+d" % direction)
    assert not np.array_equal(want, flow)


@pytest.mark.parametrize("direction", [+1, -1])
def test_sweep_with_nan_and_inf_flows(orc, direction):
    """Non-finite flows cannot arise from 8-bit inputs, but the sweep must still follow the reference on them: a pixel whose
    own error is NaN stays updatable (no proposal can win against NaN, the gradient step writes NaN) and hands NaN on;
    inf / huge flows leave the range of the branch-free exact sequences and take the IEEE-intrinsic path."""
    from panorama_opticalflow_b200 import stages
    I0, I1, A0, A1 = _pair_planes(orc, 200, 260, 7, 8.0, False, 0)
    G0, G1 = _grad(orc, I0), _grad(orc, I1)
    h, w = I0.shape
    rng = np.random.default_rng(7)
    flow = (rng.standard_normal((h, w, 2)) * 0.7).astype(np.float32)
    for k, v in enumerate((np.nan, np.inf, -np.inf, 1e30, 1e-42, -0.0)):
        ys, xs = rng.integers(0, h, 12), rng.integers(0, w, 12)
        flow[ys, xs, k % 2] = v
    blurred = orc.gaussian_blur(np.nan_to_num(flow, nan=0.0, posinf=0.0, neginf=0.0), 15, 8.0)
    want = orc.sweep(A0, A1, G0[..., 0], G0[..., 1], G1[..., 0], G1[..., 1], blurred, flow, direction)
    got = stages.sweep(A0, A1, G0, G1, blurred, flow, direction)
    assert np.isnan(want).any()
    # NaN payloads are not part of the contract: compare NaN-ness, and bits everywhere else
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.array_equal(got[m].view(np.uint32), want[m].view(np.uint32)), int((got[m].view(np.uint32) != want[m].view(np.uint32)).sum())
