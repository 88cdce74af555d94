"""First "next" row: Stitchtools::prepare without the blend smoothing -- canvas map, overlap masking,
un-smoothed blend (countblend) and MergedDis.  Byte/index work: bit-exact against the oracle."""
import numpy as np
import pytest

from conftest import assert_bit_equal


def _canvas_pair(rows, cols, seed, wrap=False):
    """two RGBA images covering different parts of a canvas, alpha strictly {0, 255} like the reference's test data"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:rows, 0:cols]
    L = rng.integers(1, 256, (rows, cols, 4), dtype=np.uint8)
    R = rng.integers(1, 256, (rows, cols, 4), dtype=np.uint8)
    if wrap:   # R wraps around the 360-degree seam
        aL = (x > 0.25 * cols) & (x < 0.7 * cols) & (y > 0.1 * rows)
        aR = ((x < 0.35 * cols) | (x > 0.6 * cols)) & (y < 0.9 * rows)
    else:
        aL = ((x - 0.4 * cols) ** 2 / (0.3 * cols) ** 2 + (y - 0.5 * rows) ** 2 / (0.45 * rows) ** 2) < 1
        aR = (x > 0.45 * cols + 20 * np.sin(y / 23.0)) & (y > 0.05 * rows)
    L[..., 3] = np.where(aL, 255, 0)
    R[..., 3] = np.where(aR, 255, 0)
    return L, R


@pytest.mark.parametrize("rows,cols,wrap", [(240, 330, False), (400, 250, True)])
def test_oracle_map_and_mask_against_numpy(orc, rows, cols, wrap):
    L, R = _canvas_pair(rows, cols, 3, wrap)
    m, oL, oR = orc.stitch_match_and_mask(L, R)
    mm = (np.where(L[..., 3] > 0, 100, 0) + np.where(R[..., 3] > 0, 50, 0)).astype(np.uint8)
    assert np.array_equal(m, mm)
    keep = (mm > 140)[..., None].astype(np.uint8)
    assert np.array_equal(oL, L * keep) and np.array_equal(oR, R * keep)
    blend, md = orc.stitch_blend_raw(m)
    assert np.all(blend[mm == 100] == 0) and np.all(blend[mm == 50] == 1) and np.all(blend[mm == 0] == 0.5)
    ov = mm == 150
    assert ov.any() and np.all((blend[ov] >= 0) & (blend[ov] <= 1)) and np.all(md[~ov] == 0)
    with pytest.raises(ValueError):
        orc.stitch_blend_raw(m[:150, :150])


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,wrap", [(240, 330, False), (400, 250, True), (601, 403, False), (333, 1000, True)])
def test_stitch_prepare_vs_oracle(orc, engine_low, rows, cols, wrap):
    import panorama_opticalflow_b200 as pf
    L, R = _canvas_pair(rows, cols, 7, wrap)
    st = pf.Stitchtools(engine_low)
    st.prepare(L, R)
    m, oL, oR = orc.stitch_match_and_mask(L, R)
    blend, md = orc.stitch_blend_raw(m)
    assert_bit_equal(st.getMap(), m, "Map")
    assert_bit_equal(st.getOverlappedL(), oL, "OverlappedL")
    assert_bit_equal(st.getOverlappedR(), oR, "OverlappedR")
    assert_bit_equal(st.getBlendUnsmoothed(), blend, "blend (un-smoothed)")
    assert_bit_equal(st.MergedDis, md, "MergedDis")


@pytest.mark.gpu
def test_stitch_prepare_rejects_small_images(engine_low):
    import panorama_opticalflow_b200 as pf
    L, R = _canvas_pair(150, 300, 1)
    with pytest.raises(pf.PixFlowError) as ei:
        pf.Stitchtools(engine_low).prepare(L, R)
    assert ei.value.code == 1 and "too small" in str(ei.value)


@pytest.mark.gpu
def test_stitch_then_flow_then_blend_like_main(orc, engine_search):
    """CPU/main.cpp:67-89 up to generateNovelView, with the un-smoothed blend standing in for getBlend()."""
    import panorama_opticalflow_b200 as pf
    L, R = _canvas_pair(260, 340, 11)
    st = pf.Stitchtools(engine_search)
    st.prepare(L, R)
    gen = pf.NovelViewGeneratorAsymmetricFlow("pixflow_search_20")
    gen.prepare(st.getOverlappedL(), st.getOverlappedR())
    gen.setBlend(st.getBlendUnsmoothed())
    merged = gen.generateNovelView()
    m, oL, oR = orc.stitch_match_and_mask(L, R)
    blend, _ = orc.stitch_blend_raw(m)
    fLR, fRL = orc.prepare_bidirectional(oL, oR, 20)
    assert_bit_equal(gen.getFlowLtoR(), fLR, "flowLtoR")
    want = orc.combine_novel_views(oL, oR, fLR, fRL, blend)
    assert np.abs(merged.astype(int) - want.astype(int)).max() <= 1
    gen.close()
