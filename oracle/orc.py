"""ctypes binding of the CPU oracle (oracle/pixflow_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborc.so")

HINT_UNKNOWN, HINT_RIGHT, HINT_DOWN, HINT_LEFT, HINT_UP = 0, 1, 2, 3, 4  # CPU/PixFlow.hpp:19

TRACE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int)
STAGE_NAMES = {0: "blurred", 1: "fwd", 2: "fwd_median", 3: "bwd", 4: "bwd_median", 5: "diffused", 6: "flow_in"}


def build(force=False):
    src = os.path.join(_HERE, "pixflow_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liborc.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_intensity_ratio.restype = C.c_float
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def _u(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def gaussian_kernel(k, sigma):
    out = np.empty(k, np.float32)
    lib().orc_gaussian_kernel(C.c_int(k), C.c_double(sigma), out.ctypes.data_as(C.c_void_p))
    return out


def gaussian_blur(src, k, sigma):
    src, p = _f(src)
    h, w = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    lib().orc_gaussian_blur(p, h, w, ch, k, C.c_double(sigma), dst.ctypes.data_as(C.c_void_p))
    return dst


def sobel(src, dx):
    src, p = _f(src)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().orc_sobel(p, h, w, int(dx), dst.ctypes.data_as(C.c_void_p))
    return dst


def median5_c2(src):
    src, p = _f(src)
    h, w, _ = src.shape
    dst = np.empty_like(src)
    lib().orc_median5_c2(p, h, w, dst.ctypes.data_as(C.c_void_p))
    return dst


def bgra2gray(bgra):
    bgra, p = _u(bgra)
    h, w, _ = bgra.shape
    dst = np.empty((h, w), np.uint8)
    lib().orc_bgra2gray(p, C.c_size_t(h * w), dst.ctypes.data_as(C.c_void_p))
    return dst


def resize_linear(src, dh, dw):
    src, p = _f(src)
    sh, sw = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.float32)
    lib().orc_resize_linear(p, sh, sw, ch, dst.ctypes.data_as(C.c_void_p), dh, dw)
    return dst


def resize_cubic_f32(src, dh, dw):
    src, p = _f(src)
    sh, sw = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.float32)
    lib().orc_resize_cubic_f32(p, sh, sw, ch, dst.ctypes.data_as(C.c_void_p), dh, dw)
    return dst


def resize_cubic_u8c4(src, dh, dw):
    src, p = _u(src)
    sh, sw, _ = src.shape
    dst = np.empty((dh, dw, 4), np.uint8)
    lib().orc_resize_cubic_u8c4(p, sh, sw, C.c_size_t(sw * 4), dst.ctypes.data_as(C.c_void_p), dh, dw)
    return dst


def downscale_size(rows, cols):
    dh, dw = C.c_int(), C.c_int()
    lib().orc_downscale_size(rows, cols, C.byref(dh), C.byref(dw))
    return dh.value, dw.value


def pyramid_sizes(w0, h0):
    ws = (C.c_int * 128)()
    hs = (C.c_int * 128)()
    n = lib().orc_pyramid_sizes(w0, h0, ws, hs, 128)
    return [(ws[i], hs[i]) for i in range(n)]


def search_distance(max_percentage):
    return lib().orc_search_distance(max_percentage)


def sweep(alpha0, alpha1, I0x, I0y, I1x, I1y, blurred, flow, direction):
    """In-place reference sweep on a copy of `flow`; direction +1 (top/left) or -1 (bottom/right)."""
    a0, pa0 = _f(alpha0)
    a1, pa1 = _f(alpha1)
    g0, pg0 = _f(I0x)
    g1, pg1 = _f(I0y)
    g2, pg2 = _f(I1x)
    g3, pg3 = _f(I1y)
    bl, pbl = _f(blurred)
    out = np.array(flow, dtype=np.float32, order="C", copy=True)
    h, w = a0.shape
    lib().orc_sweep(pa0, pa1, pg0, pg1, pg2, pg3, pbl, out.ctypes.data_as(C.c_void_p), h, w, int(direction))
    return out


def low_alpha_diffusion(alpha0, alpha1, flow):
    a0, pa0 = _f(alpha0)
    a1, pa1 = _f(alpha1)
    out = np.array(flow, dtype=np.float32, order="C", copy=True)
    h, w = a0.shape
    lib().orc_low_alpha_diffusion(pa0, pa1, out.ctypes.data_as(C.c_void_p), h, w)
    return out


def intensity_ratio(I0, a0, I1, a1):
    I0, p0 = _f(I0)
    a0, pa0 = _f(a0)
    I1, p1 = _f(I1)
    a1, pa1 = _f(a1)
    h, w = I0.shape
    return np.float32(lib().orc_intensity_ratio(p0, pa0, p1, pa1, h, w))


def adjust_initial_flow(I0, I1, alpha0, alpha1, hint, dist):
    I0, p0 = _f(I0)
    I1, p1 = _f(I1)
    a0, pa0 = _f(alpha0)
    a1, pa1 = _f(alpha1)
    h, w = I0.shape
    flow = np.zeros((h, w, 2), np.float32)
    lib().orc_adjust_initial_flow(p0, p1, pa0, pa1, flow.ctypes.data_as(C.c_void_p), h, w, int(hint), int(dist))
    return flow


def frontend(bgra):
    bgra, p = _u(bgra)
    rows, cols, _ = bgra.shape
    dh, dw = downscale_size(rows, cols)
    I = np.empty((dh, dw), np.float32)
    A = np.empty((dh, dw), np.float32)
    lib().orc_frontend(p, rows, cols, C.c_size_t(cols * 4), I.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p))
    return I, A


def level(I0, I1, a0, a1, flow, first, hint, max_percentage):
    I0, p0 = _f(I0)
    I1, p1 = _f(I1)
    a0, pa0 = _f(a0)
    a1, pa1 = _f(a1)
    h, w = I0.shape
    out = np.zeros((h, w, 2), np.float32) if flow is None else np.array(flow, dtype=np.float32, order="C", copy=True)
    lib().orc_level(p0, p1, pa0, pa1, out.ctypes.data_as(C.c_void_p), h, w, int(first), int(hint),
                    int(max_percentage), None, None, 0)
    return out


def compute_flow(i0, i1, max_percentage, hint, trace=None):
    """PixFlow<max_percentage>::computeOpticalFlow.  trace: optional dict filled with
    {(level, stage_name): array} for every per-level stage."""
    i0, p0 = _u(i0)
    i1, p1 = _u(i1)
    rows, cols, _ = i0.shape
    assert i1.shape == i0.shape
    flow = np.empty((rows, cols, 2), np.float32)
    cb = None
    if trace is not None:
        def _cb(_user, lvl, stage, data, h, w, ch):
            arr = np.ctypeslib.as_array(data, shape=(h, w, ch)).copy()
            trace[(lvl, STAGE_NAMES[stage])] = arr
        cb = TRACE_FN(_cb)
    lib().orc_compute_flow(p0, C.c_size_t(cols * 4), p1, C.c_size_t(cols * 4), rows, cols,
                           int(max_percentage), int(hint), flow.ctypes.data_as(C.c_void_p),
                           cb if cb is not None else C.cast(None, TRACE_FN), None)
    return flow


def prepare_bidirectional(L, R, max_percentage, threads=1):
    """NovelViewGeneratorAsymmetricFlow::prepare.  threads=2 runs the two (independent) directions on two host threads --
    identical results; process-wide setting, so do not mix values across concurrent callers."""
    lib().orc_set_prepare_threads(int(threads))
    L, pl = _u(L)
    R, pr = _u(R)
    rows, cols, _ = L.shape
    fLR = np.empty((rows, cols, 2), np.float32)
    fRL = np.empty((rows, cols, 2), np.float32)
    lib().orc_prepare_bidirectional(pl, C.c_size_t(cols * 4), pr, C.c_size_t(cols * 4), rows, cols,
                                    int(max_percentage), fLR.ctypes.data_as(C.c_void_p),
                                    fRL.ctypes.data_as(C.c_void_p))
    return fLR, fRL


def combine_novel_views(imageL, imageR, flowLtoR, flowRtoL, blend):
    L, pl = _u(imageL)
    R, pr = _u(imageR)
    fLR, plr = _f(flowLtoR)
    fRL, prl = _f(flowRtoL)
    bl, pb = _f(blend)
    rows, cols, _ = L.shape
    out = np.empty((rows, cols, 4), np.uint8)
    lib().orc_combine_novel_views(pl, C.c_size_t(cols * 4), pr, C.c_size_t(cols * 4), plr, prl, pb, rows, cols,
                                  out.ctypes.data_as(C.c_void_p))
    return out


def stitch_match_and_mask(imageL, imageR):
    """Stitchtools::MatchImages + the overlap masking of Stitchtools::prepare (CPU/StitchTool.cpp:7-50)."""
    L, pl = _u(imageL)
    R, pr = _u(imageR)
    rows, cols, _ = L.shape
    m = np.empty((rows, cols), np.uint8)
    oL = np.empty((rows, cols, 4), np.uint8)
    oR = np.empty((rows, cols, 4), np.uint8)
    lib().orc_stitch_match_and_mask(pl, C.c_size_t(cols * 4), pr, C.c_size_t(cols * 4), rows, cols,
                                    m.ctypes.data_as(C.c_void_p), oL.ctypes.data_as(C.c_void_p), oR.ctypes.data_as(C.c_void_p))
    return m, oL, oR


def stitch_blend_raw(map_u8):
    """GenerateBlend before the smoothing (CPU/StitchTool.cpp:98-131, countblend :148-191) -> (blend, MergedDis)."""
    m, pm = _u(map_u8)
    rows, cols = m.shape
    blend = np.empty((rows, cols), np.float32)
    md = np.empty((rows, cols), np.float32)
    rc = lib().orc_stitch_blend_raw(pm, rows, cols, blend.ctypes.data_as(C.c_void_p), md.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("image too small: the reference's search step cols/200 (rows/200) is 0")
    return blend, md


def box_blur_rect(parent, x0, y0, rw, rh, k):
    """cv::blur with a k x k kernel on the rectangle (x0, y0, rw, rh) of `parent` taken as a ROI (window pixels come
    from the parent, reflect-101 beyond its edge); (0, 0, cols, rows) = cv::blur of the whole image."""
    p, pp = _f(parent)
    rows, cols = p.shape
    out = np.empty((rh, rw), np.float32)
    lib().orc_box_blur_rect(pp, rows, cols, x0, y0, rw, rh, k, out.ctypes.data_as(C.c_void_p))
    return out


def stitch_blend_smooth(blend_raw, merged_dis):
    """The smoothing of GenerateBlend (CPU/StitchTool.cpp:133-145) -> Blend."""
    b = np.array(blend_raw, np.float32, order="C", copy=True)
    md, pmd = _f(merged_dis)
    rows, cols = b.shape
    rc = lib().orc_stitch_blend_smooth(b.ctypes.data_as(C.c_void_p), pmd, rows, cols)
    if rc != 0:
        raise ValueError("image too small: rows/400 == 0 (cv::blur with an empty kernel) or search step 0")
    return b


def stitch_gather(imageL, imageR, merged, map_u8):
    """Stitchtools::Gather (CPU/StitchTool.cpp:52-96) -> FinalResult."""
    L, pl = _u(imageL)
    R, pr = _u(imageR)
    M, pm = _u(merged)
    m, pmap = _u(map_u8)
    rows, cols = m.shape
    out = np.empty((rows, cols, 4), np.uint8)
    lib().orc_stitch_gather(pl, pr, pm, pmap, rows, cols, out.ctypes.data_as(C.c_void_p))
    return out


def stitch_iteration(imageL, imageR, max_percentage=20, threads=1):
    """One iteration of the reference driver (CPU/main.cpp:72-92): Stitchtools::prepare -> flow prepare -> setBlend ->
    generateNovelView -> setMergedmiddle -> Gather.  Returns (FinalResult, dict of intermediates).
    threads > 1: the blend map and the two flow directions -- independent computations the reference runs one after the
    other -- run on separate host threads (ctypes releases the GIL); the results are identical."""
    m, oL, oR = stitch_match_and_mask(imageL, imageR)
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as ex:
            fut = ex.submit(lambda: stitch_blend_raw(m))
            fLR, fRL = prepare_bidirectional(oL, oR, max_percentage, threads=2)
            braw, md = fut.result()
    else:
        braw, md = stitch_blend_raw(m)
        fLR, fRL = prepare_bidirectional(oL, oR, max_percentage)
    blend = stitch_blend_smooth(braw, md)
    merged = combine_novel_views(oL, oR, fLR, fRL, blend)
    final = stitch_gather(imageL, imageR, merged, m)
    return final, dict(map=m, overlappedL=oL, overlappedR=oR, blend_raw=braw, merged_dis=md, blend=blend,
                       flowLtoR=fLR, flowRtoL=fRL, merged=merged)


def four_input_frontend(img1, img2, img3, img4):
    """CPU_4Input/main.cpp:64-79 -> (colorImageL, colorImageR)"""
    arrs = [_u(a)[0] for a in (img1, img2, img3, img4)]
    rows, cols, _ = arrs[0].shape
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
    L = np.empty((rows, cols, 4), np.uint8)
    R = np.empty((rows, cols, 4), np.uint8)
    lib().orc_four_input_frontend(ptrs, rows, cols, L.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    return L, R
