"""numpy wrappers of the diagnostic single-stage entry points (pf_stage_*), used by tests/ to localise a
divergence to one kernel.  Host arrays in, host arrays out; one kernel launch each."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def frontend(bgra, pad=0):
    bgra = np.ascontiguousarray(bgra, dtype=np.uint8)
    rows, cols, _ = bgra.shape
    pc = cols + 2 * pad
    dw = int(np.float32(pc) * np.float32(0.5))
    dh = int(np.float32(rows) * np.float32(0.5))
    grey = np.empty((dh, dw), np.float32)
    alpha = np.empty((dh, dw), np.float32)
    _lib.check(_lib.load().pf_stage_frontend(_p(bgra), rows, cols, pad, _p(grey), _p(alpha), dh, dw))
    return grey, alpha


def gauss5(src):
    src = _f32(src)
    dst = np.empty_like(src)
    _lib.check(_lib.load().pf_stage_gauss5(_p(src), _p(dst), *src.shape))
    return dst


def pyr_down(src, dh, dw):
    src = _f32(src)
    dst = np.empty((dh, dw), np.float32)
    _lib.check(_lib.load().pf_stage_pyr_down(_p(src), src.shape[0], src.shape[1], _p(dst), dh, dw))
    return dst


def gradient(I):
    I = _f32(I)
    G = np.empty(I.shape + (2,), np.float32)
    _lib.check(_lib.load().pf_stage_gradient(_p(I), _p(G), *I.shape))
    return G


def blur15(flow, alpha0=None, alpha1=None):
    flow = _f32(flow)
    dst = np.empty_like(flow)
    h, w, _ = flow.shape
    if alpha0 is None:
        _lib.check(_lib.load().pf_stage_blur15(_p(flow), _p(dst), h, w, None, None))
    else:
        a0, a1 = _f32(alpha0), _f32(alpha1)
        _lib.check(_lib.load().pf_stage_blur15(_p(flow), _p(dst), h, w, _p(a0), _p(a1)))
    return dst


def median5(flow):
    flow = _f32(flow)
    dst = np.empty_like(flow)
    _lib.check(_lib.load().pf_stage_median5(_p(flow), _p(dst), flow.shape[0], flow.shape[1]))
    return dst


def sweep(alpha0, alpha1, G0, G1, blurred, flow, direction):
    a0, a1, g0, g1, bl = _f32(alpha0), _f32(alpha1), _f32(G0), _f32(G1), _f32(blurred)
    out = np.array(flow, dtype=np.float32, order="C", copy=True)
    h, w = a0.shape
    _lib.check(_lib.load().pf_stage_sweep(_p(a0), _p(a1), _p(g0), _p(g1), _p(bl), _p(out), h, w, int(direction)))
    return out


def upsample_cubic(src, dh, dw):
    src = _f32(src)
    dst = np.empty((dh, dw, 2), np.float32)
    _lib.check(_lib.load().pf_stage_upsample_cubic(_p(src), src.shape[0], src.shape[1], _p(dst), dh, dw))
    return dst


def tail(flow0, rows, pcols, pad, cols):
    flow0 = _f32(flow0)
    out = np.empty((rows, cols, 2), np.float32)
    _lib.check(_lib.load().pf_stage_tail(_p(flow0), flow0.shape[0], flow0.shape[1], rows, pcols, pad, cols, _p(out)))
    return out


def initial_flow(I0, I1, alpha0, alpha1, hint, dist):
    I0, I1, a0, a1 = _f32(I0), _f32(I1), _f32(alpha0), _f32(alpha1)
    h, w = I0.shape
    flow = np.empty((h, w, 2), np.float32)
    _lib.check(_lib.load().pf_stage_initial_flow(_p(I0), _p(I1), _p(a0), _p(a1), _p(flow), h, w, int(hint), int(dist)))
    return flow
