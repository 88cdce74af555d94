"""Second, independent composition of the reference hot path: the OpenCV calls are made through the
same-named cv2 4.13.0 functions in scalar mode, the reference's own loops through oracle/pixflow_oracle.c.
TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/orc.py).  Follows SURVEY.md Appendix C row by row.

Used by tests/test_oracle_vs_cv2.py to pin the pure-C oracle (which is what travels to the GPU box and
what bench.py times as the CPU baseline) against the real OpenCV arithmetic.
"""
import cv2
import numpy as np

from . import orc

cv2.setUseOptimized(False)  # scalar code paths: deterministic, fully characterised

F32 = np.float32


def frontend(bgra):
    """CPU/PixFlow.hpp:78-103"""
    rows, cols = bgra.shape[:2]
    dw = int(F32(cols) * F32(0.5))
    dh = int(F32(rows) * F32(0.5))
    small = cv2.resize(bgra, (dw, dh), interpolation=cv2.INTER_CUBIC)
    grey = cv2.cvtColor(small, cv2.COLOR_BGRA2GRAY)
    inv255 = F32(1.0 / 255.0)
    I = grey.astype(F32) * inv255
    A = small[..., 3].astype(F32) * inv255
    I = cv2.GaussianBlur(I, (5, 5), 0.25)
    return I, A


def build_pyramid(img):
    """CPU/PixFlow.hpp:137-151"""
    pyr = [img]
    while len(pyr) < 1000:
        h, w = pyr[-1].shape[:2]
        nw = int(F32(w) * F32(0.9) + F32(0.5))
        nh = int(F32(h) * F32(0.9) + F32(0.5))
        if nh <= 24 or nw <= 24:
            break
        pyr.append(cv2.resize(pyr[-1], (nw, nh), interpolation=cv2.INTER_LINEAR))
    return pyr


def gradients(I):
    """CPU/PixFlow.hpp:284-294"""
    gx = cv2.Sobel(I, -1, 1, 0, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    gy = cv2.Sobel(I, -1, 0, 1, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    return cv2.GaussianBlur(gx, (3, 3), 0.5), cv2.GaussianBlur(gy, (3, 3), 0.5)


def level(I0, I1, a0, a1, flow, hint, max_percentage, trace=None, lvl=0):
    """CPU/PixFlow.hpp:272-340"""
    I0x, I0y = gradients(I0)
    I1x, I1y = gradients(I1)
    if flow is None:
        if max_percentage > 0 and hint != orc.HINT_UNKNOWN:
            flow = orc.adjust_initial_flow(I0, I1, a0, a1, hint, orc.search_distance(max_percentage))
        else:
            flow = np.zeros(I0.shape + (2,), F32)
    blurred = cv2.GaussianBlur(flow, (15, 15), 8.0)
    if trace is not None:
        trace[(lvl, "flow_in")] = flow.copy()
        trace[(lvl, "blurred")] = blurred.copy()
    flow = orc.sweep(a0, a1, I0x, I0y, I1x, I1y, blurred, flow, +1)
    if trace is not None:
        trace[(lvl, "fwd")] = flow.copy()
    flow = cv2.medianBlur(flow, 5)
    flow = orc.sweep(a0, a1, I0x, I0y, I1x, I1y, blurred, flow, -1)
    if trace is not None:
        trace[(lvl, "bwd")] = flow.copy()
    flow = cv2.medianBlur(flow, 5)
    # lowAlphaFlowDiffusion, CPU/PixFlow.hpp:388-405
    bl = cv2.GaussianBlur(flow, (15, 15), 8.0)
    d = (F32(1.0) - a0 * a1)[..., None]
    flow = d * bl + (F32(1.0) - d) * flow
    if trace is not None:
        trace[(lvl, "diffused")] = flow.copy()
    return flow


def compute_flow(i0, i1, max_percentage, hint, trace=None):
    """CPU/PixFlow.hpp:72-135"""
    rows, cols = i0.shape[:2]
    I0, A0 = frontend(i0)
    I1, A1 = frontend(i1)
    pI0, pI1, pA0, pA1 = (build_pyramid(x) for x in (I0, I1, A0, A1))
    flow = None
    for l in range(len(pI0) - 1, -1, -1):
        flow = level(pI0[l], pI1[l], pA0[l], pA1[l], flow, hint, max_percentage, trace, l)
        if l > 0:
            h, w = pI0[l - 1].shape
            flow = cv2.resize(flow, (w, h), interpolation=cv2.INTER_CUBIC) * F32(F32(1.0) / F32(0.9))
    flow = cv2.resize(flow, (cols, rows), interpolation=cv2.INTER_LINEAR) * F32(F32(1.0) / F32(0.5))
    return cv2.GaussianBlur(flow, (3, 3), 1.0)


def prepare_bidirectional(L, R, max_percentage):
    """CPU/OpticalFlow.cpp:102-145"""
    cols = L.shape[1]
    l = cols // 20
    pad = lambda M: np.ascontiguousarray(np.concatenate([M[:, cols - l:], M, M[:, :l]], 1))
    nL, nR = pad(L), pad(R)
    fLR = compute_flow(nL, nR, max_percentage, orc.HINT_LEFT)
    fRL = compute_flow(nR, nL, max_percentage, orc.HINT_RIGHT)
    return np.ascontiguousarray(fLR[:, l:l + cols]), np.ascontiguousarray(fRL[:, l:l + cols])
