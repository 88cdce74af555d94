"""Seeded synthetic overlap pairs (numpy only) -- the workload generator of SURVEY.md section 8d / App. C.

A smooth random texture plus colour ramps forms image L; R is L resampled through a smooth disparity
field d(x,y) = A*(0.5+0.5*sin(y/97)*cos(x/131)), so the true L->R flow is (-d, -0.3d): this matches the
LEFT hint's search box for L->R and the RIGHT hint for R->L (CPU/OpticalFlow.cpp:130-139).
"""
import numpy as np


def _gauss1d(sigma):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def _blur(img, sigma):
    k = _gauss1d(sigma)
    r = len(k) // 2
    out = img
    for axis in (0, 1):
        p = np.pad(out, [(r, r) if a == axis else (0, 0) for a in range(out.ndim)], mode="reflect")
        acc = np.zeros_like(out)
        n = out.shape[axis]
        for i, kv in enumerate(k):
            sl = [slice(None)] * out.ndim
            sl[axis] = slice(i, i + n)
            acc += kv * p[tuple(sl)]
        out = acc
    return out


def _remap_linear(img, mx, my):
    h, w = img.shape[:2]

    def refl(v, n):
        v = np.abs(v)
        v = np.where(v > n - 1, 2 * (n - 1) - v, v)
        return np.clip(v, 0, n - 1)

    mx = refl(mx, w)
    my = refl(my, h)
    x0 = np.floor(mx).astype(np.int64)
    y0 = np.floor(my).astype(np.int64)
    fx = (mx - x0).astype(np.float32)[..., None]
    fy = (my - y0).astype(np.float32)[..., None]
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    top = img[y0, x0] * (1 - fx) + img[y0, x1] * fx
    bot = img[y1, x0] * (1 - fx) + img[y1, x1] * fx
    return top * (1 - fy) + bot * fy


def make_pair(rows, cols, seed=0, amplitude=6.0, sparse=False):
    """Returns (L, R): two uint8 BGRA images rows x cols.  sparse=True zeroes alpha outside a blob."""
    rng = np.random.default_rng(seed)
    m = 32
    H, W = rows + 2 * m, cols + 2 * m
    tex = _blur(rng.random((H, W, 3), dtype=np.float32), 3.0)
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    xr = (np.arange(W, dtype=np.float32) / W)[None, :, None]
    yr = (np.arange(H, dtype=np.float32) / H)[:, None, None]
    img = 0.6 * tex + 0.4 * (xr * np.array([1.0, 0.5, 0.2], np.float32) + yr * np.array([0.1, 0.4, 0.8], np.float32))
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    d = amplitude * (0.5 + 0.5 * np.sin(yy / 97.0) * np.cos(xx / 131.0))
    imgR = _remap_linear(img, xx + d, yy + 0.3 * d)

    def to_bgra(a):
        a = a[m:m + rows, m:m + cols]
        u = np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)
        return np.concatenate([u, np.full((rows, cols, 1), 255, np.uint8)], axis=2)

    L, R = to_bgra(img), to_bgra(imgR)
    if sparse:
        y, x = np.mgrid[0:rows, 0:cols]
        inside = ((x - 0.55 * cols) / (0.35 * cols)) ** 2 + ((y - 0.5 * rows) / (0.42 * rows)) ** 2 < 1.0
        insideR = ((x - 0.45 * cols) / (0.38 * cols)) ** 2 + ((y - 0.5 * rows) / (0.45 * rows)) ** 2 < 1.0
        L = L * inside[..., None].astype(np.uint8)
        R = R * insideR[..., None].astype(np.uint8)
    return np.ascontiguousarray(L), np.ascontiguousarray(R)


def make_blend(rows, cols):
    """A smooth left-to-right blend-weight ramp in [0,1] (stand-in for Stitchtools::GenerateBlend)."""
    ramp = np.linspace(0.0, 1.0, cols, dtype=np.float32)[None, :]
    return np.ascontiguousarray(np.repeat(ramp, rows, axis=0))
