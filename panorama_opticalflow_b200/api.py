"""Host-side mirror of the reference's operator interface for the flow / novel-view path, over the C-ABI.

Same names, argument meaning and error behaviour as the reference (C++ names kept on purpose):
  OpticalFlowInterface / DirectionHint / makeOpticalFlowByName      CPU/PixFlow.hpp:15-26, :459-500
  NovelViewUtil.combineNovelViews                                    CPU/OpticalFlow.hpp:26-31
  NovelViewGenerator / NovelViewGeneratorAsymmetricFlow              CPU/OpticalFlow.hpp:34-70
Images are numpy uint8 (rows, cols, 4) BGRA arrays or CUDA tensors exposing data_ptr()/stride()
(device-resident, zero copy); flows are float32 (rows, cols, 2); blend float32 (rows, cols).
The C++ twin of this file is include/pixflow_b200.hpp.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import PixFlowError  # noqa: F401  (re-export)


class DirectionHint(enum.IntEnum):
    """OpticalFlowInterface::DirectionHint, CPU/PixFlow.hpp:19"""
    UNKNOWN = 0
    RIGHT = 1
    DOWN = 2
    LEFT = 3
    UP = 4


def _is_device_tensor(a):
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda")


def _view(a, dtype, channels, name, out=False):
    """-> (keepalive, pointer, row_stride_bytes, rows, cols).  Inputs whose pixels are not densely packed are copied;
    an OUTPUT (out=True) must be usable in place -- a copy would leave the caller's buffer unfilled -- so it raises."""
    if _is_device_tensor(a):
        if not a.is_cuda:
            a = a.numpy()
        else:
            exp = {np.uint8: "torch.uint8", np.float32: "torch.float32"}[dtype]
            if str(a.dtype) != exp:
                raise TypeError("%s must be %s, got %s" % (name, exp, a.dtype))
            if a.stride(-1) != 1 or (channels > 1 and (a.dim() != 3 or a.shape[2] != channels or a.stride(1) != channels)):
                raise ValueError("%s must have densely packed pixels" % name)
            return a, C.c_void_p(a.data_ptr()), a.stride(0) * a.element_size(), a.shape[0], a.shape[1]
    a = np.asarray(a)
    if a.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, np.dtype(dtype), a.dtype))
    want_ndim = 3 if channels > 1 else 2
    if a.ndim != want_ndim or (channels > 1 and a.shape[2] != channels):
        raise ValueError("%s must have shape (rows, cols%s)" % (name, ", %d" % channels if channels > 1 else ""))
    item = a.itemsize
    inner_ok = a.strides[-1] == item and (channels == 1 or a.strides[1] == item * channels)
    if not inner_ok or a.strides[0] < a.shape[1] * item * channels:
        if out:
            raise ValueError("%s is an output and must have densely packed pixels and non-overlapping rows" % name)
        a = np.ascontiguousarray(a)
    if out and not a.flags.writeable:
        raise ValueError("%s is an output and must be writeable" % name)
    return a, C.c_void_p(a.ctypes.data), a.strides[0], a.shape[0], a.shape[1]


def _same_size(what, size, **others):
    """every operand of a call must have imageL's (rows, cols): the C ABI trusts the sizes it is given"""
    for name, sz in others.items():
        if tuple(sz) != tuple(size):
            raise ValueError("%s: %s is %d x %d but imageL is %d x %d" % (what, name, sz[0], sz[1], size[0], size[1]))


class OpticalFlowInterface:
    """Abstract boundary, CPU/PixFlow.hpp:15-26"""
    DirectionHint = DirectionHint

    def computeOpticalFlow(self, I0BGRA, I1BGRA, hint, flow=None):
        raise NotImplementedError


class PixFlow(OpticalFlowInterface):
    """PixFlow<MaxPercentage> (CPU/PixFlow.hpp:28-457) running on a B200 through libpixflow_b200."""

    def __init__(self, flowAlgName, device=-1):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._pending = {}
        self.flowAlgName = flowAlgName
        _lib.check(self._lib.pf_engine_create(flowAlgName.encode(), int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pf_engine_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- OpticalFlowInterface::computeOpticalFlow, CPU/PixFlow.hpp:72-135 --
    def computeOpticalFlow(self, I0BGRA, I1BGRA, hint=DirectionHint.UNKNOWN, flow=None):
        k0, p0, s0, rows, cols = _view(I0BGRA, np.uint8, 4, "I0BGRA")
        k1, p1, s1, r1, c1 = _view(I1BGRA, np.uint8, 4, "I1BGRA")
        if (rows, cols) != (r1, c1):
            raise ValueError("I0BGRA and I1BGRA must have the same size")
        if flow is None:
            flow = np.empty((rows, cols, 2), np.float32)
        kf, pf_, sf, rf, cf = _view(flow, np.float32, 2, "flow", out=True)
        if (rf, cf) != (rows, cols):
            raise ValueError("flow must be (rows, cols, 2)")
        _lib.check(self._lib.pf_compute_flow(self._h, p0, s0, p1, s1, rows, cols, int(hint), pf_, sf))
        return kf

    # -- NovelViewGeneratorAsymmetricFlow::prepare semantics, CPU/OpticalFlow.cpp:102-145 --
    def prepareBidirectional(self, imageL, imageR, flowLtoR=None, flowRtoL=None):
        kl, pl, sl, rows, cols = _view(imageL, np.uint8, 4, "colorImageL")
        kr, pr, sr, r1, c1 = _view(imageR, np.uint8, 4, "colorImageR")
        if (rows, cols) != (r1, c1):
            raise ValueError("colorImageL and colorImageR must have the same size")
        if flowLtoR is None:
            flowLtoR = np.empty((rows, cols, 2), np.float32)
        if flowRtoL is None:
            flowRtoL = np.empty((rows, cols, 2), np.float32)
        ka, pa, sa, ra, ca = _view(flowLtoR, np.float32, 2, "flowLtoR", out=True)
        kb, pb, sb, rb, cb = _view(flowRtoL, np.float32, 2, "flowRtoL", out=True)
        if (ra, ca) != (rows, cols) or (rb, cb) != (rows, cols):
            raise ValueError("flowLtoR and flowRtoL must be (rows, cols, 2)")
        _lib.check(self._lib.pf_prepare_bidirectional(self._h, pl, sl, pr, sr, rows, cols, pa, sa, pb, sb))
        return ka, kb

    def prepareBidirectionalBatch(self, imagesL, imagesR, flowsLtoR=None, flowsRtoL=None, slot=None):
        """n independent pairs of identical size, all in flight concurrently on this engine's device."""
        n = len(imagesL)
        vl = [_view(a, np.uint8, 4, "imagesL[%d]" % i) for i, a in enumerate(imagesL)]
        vr = [_view(a, np.uint8, 4, "imagesR[%d]" % i) for i, a in enumerate(imagesR)]
        rows, cols = vl[0][3], vl[0][4]
        if flowsLtoR is None:
            flowsLtoR = [np.empty((rows, cols, 2), np.float32) for _ in range(n)]
        if flowsRtoL is None:
            flowsRtoL = [np.empty((rows, cols, 2), np.float32) for _ in range(n)]
        va = [_view(a, np.float32, 2, "flowsLtoR[%d]" % i, out=True) for i, a in enumerate(flowsLtoR)]
        vb = [_view(a, np.float32, 2, "flowsRtoL[%d]" % i, out=True) for i, a in enumerate(flowsRtoL)]
        for group in (vl, vr, va, vb):
            if len(group) != n or any(v[2] != group[0][2] or (v[3], v[4]) != (rows, cols) for v in group):
                raise ValueError("all pairs of a batch must share size and stride")
        arr = lambda vs: (C.c_void_p * n)(*[v[1].value for v in vs])
        if slot is None:
            _lib.check(self._lib.pf_prepare_bidirectional_batch(
                self._h, n, arr(vl), vl[0][2], arr(vr), vr[0][2], rows, cols, arr(va), va[0][2], arr(vb), vb[0][2]))
        else:
            _lib.check(self._lib.pf_prepare_bidirectional_batch_async(
                self._h, int(slot), n, arr(vl), vl[0][2], arr(vr), vr[0][2], rows, cols, arr(va), va[0][2], arr(vb), vb[0][2]))
            self._pending[int(slot)] = (vl, vr, va, vb)          # keep the buffers alive until wait(slot)
        return [v[0] for v in va], [v[0] for v in vb]

    def prepareBidirectionalBatchAsync(self, slot, imagesL, imagesR, flowsLtoR, flowsRtoL):
        """pf_prepare_bidirectional_batch_async: returns at once; the flows are complete after wait(slot).  slot is 0 or 1."""
        return self.prepareBidirectionalBatch(imagesL, imagesR, flowsLtoR, flowsRtoL, slot=slot)

    def wait(self, slot):
        _lib.check(self._lib.pf_wait(self._h, int(slot)))
        self._pending.pop(int(slot), None)

    def combineNovelViews(self, imageL, imageR, flowLtoR, flowRtoL, blend, out=None):
        kl, pl, sl, rows, cols = _view(imageL, np.uint8, 4, "imageL")
        kr, pr, sr, r1, c1 = _view(imageR, np.uint8, 4, "imageR")
        ka, pa, sa, r2, c2 = _view(flowLtoR, np.float32, 2, "flowLtoR")
        kb, pb, sb, r3, c3 = _view(flowRtoL, np.float32, 2, "flowRtoL")
        kc, pc, sc, r4, c4 = _view(blend, np.float32, 1, "blend")
        if out is None:
            out = np.empty((rows, cols, 4), np.uint8)
        ko, po, so, r5, c5 = _view(out, np.uint8, 4, "out", out=True)
        _same_size("combineNovelViews", (rows, cols), imageR=(r1, c1), flowLtoR=(r2, c2), flowRtoL=(r3, c3), blend=(r4, c4), out=(r5, c5))
        _lib.check(self._lib.pf_combine_novel_views(self._h, pl, sl, pr, sr, pa, sa, pb, sb, pc, sc, rows, cols, po, so))
        return ko

    def novelView(self, imageL, imageR, blend, out=None, flowLtoR=None, flowRtoL=None):
        """prepare + setBlend + generateNovelView fused on the device (CPU/main.cpp:82-89)."""
        kl, pl, sl, rows, cols = _view(imageL, np.uint8, 4, "imageL")
        kr, pr, sr, r1, c1 = _view(imageR, np.uint8, 4, "imageR")
        kc, pc, sc, r2, c2 = _view(blend, np.float32, 1, "blend")
        if out is None:
            out = np.empty((rows, cols, 4), np.uint8)
        ko, po, so, r3, c3 = _view(out, np.uint8, 4, "out", out=True)
        sizes = dict(imageR=(r1, c1), blend=(r2, c2), out=(r3, c3))
        pa = pb = C.c_void_p()
        sa = sb = 0
        if flowLtoR is not None:
            ka, pa, sa, r4, c4 = _view(flowLtoR, np.float32, 2, "flowLtoR", out=True)
            sizes["flowLtoR"] = (r4, c4)
        if flowRtoL is not None:
            kb, pb, sb, r5, c5 = _view(flowRtoL, np.float32, 2, "flowRtoL", out=True)
            sizes["flowRtoL"] = (r5, c5)
        _same_size("novelView", (rows, cols), **sizes)
        _lib.check(self._lib.pf_novel_view(self._h, pl, sl, pr, sr, pc, sc, rows, cols, po, so, pa, sa, pb, sb))
        return ko

    # -- instrumentation used by bench.py --
    def setSweepTiming(self, enabled):
        _lib.check(self._lib.pf_set_sweep_timing(self._h, int(bool(enabled))))

    def timerStart(self):
        _lib.check(self._lib.pf_timer_start(self._h))

    def timerStop(self):
        ms = C.c_double()
        _lib.check(self._lib.pf_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def lastSweepMs(self):
        return float(self._lib.pf_last_sweep_ms(self._h)), int(self._lib.pf_last_sweep_launches(self._h))


def makeOpticalFlowByName(flowAlgName, device=-1):
    """CPU/PixFlow.hpp:459-500: "pixflow_low" -> PixFlow<0>, "pixflow_search_20" -> PixFlow<20>;
    anything else raises (reference: throw VrCamException("unrecognized flow algorithm name: ..."))."""
    return PixFlow(flowAlgName, device)


class NovelViewUtil:
    """CPU/OpticalFlow.hpp:19-32"""

    @staticmethod
    def combineNovelViews(imageL, imageR, flowLtoR, flowRtoL, blend, flowAlg=None):
        own = flowAlg is None
        alg = makeOpticalFlowByName("pixflow_low") if own else flowAlg
        try:
            return alg.combineNovelViews(imageL, imageR, flowLtoR, flowRtoL, blend)
        finally:
            if own:
                alg.close()


class NovelViewGenerator:
    """CPU/OpticalFlow.hpp:34-48"""

    def prepare(self, colorImageL, colorImageR):
        raise NotImplementedError

    def generateNovelView(self):
        raise NotImplementedError

    def getFlowLtoR(self):
        return None

    def getFlowRtoL(self):
        return None

    def setBlend(self, blend):
        raise NotImplementedError


class NovelViewGeneratorAsymmetricFlow(NovelViewGenerator):
    """CPU/OpticalFlow.hpp:50-70, CPU/OpticalFlow.cpp:94-145"""

    def __init__(self, flowAlgName, device=-1):
        self.flowAlgName = flowAlgName
        self._alg = makeOpticalFlowByName(flowAlgName, device)   # raises on an unknown name, like prepare() would
        self.imageL = self.imageR = None
        self.flowLtoR = self.flowRtoL = None
        self.Blend = None

    def prepare(self, colorImageL, colorImageR):
        self.imageL = np.array(colorImageL, copy=True)   # .clone(), CPU/OpticalFlow.cpp:106-107
        self.imageR = np.array(colorImageR, copy=True)
        self.flowLtoR, self.flowRtoL = self._alg.prepareBidirectional(self.imageL, self.imageR)

    def setBlend(self, blend):
        self.Blend = np.array(blend, dtype=np.float32, copy=True)   # blend.clone(), CPU/OpticalFlow.hpp:69

    def generateNovelView(self):
        return self._alg.combineNovelViews(self.imageL, self.imageR, self.flowLtoR, self.flowRtoL, self.Blend)

    def getFlowLtoR(self):
        return self.flowLtoR

    def getFlowRtoL(self):
        return self.flowRtoL

    def close(self):
        self._alg.close()


class Stitchtools:
    """Stitchtools (CPU/StitchTool.hpp:21-61) on the B200: prepare() (MatchImages, overlap masking, GenerateBlend with
    countblend and the blend smoothing), setMergedmiddle() + Gather(), and the getters of the reference.  Member names are
    the reference's (ImageL, ImageR, Blend, OverlappedL, OverlappedR, Mergedmiddle, Map, FinalResult, MergedDis)."""

    def __init__(self, flowAlg=None, device=-1):
        self._own = flowAlg is None
        self._alg = makeOpticalFlowByName("pixflow_low", device) if flowAlg is None else flowAlg
        self.ImageL = self.ImageR = None
        self.Map = self.OverlappedL = self.OverlappedR = self.Blend = self.BlendUnsmoothed = self.MergedDis = None
        self.Mergedmiddle = self.FinalResult = None

    def prepare(self, colorImageL, colorImageR):
        """CPU/StitchTool.cpp:7-36.  Raises PixFlowError for sizes the reference itself cannot run (rows < 400 or a shorter
        side < 200)."""
        self.ImageL = np.array(colorImageL, copy=True)
        self.ImageR = np.array(colorImageR, copy=True)
        kl, pl, sl, rows, cols = _view(self.ImageL, np.uint8, 4, "colorImageL")
        kr, pr, sr, r1, c1 = _view(self.ImageR, np.uint8, 4, "colorImageR")
        if (rows, cols) != (r1, c1):
            raise ValueError("colorImageL and colorImageR must have the same size")
        self.Map = np.empty((rows, cols), np.uint8)
        self.OverlappedL = np.empty((rows, cols, 4), np.uint8)
        self.OverlappedR = np.empty((rows, cols, 4), np.uint8)
        self.BlendUnsmoothed = np.empty((rows, cols), np.float32)
        self.MergedDis = np.empty((rows, cols), np.float32)
        self.Blend = np.empty((rows, cols), np.float32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(self._alg._lib.pf_stitch_prepare(
            self._alg._h, pl, sl, pr, sr, rows, cols, p(self.Map), cols, p(self.OverlappedL), cols * 4,
            p(self.OverlappedR), cols * 4, p(self.BlendUnsmoothed), cols * 4, p(self.MergedDis), cols * 4,
            p(self.Blend), cols * 4))

    def setMergedmiddle(self, image):
        self.Mergedmiddle = np.array(image, copy=True)

    def Gather(self):
        """CPU/StitchTool.cpp:52-96 -> FinalResult"""
        if self.Map is None or self.Mergedmiddle is None:
            raise RuntimeError("Gather() needs prepare() and setMergedmiddle() first")
        rows, cols = self.Map.shape
        kl, pl, sl, _, _ = _view(self.ImageL, np.uint8, 4, "ImageL")
        kr, pr, sr, _, _ = _view(self.ImageR, np.uint8, 4, "ImageR")
        km, pm, sm, r1, c1 = _view(self.Mergedmiddle, np.uint8, 4, "Mergedmiddle")
        if (rows, cols) != (r1, c1):
            raise ValueError("Mergedmiddle must have the size of the canvas")
        self.FinalResult = np.empty((rows, cols, 4), np.uint8)
        _lib.check(self._alg._lib.pf_stitch_gather(
            self._alg._h, pl, sl, pr, sr, pm, sm, C.c_void_p(self.Map.ctypes.data), cols, rows, cols,
            C.c_void_p(self.FinalResult.ctypes.data), cols * 4))

    def getImageL(self):
        return self.ImageL

    def getImageR(self):
        return self.ImageR

    def getBlend(self):
        return self.Blend

    def getMap(self):
        return self.Map

    def getOverlappedL(self):
        return self.OverlappedL

    def getOverlappedR(self):
        return self.OverlappedR

    def getFinalResult(self):
        return self.FinalResult

    def getBlendUnsmoothed(self):
        """GenerateBlend's map before its smoothing (CPU/StitchTool.cpp:113-124); not a member of the reference class"""
        return self.BlendUnsmoothed

    def close(self):
        if self._own:
            self._alg.close()


def stitch_iteration(flowAlg, colorImageL, colorImageR, out=None, want_intermediates=False):
    """One pass of the loop body of the reference driver (CPU/main.cpp:72-95) as ONE call, every intermediate resident in
    HBM: Stitchtools::prepare -> NovelViewGeneratorAsymmetricFlow::prepare -> setBlend -> generateNovelView ->
    setMergedmiddle -> Gather.  Images may be numpy arrays or CUDA tensors; `out` may be a CUDA uint8 tensor (rows, cols, 4)
    that is fed back as colorImageR of the next iteration.  Returns FinalResult (and a dict of Blend / Mergedmiddle / Map)."""
    kl, pl, sl, rows, cols = _view(colorImageL, np.uint8, 4, "colorImageL")
    kr, pr, sr, r1, c1 = _view(colorImageR, np.uint8, 4, "colorImageR")
    if (rows, cols) != (r1, c1):
        raise ValueError("colorImageL and colorImageR must have the same size")
    if out is None:
        out = np.empty((rows, cols, 4), np.uint8)
    ko, po, so, r2, c2 = _view(out, np.uint8, 4, "out", out=True)
    if (rows, cols) != (r2, c2) or (not _is_device_tensor(out) and ko is not out):
        raise ValueError("out must be a dense (rows, cols, 4) uint8 buffer of the canvas size")
    null = C.c_void_p(None)
    extra = {}
    args = [null, 0, null, 0, null, 0]
    if want_intermediates:
        extra = dict(Blend=np.empty((rows, cols), np.float32), Mergedmiddle=np.empty((rows, cols, 4), np.uint8),
                     Map=np.empty((rows, cols), np.uint8))
        args = [C.c_void_p(extra["Blend"].ctypes.data), cols * 4, C.c_void_p(extra["Mergedmiddle"].ctypes.data), cols * 4,
                C.c_void_p(extra["Map"].ctypes.data), cols]
    _lib.check(flowAlg._lib.pf_stitch_iteration(flowAlg._h, pl, sl, pr, sr, rows, cols, po, so, *args))
    return (out, extra) if want_intermediates else out


def four_input_frontend(flowAlg, colorImage1, colorImage2, colorImage3, colorImage4, device_out=False):
    """The input preparation of the 4-input driver (CPU_4Input/main.cpp:64-79) -> (colorImageL, colorImageR).
    device_out=True (CUDA tensor inputs): the two outputs are CUDA tensors and nothing crosses PCIe."""
    views = [_view(a, np.uint8, 4, "colorImage%d" % (k + 1)) for k, a in enumerate((colorImage1, colorImage2, colorImage3, colorImage4))]
    rows, cols = views[0][3], views[0][4]
    if any((v[3], v[4]) != (rows, cols) for v in views) or any(v[2] != views[0][2] for v in views):
        raise ValueError("the four inputs must have the same size and row stride")
    ptrs = (C.c_void_p * 4)(*[v[1] for v in views])
    if device_out:
        import torch
        L = torch.empty((rows, cols, 4), dtype=torch.uint8, device=colorImage1.device)
        R = torch.empty_like(L)
        pl, pr = C.c_void_p(L.data_ptr()), C.c_void_p(R.data_ptr())
    else:
        L = np.empty((rows, cols, 4), np.uint8)
        R = np.empty((rows, cols, 4), np.uint8)
        pl, pr = C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data)
    _lib.check(flowAlg._lib.pf_four_input_frontend(flowAlg._h, ptrs, views[0][2], rows, cols, pl, cols * 4, pr, cols * 4))
    return L, R


def _blend_smooth_for_tests(flowAlg, blend_raw, merged_dis):
    """The smoothing of GenerateBlend (CPU/StitchTool.cpp:133-145) alone, on caller-supplied values (diagnostic entry point
    pf_stage_blend_smooth; used by the parity tests to drive the box filters with adversarial data)."""
    b = np.array(blend_raw, np.float32, order="C", copy=True)
    md = np.ascontiguousarray(merged_dis, np.float32)
    rows, cols = b.shape
    _lib.check(flowAlg._lib.pf_stage_blend_smooth(flowAlg._h, C.c_void_p(b.ctypes.data), C.c_void_p(md.ctypes.data), rows, cols))
    return b
