"""ctypes loader of libpixflow_b200.so (the C-ABI declared in include/pixflow_b200.h).

Fails loudly: if the shared library is missing it is built with nvcc; if that is impossible an
ImportError/RuntimeError propagates.  There is no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os

from . import build as _build

PF_OK, PF_ERR_INVALID_ARGUMENT, PF_ERR_UNKNOWN_ALGORITHM, PF_ERR_CUDA, PF_ERR_NO_DEVICE = range(5)

_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); keep in sync with include/pixflow_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "pf_engine_create": (_i, [C.c_char_p, _i, C.POINTER(_vp)]),
    "pf_engine_destroy": (None, [_vp]),
    "pf_compute_flow": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _i, _vp, _sz]),
    "pf_prepare_bidirectional": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz, _vp, _sz]),
    "pf_prepare_bidirectional_batch": (_i, [_vp, _i, C.POINTER(_vp), _sz, C.POINTER(_vp), _sz, _i, _i,
                                            C.POINTER(_vp), _sz, C.POINTER(_vp), _sz]),
    "pf_prepare_bidirectional_batch_async": (_i, [_vp, _i, _i, C.POINTER(_vp), _sz, C.POINTER(_vp), _sz, _i, _i,
                                                  C.POINTER(_vp), _sz, C.POINTER(_vp), _sz]),
    "pf_wait": (_i, [_vp, _i]),
    "pf_combine_novel_views": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "pf_novel_view": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz, _vp, _sz, _vp, _sz]),
    "pf_stitch_prepare": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "pf_stitch_gather": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "pf_four_input_frontend": (_i, [_vp, C.POINTER(_vp), _sz, _i, _i, _vp, _sz, _vp, _sz]),
    "pf_stitch_iteration": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "pf_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "pf_host_free": (_i, [_vp]),
    "pf_kernel_launch_count": (C.c_uint64, []),
    "pf_set_sweep_timing": (_i, [_vp, _i]),
    "pf_last_sweep_ms": (C.c_double, [_vp]),
    "pf_last_sweep_launches": (C.c_uint64, [_vp]),
    "pf_timer_start": (_i, [_vp]),
    "pf_timer_stop": (_i, [_vp, C.POINTER(C.c_double)]),
    "pf_selftest_exact_math": (_i, [_i, _i, C.POINTER(C.c_uint64)]),
    "pf_last_error": (C.c_char_p, []),
    "pf_version": (C.c_char_p, []),
    "pf_stage_frontend": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i]),
    "pf_stage_gauss5": (_i, [_vp, _vp, _i, _i]),
    "pf_stage_pyr_down": (_i, [_vp, _i, _i, _vp, _i, _i]),
    "pf_stage_gradient": (_i, [_vp, _vp, _i, _i]),
    "pf_stage_blur15": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "pf_stage_median5": (_i, [_vp, _vp, _i, _i]),
    "pf_stage_sweep": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "pf_stage_upsample_cubic": (_i, [_vp, _i, _i, _vp, _i, _i]),
    "pf_stage_tail": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pf_stage_blend_smooth": (_i, [_vp, _vp, _vp, _i, _i]),
    "pf_stage_initial_flow": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) and return the ctypes handle with typed signatures."""
    global _lib
    if _lib is None:
        path = os.environ.get("PF_LIB_PATH") or _build.build_lib()      # PF_LIB_PATH: a variant build, for experiments only
        if not os.path.exists(path):
            raise ImportError("libpixflow_b200.so is missing and could not be built: " + path)
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class PixFlowError(RuntimeError):
    """Python face of a non-zero pf_status (the reference throws VrCamException / aborts via glog)."""

    def __init__(self, code, msg):
        super().__init__("pixflow_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != PF_OK:
        raise PixFlowError(rc, load().pf_last_error().decode("utf-8", "replace"))
