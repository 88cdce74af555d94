"""Exhaustive check (every float in the validity range) that the branch-free sqrt / division-by-constant used on
the sweep's critical path return exactly the IEEE correctly rounded results."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_exact_math_exhaustive():
    from panorama_opticalflow_b200 import _lib
    out = (C.c_uint64 * 4)()
    # every level width the fast path is used for (pf_math.cuh: PF_EXACT_W_MIN .. PF_EXACT_W_MAX; BASELINE config 3's level 0
    # is 4948 wide); widths outside take the IEEE-intrinsic path (tests/test_gpu_pipeline.py::test_tiny_images_vs_oracle)
    _lib.check(_lib.load().pf_selftest_exact_math(24, 8192, out))
    assert out[0] == 0, "sqrt_exact_fast differs from __fsqrt_rn on %d inputs" % out[0]
    assert out[1] == 0, "x/0.001f differs on %d inputs" % out[1]
    assert out[2] == 0, "x/float(w) differs on %d inputs, first bad w = %d" % (out[2], out[3])
