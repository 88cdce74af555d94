"""Host-side argument handling of the Python mirror (no GPU needed): outputs must be usable in place, operand sizes must agree."""
import numpy as np
import pytest

from panorama_opticalflow_b200 import api


def test_view_copies_non_dense_inputs_but_rejects_non_dense_outputs():
    a = np.zeros((10, 12, 8), np.uint8)[:, :, ::2]            # pixel stride 8 bytes: not densely packed
    keep, ptr, stride, rows, cols = api._view(a, np.uint8, 4, "image")
    assert keep is not a and keep.flags.c_contiguous and (rows, cols) == (10, 12)
    with pytest.raises(ValueError, match="output"):
        api._view(a, np.uint8, 4, "flow", out=True)
    ro = np.zeros((4, 4, 2), np.float32)
    ro.setflags(write=False)
    with pytest.raises(ValueError, match="writeable"):
        api._view(ro, np.float32, 2, "flow", out=True)


def test_view_keeps_row_strided_outputs_in_place():
    big = np.zeros((10, 20, 2), np.float32)
    v = big[:, :12]
    keep, ptr, stride, rows, cols = api._view(v, np.float32, 2, "flow", out=True)
    assert keep is v and stride == 20 * 8 and (rows, cols) == (10, 12)


def test_view_type_and_shape_errors():
    with pytest.raises(TypeError):
        api._view(np.zeros((4, 4, 4), np.float32), np.uint8, 4, "image")
    with pytest.raises(ValueError):
        api._view(np.zeros((4, 4, 3), np.uint8), np.uint8, 4, "image")
    with pytest.raises(ValueError):
        api._view(np.zeros((4, 4, 1), np.float32), np.float32, 1, "blend")


def test_same_size_check_names_the_offending_operand():
    api._same_size("combineNovelViews", (4, 5), imageR=(4, 5), blend=(4, 5))
    with pytest.raises(ValueError, match="blend is 4 x 6"):
        api._same_size("combineNovelViews", (4, 5), imageR=(4, 5), blend=(4, 6))
