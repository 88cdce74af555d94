import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure), built on demand with gcc."""
    from oracle import orc as _orc
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def engine_low():
    import panorama_opticalflow_b200 as pf
    e = pf.makeOpticalFlowByName("pixflow_low")
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_search():
    import panorama_opticalflow_b200 as pf
    e = pf.makeOpticalFlowByName("pixflow_search_20")
    yield e
    e.close()


def assert_bit_equal(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, "%s shape %s vs %s" % (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
        nbad = int((got != want).sum())
        idx = np.unravel_index(np.argmax(diff), diff.shape)
        raise AssertionError("%s: %d/%d elements differ, max abs diff %.9g at %s (got %r want %r)"
                             % (what, nbad, got.size, diff.max(), idx, got[idx], want[idx]))
