"""Host logic, ABI surface and error behaviour -- no GPU needed."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from panorama_opticalflow_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pixflow_b200.h")).read()
    declared = set(re.findall(r"PF_API\s+[\w\s\*]+?\b(pf_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.lib_path()], text=True)
    exported = set(re.findall(r" T (pf_\w+)", out))
    assert declared <= exported, declared - exported


def test_library_is_sm100a_only():
    from panorama_opticalflow_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_unknown_algorithm_name_raises_like_the_reference():
    import panorama_opticalflow_b200 as pf
    with pytest.raises(pf.PixFlowError) as ei:
        pf.makeOpticalFlowByName("pixflow_medium")
    assert ei.value.code == 2 and "unrecognized flow algorithm name: pixflow_medium" in str(ei.value)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import panorama_opticalflow_b200 as pf
    with pytest.raises(pf.PixFlowError) as ei:
        pf.makeOpticalFlowByName("pixflow_low")
    assert ei.value.code == 4


def test_product_code_never_touches_the_oracle():
    """The product path must not import, load, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "panorama_opticalflow_b200")
    pat = re.compile(r"(from\s+oracle|import\s+oracle|liborc|\borc_\w+\s*\(|oracle/|oracle\.)")
    files = [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    for dirpath, _, names in os.walk(pkg):
        files += [os.path.join(dirpath, fn) for fn in names if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    for path in files:
        m = pat.search(open(path).read())
        assert m is None, "%s references the oracle: %r" % (path, m.group(0))


def test_gaussian_constants_in_kernels_match_oracle(orc):
    txt = open(os.path.join(ROOT, "panorama_opticalflow_b200", "csrc", "pf_math.cuh")).read()
    for name, (k, s) in {"kG5": (5, 0.25), "kG3H": (3, 0.5), "kG3O": (3, 1.0), "kG15": (15, 8.0)}.items():
        m = re.search(name + r"\[\d+\]\s*=\s*\{([^}]*)\}", txt)
        vals = [float.fromhex(v.strip().rstrip("f")) for v in m.group(1).replace("\\", "").split(",")]
        want = orc.gaussian_kernel(k, s)[k // 2:]
        assert np.array_equal(np.array(vals, np.float32), want), name
        assert all(np.float32(v) == v for v in vals)
    inv255 = float.fromhex(re.search(r"PF_INV255\s+(\S+?)f\s", txt).group(1))
    assert np.float32(inv255) == np.float32(1.0 / 255.0)
    invpyr = float.fromhex(re.search(r"PF_INV_PYR\s+(\S+?)f\s", txt).group(1))
    assert np.float32(invpyr) == np.float32(1.0) / np.float32(0.9)


def test_median_network_selects_the_median():
    """0-1 principle on a random subset plus structured cases (the exhaustive 2^25 check ran at development
    time)."""
    txt = open(os.path.join(ROOT, "panorama_opticalflow_b200", "csrc", "pf_math.cuh")).read()
    body = txt[txt.index("float median25"):txt.index("return v[12]")]
    net = [(int(a), int(b)) for a, b in re.findall(r"PF_CSWAP\((\d+),(\d+)\)", body)]
    assert len(net) == 99
    rng = np.random.default_rng(0)
    v = rng.integers(0, 2, (200000, 25)).astype(np.float32)
    v = np.concatenate([v, rng.standard_normal((50000, 25)).astype(np.float32), np.tril(np.ones((25, 25), np.float32))])
    want = np.sort(v, axis=1)[:, 12]
    for a, b in net:
        lo, hi = np.minimum(v[:, a], v[:, b]), np.maximum(v[:, a], v[:, b])
        v[:, a], v[:, b] = lo, hi
    assert np.array_equal(v[:, 12], want)


def test_median_network_with_three_element_sorts():
    """median25_s3: the same network with its bubble triples replaced by 3-input sorts (min3 / max3 / xor) -- must select the
    median as well, and must be the 99-comparator network regrouped (every triple expands to (j,k),(i,k),(i,j))."""
    txt = open(os.path.join(ROOT, "panorama_opticalflow_b200", "csrc", "pf_math.cuh")).read()
    base = txt[txt.index("float median25("):txt.index("return v[12]")]
    net = [(int(a), int(b)) for a, b in re.findall(r"PF_CSWAP\((\d+),(\d+)\)", base)]
    body = txt[txt.index("float median25_s3"):]
    body = body[:body.index("return v[12]")]
    ops = re.findall(r"PF_(CSWAP|SORT3)\(([\d,]+)\)", body)
    expanded = []
    for kind, args in ops:
        idx = [int(t) for t in args.split(",")]
        if kind == "CSWAP":
            expanded.append((idx[0], idx[1]))
        else:
            i, j, k = idx
            assert i < j < k
            expanded += [(j, k), (i, k), (i, j)]
    assert expanded == net and sum(1 for k, _ in ops if k == "SORT3") == 22
    rng = np.random.default_rng(1)
    v = rng.integers(0, 2, (100000, 25)).astype(np.float32)
    v = np.concatenate([v, rng.standard_normal((50000, 25)).astype(np.float32), np.tril(np.ones((25, 25), np.float32))])
    want = np.sort(v, axis=1)[:, 12]
    for kind, args in ops:
        idx = [int(t) for t in args.split(",")]
        v[:, idx] = np.sort(v[:, idx], axis=1)
    assert np.array_equal(v[:, 12], want)


def test_median_networks_exhaustively_by_the_zero_one_principle():
    """A comparator network selects the median of every input iff it does so for all 2^25 zero-one inputs.  Bit-parallel:
    each wire is a 2^25-bit vector (input number i has bit w of i on wire w), min = AND, max = OR, and the middle of a
    3-input sort a ^ b ^ c ^ lo ^ hi is the majority.  Checks both median25 (2-input) and median25_s3 (3-input sorts)."""
    txt = open(os.path.join(ROOT, "panorama_opticalflow_b200", "csrc", "pf_math.cuh")).read()
    n = 1 << 25
    idx = np.arange(n, dtype=np.uint32)
    wires0 = [np.packbits(((idx >> w) & 1).astype(np.uint8)) for w in range(25)]
    count = np.zeros(n, np.uint8)
    for w in range(25):
        count += ((idx >> w) & 1).astype(np.uint8)
    want = np.packbits((count >= 13).astype(np.uint8))      # the 13th smallest of 25 zero-one values is 1 iff >= 13 ones
    del idx, count
    for name in ("float median25(", "float median25_s3("):
        body = txt[txt.index(name):]
        body = body[:body.index("return v[12]")]
        v = [w.copy() for w in wires0]
        for kind, args in re.findall(r"PF_(CSWAP|SORT3)\(([\d,]+)\)", body):
            t = [int(q) for q in args.split(",")]
            if kind == "CSWAP":
                a, b = t
                v[a], v[b] = v[a] & v[b], v[a] | v[b]
            else:
                a, b, c = t
                lo, hi = v[a] & v[b] & v[c], v[a] | v[b] | v[c]
                v[b] = v[a] ^ v[b] ^ v[c] ^ lo ^ hi
                v[a], v[c] = lo, hi
        assert np.array_equal(v[12], want), name


def test_pyramid_plan_matches_survey(orc):
    # config 2 through prepare -> level-0 1100x2000, 37 levels, coarsest 25x45
    rows, cols = 4000, 2000
    pc = cols + 2 * (cols // 20)
    dh, dw = orc.downscale_size(rows, pc)
    sizes = orc.pyramid_sizes(dw, dh)
    assert (dw, dh) == (1100, 2000) and len(sizes) == 37 and sizes[-1] == (25, 45)
    assert sum(w + h - 1 for w, h in sizes) == 30363
    # config 1: 512^2 direct -> 256^2, 23 levels, coarsest 26^2
    sizes = orc.pyramid_sizes(256, 256)
    assert len(sizes) == 23 and sizes[-1] == (26, 26)
    assert orc.search_distance(20) == 5 and orc.search_distance(0) == 0


def test_argument_validation_before_any_device_work():
    from panorama_opticalflow_b200 import _lib
    lib = _lib.load()
    assert lib.pf_compute_flow(None, None, 0, None, 0, 4, 4, 0, None, 0) == _lib.PF_ERR_INVALID_ARGUMENT
    assert b"engine is NULL" in lib.pf_last_error()
    h = C.c_void_p()
    assert lib.pf_engine_create(None, 0, C.byref(h)) == _lib.PF_ERR_INVALID_ARGUMENT
