"""Build recipe of libpixflow_b200.so (sm_100a only, in-tree so that it travels with the repo snapshot).

    python -m panorama_opticalflow_b200.build [--force]

-fmad=false is part of the arithmetic contract (bit-compatibility with the reference CPU path), not a
tuning flag; --use_fast_math must never be added.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpixflow_b200.so")
SOURCES = ["pf_kernels.cu", "pf_sweep.cu", "pf_fused.cu", "pf_stitch.cu", "pf_selftest.cu", "pf_engine.cu"]
HEADERS = ["pf_kernels.cuh", "pf_math.cuh", "pf_prep.cuh", os.path.join("..", "..", "include", "pixflow_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
    "--shared", "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, out=None):
    """out: build a variant (with PF_EXTRA_NVCC_FLAGS) to another path, for experiments (select it with PF_LIB_PATH)"""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("PF_EXTRA_NVCC_FLAGS", "").split() + ["-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return out or LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
