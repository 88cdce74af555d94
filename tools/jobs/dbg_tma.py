import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
from panorama_opticalflow_b200 import stages
from oracle import orc
rng = np.random.default_rng(0)
for name, shape in (("median", (64, 200)), ("blur", (150, 200))):
    f = rng.standard_normal(shape + (2,)).astype(np.float32)
    try:
        if name == "median":
            got = stages.median5(f); want = orc.median5_c2(f)
        else:
            got = stages.blur15(f); want = orc.gaussian_blur(f, 15, 8.0)
        print(name, "ok" if np.array_equal(got, want) else "MISMATCH %d" % int((got != want).sum()), flush=True)
    except Exception as e:
        print(name, "ERROR", e, flush=True)
        break
