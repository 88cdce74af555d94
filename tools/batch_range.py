"""One batch step (B pairs in flight) between cudaProfilerStart/Stop, for `ncu --replay-mode range`: hardware counters over
the WHOLE concurrent step (what no per-kernel capture can show, because ncu serialises kernels).

    ncu --replay-mode range --metrics <list> python tools/batch_range.py [B]
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import panorama_opticalflow_b200 as pf  # noqa: E402
from panorama_opticalflow_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rows, cols = 4000, 2000
L, R = synth.make_pair(rows, cols, seed=1, amplitude=cols / 12.0 + 1.0)
base = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
dL = [torch.roll(base[0], 37 * i, dims=0).contiguous() for i in range(B)]
dR = [torch.roll(base[1], 37 * i, dims=0).contiguous() for i in range(B)]
oLR = [torch.empty((rows, cols, 2), dtype=torch.float32, device="cuda") for _ in range(B)]
oRL = [torch.empty((rows, cols, 2), dtype=torch.float32, device="cuda") for _ in range(B)]
eng = pf.makeOpticalFlowByName("pixflow_search_20")
for _ in range(2):
    eng.prepareBidirectionalBatch(dL, dR, oLR, oRL)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.prepareBidirectionalBatch(dL, dR, oLR, oRL)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
eng.close()
