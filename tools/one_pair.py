"""One bidirectional flow of a synthetic rows x cols pair through the public API (device-resident), for ncu launch lists:
    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file out.csv python tools/one_pair.py 4000 2000
"""
import os
import sys

os.environ.setdefault("PF_NO_GRAPHS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import panorama_opticalflow_b200 as pf  # noqa: E402
from panorama_opticalflow_b200 import synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L, R = synth.make_pair(rows, cols, seed=1, amplitude=cols / 12.0 + 1.0)
dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
oLR = torch.empty((rows, cols, 2), dtype=torch.float32, device="cuda")
oRL = torch.empty_like(oLR)
eng = pf.makeOpticalFlowByName("pixflow_search_20")
for _ in range(reps):
    eng.prepareBidirectionalBatch([dL], [dR], [oLR], [oRL])
torch.cuda.synchronize()
print("done", float(oLR.abs().mean()))
eng.close()
