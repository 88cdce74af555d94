"""One stitching iteration of BASELINE config 3 (top.tif + 1.tif from data/Test_data_1, 4000 x 8998 canvas) through the public
API, for an ncu launch list of the whole iteration (PF_NO_GRAPHS=1 so that every kernel is a launch):
    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file out.csv python tools/stitch_real_profile.py
"""
import os
import sys

os.environ.setdefault("PF_NO_GRAPHS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import panorama_opticalflow_b200 as pf  # noqa: E402
from panorama_opticalflow_b200 import testdata  # noqa: E402

it = int(sys.argv[1]) if len(sys.argv) > 1 else 1
R = torch.from_numpy(testdata.load("Test_data_1", "top")).cuda()
eng = pf.makeOpticalFlowByName("pixflow_search_20")
out = torch.empty_like(R)
for i in range(1, it + 1):
    L = torch.from_numpy(testdata.load("Test_data_1", str(i))).cuda()
    torch.cuda.synchronize()
    pf.stitch_iteration(eng, L, R, out=out)
    R, out = out, torch.empty_like(out)
torch.cuda.synchronize()
print("done", int(R.sum()))
eng.close()
