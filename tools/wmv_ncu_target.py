"""16 x 1080p streams through WMV with retainInput (steady state): target of `ncu -k regex:wmv_kernel` captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
S, w, h, NT = 16, 1920, 1080, 6
st = torch.cuda.current_stream().cuda_stream
frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
for t in range(NT):
    synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
fg = torch.empty((S, h, w), dtype=torch.uint8, device="cuda")
p = tb.WeightedMovingVarianceBGS(nstreams=S, retainInput=1)
for k in range(8):
    p.process_dev(frames[k % NT].data_ptr(), w, h, fg.data_ptr(), None, stream=st)
torch.cuda.synchronize()
p.close()
print("ok")
