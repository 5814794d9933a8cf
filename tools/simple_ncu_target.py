"""16 x 1080p streams through FD, ABL, WMV and ASBL (steady state): target of `ncu -k regex:<kernel>` captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
S, w, h, NT = 16, 1920, 1080, 4
st = torch.cuda.current_stream().cuda_stream
frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
for t in range(NT):
    synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
fg = torch.empty((S, h, w), dtype=torch.uint8, device="cuda")
bg = torch.empty((S, h, w, 3), dtype=torch.uint8, device="cuda")
for cls in (tb.FrameDifferenceBGS, tb.AdaptiveBackgroundLearning, tb.WeightedMovingVarianceBGS,
            tb.AdaptiveSelectiveBackgroundLearning):
    p = cls(nstreams=S)
    for k in range(6):
        p.process_dev(frames[k % NT].data_ptr(), w, h, fg.data_ptr(), bg.data_ptr(), stream=st)
    torch.cuda.synchronize()
    p.close()
print("ok")
