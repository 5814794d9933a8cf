"""Warm a 1080p MOG2 model for 128 frames, then run a few more: the target of ncu captures
(usage: ncu --launch-skip 129 -c 1 ... python tools/ncu_target.py [kernelVariant])."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
W, H, NF = 1920, 1080, 64
v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF, W, H, stream=st)
fg = torch.empty((H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
p = tb.MixtureOfGaussianV2BGS()
p.set("kernelVariant", 0)
for k in range(128):
    p.process_dev(d[k % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st)
p.set("kernelVariant", v)
for k in range(128, 136):
    p.process_dev(d[k % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st)
torch.cuda.synchronize()
print("done")
