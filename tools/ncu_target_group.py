"""ncu target: MOG2 stream-group kernel (mog2_t1_kernel<.,0,true>) on S x 1080p streams, steady state.
usage: ncu --launch-skip 60 -c 2 -k regex:mog2_t1 ... python tools/ncu_target_group.py [S]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H, NF = 1920, 1080, 16
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF, S, H, W, 3), dtype=torch.uint8, device="cuda")
for t in range(NF):
    synth.frames_dev(d[t].data_ptr(), S, 1, W, H, t0=t, stream=st)
fg = torch.empty((S, H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((S, H, W, 3), dtype=torch.uint8, device="cuda")
p = tb.MixtureOfGaussianV2BGS(nstreams=S)
for k in range(66):
    p.process_dev(d[k % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st)
torch.cuda.synchronize()
print("done")
