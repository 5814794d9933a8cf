"""Where the MOG2 frame time goes: production kernel vs two timing instruments on the same warmed 1080p model
(kernelVariant 9: same loads/stores, no arithmetic; 8: no generic phase).  GPU box, measurement tooling.
Needs an instrumented library: `python -m tracking_b200._build --force --instrument` first (the shipped build rejects 8 / 9)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
W, H, NF = 1920, 1080, 64
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF, W, H, stream=st)
fg = torch.empty((H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
p = tb.MixtureOfGaussianV2BGS()
k = [0]
def run(n):
    for _ in range(n):
        p.process_dev(d[k[0] % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st); k[0] += 1
run(128)
def timed(n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(n); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("production us/frame %.2f" % timed(256))
p.set("kernelVariant", 1); run(16)
print("straight restatement kernel us/frame %.2f" % timed(128))
p.set("kernelVariant", 0); run(16)
print("production again us/frame %.2f" % timed(256))
p.set("kernelVariant", 9)
print("no-arithmetic floor (same loads/stores on the warmed model) us/frame %.2f" % timed(64))
p.set("kernelVariant", 0); run(64)
p.set("kernelVariant", 8)
print("without the generic phase (ineligible pixels left untouched) us/frame %.2f" % timed(64))
p.set("kernelVariant", 0)
