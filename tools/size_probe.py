"""MOG2 T=1 time per frame vs frame size (does the model state fit the L2?).  GPU box, measurement tooling."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
st = torch.cuda.current_stream().cuda_stream
for (W, H) in ((960, 540), (1280, 720), (1600, 900), (1920, 1080), (2560, 1440), (3840, 2160)):
    NF = 32
    d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), 1, NF, W, H, stream=st)
    fg = torch.empty((H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    p = tb.MixtureOfGaussianV2BGS()
    k = [0]
    def run(n):
        for _ in range(n):
            p.process_dev(d[k[0] % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st); k[0] += 1
    run(128)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(256); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 256 * 1e3
    print("%dx%d  %.2f us/frame  %.2f ns/kpx  %.1f Gpx/s  (slot-0 state %.0f MB)" % (W, H, us, us * 1e3 / (W * H / 1e3), W * H / us / 1e3, W * H * 20 / 1e6))
    p.close(); del d
