"""DPPratiMediod on 8 x 1080p streams: the subtract pass alone (samplingRate 1000: the ring holds one sample), the update
pass on every frame (samplingRate 1, ring of 16 full), and the default mix.  GPU box, measurement tooling."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
S, w, h, NT = 8, 1920, 1080, 6
st = torch.cuda.current_stream().cuda_stream
frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
for t in range(NT):
    synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
fg = torch.empty((S, h, w), dtype=torch.uint8, device="cuda")
for name, kw, warm in (("subtract only", dict(samplingRate=1000), 20), ("update every frame", dict(samplingRate=1), 20), ("default", {}, 80)):
    p = tb.DPPratiMediodBGS(nstreams=S, **kw)
    k = [0]
    def run(n):
        for _ in range(n):
            p.process_dev(frames[k[0] % NT].data_ptr(), w, h, fg.data_ptr(), None, stream=st); k[0] += 1
    run(warm)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record(); run(n); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print("%-20s %.1f us per step of %d x 1080p  (%.1f Gpx/s), fg px %d" % (name, us, S, S * w * h / us / 1e3, int((fg != 0).sum())))
    p.close()
