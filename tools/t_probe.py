"""MOG2 temporal batches: us per 1080p frame for T in {1,2,4,8,16,32}, one stream (GPU box, measurement tooling)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
W, H, NF = 1920, 1080, 64
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF, W, H, stream=st)
for T in (1, 2, 4, 8, 16, 32):
    fg = torch.empty((T, H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((T, H, W, 3), dtype=torch.uint8, device="cuda")
    p = tb.MixtureOfGaussianV2BGS()
    k = [0]
    def run(n):
        for _ in range(n):
            p.process_batch_dev(d[k[0] % NF].data_ptr(), T, W, H, fg.data_ptr(), bg.data_ptr(), stream=st); k[0] = (k[0] + T) % NF
    run(128 // T)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 256 // T
    e0.record(); run(n); e1.record(); torch.cuda.synchronize()
    print("T=%d us/frame %.2f" % (T, e0.elapsed_time(e1) / (n * T) * 1e3))
    p.close()
