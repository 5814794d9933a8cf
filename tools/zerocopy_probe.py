"""MOG2 T=1 kernel reading its frame from / writing its outputs to mapped pinned host memory directly (no copy
engine): us per 1080p frame.  GPU box, measurement tooling."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
W, H, NF = 1920, 1080, 16
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF, W, H, stream=st)
torch.cuda.synchronize()
h_in = d.cpu().pin_memory()
h_fg = torch.empty((H, W), dtype=torch.uint8).pin_memory(); h_bg = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
d_fg = torch.empty((H, W), dtype=torch.uint8, device="cuda"); d_bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
for name, src, fg, bg in (("all on device", d, d_fg, d_bg), ("input from host", h_in, d_fg, d_bg),
                          ("outputs to host", d, h_fg, h_bg), ("input from host, outputs to host", h_in, h_fg, h_bg),
                          ("input from host, mask to host, no bg", h_in, h_fg, None)):
    p = tb.MixtureOfGaussianV2BGS()
    def run(n, k0):
        for k in range(n):
            p.process_dev(src[(k0 + k) % NF].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr() if bg is not None else None, stream=st)
            torch.cuda.synchronize()          # synchronous per frame, like IBGS::process
    run(40, 0)
    t0 = time.perf_counter(); run(100, 40); dt = (time.perf_counter() - t0) / 100
    print("%-40s %.1f us/frame  %.2f Gpx/s" % (name, dt * 1e6, W * H / dt / 1e9))
    if bg is not None and fg is h_fg:
        ref = tb.MixtureOfGaussianV2BGS(); 
        for k in range(140):
            ref.process_dev(d[k % NF].data_ptr(), W, H, d_fg.data_ptr(), d_bg.data_ptr(), stream=st)
        torch.cuda.synchronize()
        print("   same outputs as the device run:", bool((d_fg.cpu() == h_fg).all()), bool((d_bg.cpu() == h_bg).all()))
        ref.close()
    p.close()
