"""H2D bandwidth: one copy vs the frame split over k concurrent streams, pinned vs write-combined host memory."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracking_b200 import capi
w, h = 1920, 1080
n = w * h * 3
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(7)
def timed(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
t = timed(lambda: dev.copy_(host, non_blocking=True))
print("1 stream: %.1f us  %.1f GB/s" % (t * 1e6, n / t / 1e9))
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    cs = (n // k + 15) // 16 * 16
    def f():
        ev = torch.cuda.Event(); ev.record()
        for i, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                a, b = i * cs, min(n, (i + 1) * cs)
                dev[a:b].copy_(host[a:b], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    t = timed(f)
    print("%d streams: %.1f us  %.1f GB/s" % (k, t * 1e6, n / t / 1e9))
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
t = timed(lambda: dbig.copy_(big, non_blocking=True), 10)
print("256 MB copy: %.1f GB/s" % (big.numel() / t / 1e9))
t = timed(lambda: big.copy_(dbig, non_blocking=True), 10)
print("256 MB D2H: %.1f GB/s" % (big.numel() / t / 1e9))
