"""OPEN 3x3 (erode + dilate) on 64 x 1080p masks in one launch: us per batch (GPU box, measurement tooling)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracking_b200 import blobs
w, h, S = 1920, 1080, 64
rng = np.random.default_rng(0)
m = (rng.random((S, h, w)) < 0.02).astype(np.uint8) * 255
for s in range(S):
    m[s, 100 + s:400 + s, 200:900] = 255
d = torch.from_numpy(m).cuda(); o = torch.empty_like(d)
st = torch.cuda.current_stream().cuda_stream
ops = [("erode", 1), ("dilate", 1)]
for _ in range(3):
    blobs.morph_dev(d.data_ptr(), w, h, S, ops, o.data_ptr(), stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    blobs.morph_dev(d.data_ptr(), w, h, S, ops, o.data_ptr(), stream=st)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 30 * 1e3
print("OPEN on %d masks: %.1f us  %.0f GB/s (2 B/px)" % (S, us, S * w * h * 2 / us / 1e3))
