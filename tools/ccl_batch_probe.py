import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracking_b200 import blobs
w, h, S = 1920, 1080, 64
rng = np.random.default_rng(0)
ms = np.zeros((S, h, w), np.uint8)
for s in range(S):
    for r in range(12):
        y0 = (60 + 83 * r + 7 * s) % (h - 90); x0 = (100 + 150 * r + 31 * s) % (w - 120)
        ms[s, y0:y0 + 80, x0:x0 + 100] = 255
d = torch.from_numpy(ms).cuda()
cc = blobs.ConnectedComponents(w, h, max_images=S)
for _ in range(3):
    cc.label_batch_dev(d.data_ptr(), w, h, S, True, None)
torch.cuda.synchronize()
print(len(cc.components(5)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    cc.label_batch_dev(d.data_ptr(), w, h, S, True, None)
e1.record(); torch.cuda.synchronize()
print("batch of %d labelled in %.1f us (%.2f us per image)" % (S, e0.elapsed_time(e1) / 30 * 1e3, e0.elapsed_time(e1) / 30 / S * 1e3))
