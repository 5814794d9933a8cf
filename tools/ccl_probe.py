import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracking_b200 import blobs
w, h = 1920, 1080
m = np.zeros((h, w), np.uint8)
for r in range(12):
    y0 = (60 + 83 * r) % (h - 90); x0 = (100 + 150 * r) % (w - 120)
    m[y0:y0 + 80, x0:x0 + 100] = 255
rng = np.random.default_rng(0)
m[rng.random((h, w)) < 0.002] = 255
d = torch.from_numpy(m).cuda()
lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
cc = blobs.ConnectedComponents(w, h)
for _ in range(3):
    cc.label_dev(d.data_ptr(), w, h, True, lab.data_ptr())
torch.cuda.synchronize()
print(len(cc.components()))
