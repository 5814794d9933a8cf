"""ncu target: the labeller on one 1080p mask and on a batch of 64 (12 blobs each), table only and with labels.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ccl_launches.csv python tools/ccl_ncu_target.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracking_b200 import blobs
w, h, S = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = np.zeros((h, w), np.uint8)
for r in range(12):
    y0, x0 = (60 + 83 * r) % (h - 90), (100 + 150 * r) % (w - 120)
    m[y0:y0 + 80, x0:x0 + 100] = 255
d1 = torch.from_numpy(m).cuda()
dS = torch.from_numpy(np.stack([np.roll(m, 7 * s, 1) for s in range(S)])).cuda()
lab = torch.empty((S, h, w), dtype=torch.int32, device="cuda")
c1 = blobs.ConnectedComponents(w, h); cS = blobs.ConnectedComponents(w, h, max_images=S)
for _ in range(3):
    c1.label_dev(d1.data_ptr(), w, h, True, None)
    c1.label_dev(d1.data_ptr(), w, h, True, lab.data_ptr())
for _ in range(2):
    cS.label_batch_dev(dS.data_ptr(), w, h, S, True, None)
    cS.label_batch_dev(dS.data_ptr(), w, h, S, True, lab.data_ptr())
torch.cuda.synchronize()
print(len(c1.components()), len(cS.components(S - 1)))
