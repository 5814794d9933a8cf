import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracking_b200 import blobs
w, h, S = 1920, 1080, 64
m = np.zeros((h, w), np.uint8)
for r in range(12):
    y0, x0 = (60 + 83 * r) % (h - 90), (100 + 150 * r) % (w - 120)
    m[y0:y0 + 80, x0:x0 + 100] = 255
d1 = torch.from_numpy(m).cuda()
dS = torch.from_numpy(np.stack([np.roll(m, 7 * s, 1) for s in range(S)])).cuda()
c1 = blobs.ConnectedComponents(w, h); cS = blobs.ConnectedComponents(w, h, max_images=S)
def timed(fn, iters=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
a = timed(lambda: c1.label_dev(d1.data_ptr(), w, h, True, None))
b = timed(lambda: cS.label_batch_dev(dS.data_ptr(), w, h, S, True, None), 20)
print(json.dumps(dict(dbg=os.environ.get("BGSB_CCL_DBG", "0"), pdl=os.environ.get("BGSB_NO_PDL", "0"), single_us=a, batch64_us=b)))
