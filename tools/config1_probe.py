"""Per-frame latency of FD + OPEN + CvBlobDetectorCC through the host API on the reference clip (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import blobs
clip = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "clips.npz"))["video_clip"]
fd = tb.FrameDifferenceBGS(); bd = blobs.CvBlobDetectorCC()
ts = []
for rep in range(3):
    for f in clip:
        t0 = time.perf_counter()
        fg, _ = fd.process(f)
        t1 = time.perf_counter()
        if fg is not None:
            m = blobs.morph(fg, [("erode", 1), ("dilate", 1)])
            t2 = time.perf_counter()
            bd.DetectNewBlob(m, [])
            t3 = time.perf_counter()
            ts.append((t1 - t0, t2 - t1, t3 - t2))
a = np.array(ts) * 1e3
print("frames", len(a), "median ms fd/morph/detect", np.median(a, 0), "max", a.max(0), "mean", a.mean(0))
print("first 6:", a[:6].round(2).tolist())
