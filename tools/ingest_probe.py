"""Host-buffer throughput of one 1080p stream: synchronous bgsb_process vs the queued bgsb_submit / bgsb_wait.
Pinned host frames in, mask + background image out, every frame; wall clock around the whole run."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import tracking_b200 as tb

H, W, NF, K = 1080, 1920, 8, 200
rng = np.random.default_rng(0)
base = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
ins = []
for i in range(NF):
    a = tb.pinned_empty((H, W, 3))
    a[...] = np.clip(base.astype(np.int16) + rng.integers(-10, 11, (H, W, 3)), 0, 255).astype(np.uint8)
    ins.append(a)
fgs = [tb.pinned_empty((H, W)) for _ in range(NF)]

for name in sys.argv[1:] or ["MixtureOfGaussianV2BGS", "AdaptiveBackgroundLearning", "FrameDifferenceBGS", "WeightedMovingVarianceBGS"]:
    p = getattr(tb, name)()
    bgs = [tb.pinned_empty((H, W, 3) if p.BG_CHANNELS == 3 else (H, W)) for _ in range(NF)]
    import ctypes as C
    from tracking_b200 import capi
    L = capi.lib()
    fv, bv = C.c_int(0), C.c_int(0)
    def sync(i):
        j = i % NF
        capi.check(L.bgsb_process(p._h, ins[j].ctypes.data, W, H, W * 3, fgs[j].ctypes.data, W, bgs[j].ctypes.data,
                                  bgs[j].strides[0], C.byref(fv), C.byref(bv)))
    def sub(i):
        j = i % NF
        capi.check(L.bgsb_submit(p._h, ins[j].ctypes.data, W, H, W * 3, fgs[j].ctypes.data, W, bgs[j].ctypes.data,
                                 bgs[j].strides[0], C.byref(fv), C.byref(bv)))
    for i in range(20):
        sync(i)
    t0 = time.perf_counter()
    for i in range(K):
        sync(i)
    ts = (time.perf_counter() - t0) / K
    for i in range(20):
        sub(i)
    capi.check(L.bgsb_wait(p._h))
    t0 = time.perf_counter()
    for i in range(K):
        sub(i)
    capi.check(L.bgsb_wait(p._h))
    tq = (time.perf_counter() - t0) / K
    # bounded queue depth (what a capture loop does: wait for frame t-2 before reusing its buffers)
    print("%-28s process %.1f us/frame (%.2f Gpx/s)   submit/wait %.1f us/frame (%.2f Gpx/s)"
          % (name, ts * 1e6, H * W / ts / 1e9, tq * 1e6, H * W / tq / 1e9), flush=True)
    p.close()
