"""FrameProcessor-style fan-out vs one bgsb_process per plugin: 1080p, pinned host buffers, FD + WMV + MOG2 + ABL
(the four hot-path plugins FrameProcessor.cpp:176-195 can enable together).  GPU box, measurement tooling."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb                      # noqa: E402
from tracking_b200 import capi, synth           # noqa: E402


def main(w=1920, h=1080, NF=16, iters=60):
    d = torch.empty((NF, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), 1, NF, w, h)
    torch.cuda.synchronize()
    host = d.cpu().pin_memory()
    classes = [tb.FrameDifferenceBGS, tb.WeightedMovingVarianceBGS, tb.MixtureOfGaussianV2BGS, tb.AdaptiveBackgroundLearning]
    n = len(classes)
    fgs = [torch.empty((h, w), dtype=torch.uint8).pin_memory() for _ in range(n)]
    bgs = [torch.empty((h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(n)]
    lib = capi.lib()
    res = {}
    for mode in ("separate", "fanout"):
        ps = [c() for c in classes]
        ctxs = (C.c_void_p * n)(*[p._h for p in ps])
        fgp = (C.c_void_p * n)(*[f.data_ptr() for f in fgs])
        bgp = (C.c_void_p * n)(*[b.data_ptr() for b in bgs])
        fgst = (C.c_size_t * n)(*[w] * n)
        bgst = (C.c_size_t * n)(*[3 * w] * n)
        fv, bv = (C.c_int * n)(), (C.c_int * n)()
        a, b = C.c_int(0), C.c_int(0)

        def step(k):
            f = host[k % NF]
            if mode == "fanout":
                capi.check(lib.bgsb_process_fanout(ctxs, n, C.c_void_p(f.data_ptr()), w, h, 3 * w, fgp, fgst, bgp, bgst, fv, bv))
            else:
                for i, p in enumerate(ps):
                    capi.check(lib.bgsb_process(p._h, C.c_void_p(f.data_ptr()), w, h, 3 * w, C.c_void_p(fgs[i].data_ptr()), w,
                                                C.c_void_p(bgs[i].data_ptr()), 3 * w, C.byref(a), C.byref(b)))
        for k in range(8):
            step(k)
        t0 = time.perf_counter()
        for k in range(8, 8 + iters):
            step(k)
        dt = (time.perf_counter() - t0) / iters
        res[mode] = dt
        sig = [int(f.long().sum()) for f in fgs]
        res[mode + "_sig"] = sig
        for p in ps:
            p.close()
    out = {"config": "fanout", "algos": "FD+WMV+MOG2+ABL", "resolution": [w, h],
           "separate_us_per_frame": res["separate"] * 1e6, "fanout_us_per_frame": res["fanout"] * 1e6,
           "speedup": res["separate"] / res["fanout"], "same_masks": res["separate_sig"] == res["fanout_sig"],
           "h2d_bytes_per_frame": {"separate": 4 * w * h * 3, "fanout": w * h * 3}, "d2h_bytes_per_frame": 4 * w * h + 2 * w * h * 3,
           "mpixel_s_per_plugin_fanout": w * h / res["fanout"] / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
