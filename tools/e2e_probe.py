import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import capi, synth
W, H, NF = 1920, 1080, 16
d = torch.empty((NF, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF, W, H)
h_in = torch.empty((NF, H, W, 3), dtype=torch.uint8).pin_memory(); h_in.copy_(d.cpu())
h_fg = torch.empty((H, W), dtype=torch.uint8).pin_memory()
h_bg = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
L = capi.lib()
import numpy as np
wc = C.c_void_p()
capi.check(L.bgsb_host_alloc(C.byref(wc), NF * H * W * 3, 1))
wc_np = np.ctypeslib.as_array(C.cast(wc, C.POINTER(C.c_uint8)), shape=(NF, H, W, 3))
wc_np[:] = h_in.numpy()
for name, src in (("pinned", [h_in[i].data_ptr() for i in range(NF)]), ("write-combined", [wc.value + i * H * W * 3 for i in range(NF)])):
    p = tb.MixtureOfGaussianV2BGS()
    fv, bv = C.c_int(0), C.c_int(0)
    def run(n):
        for i in range(n):
            L.bgsb_process(p._h, C.c_void_p(src[i % NF]), W, H, W * 3, C.c_void_p(h_fg.data_ptr()), W,
                           C.c_void_p(h_bg.data_ptr()), W * 3, C.byref(fv), C.byref(bv))
    run(64); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(256); dt = (time.perf_counter() - t0) / 256
    print(name, "input: us/frame %.1f  Gpx/s %.2f" % (dt * 1e6, W * H / dt / 1e9), flush=True)
    p.close()
for bands in (1, 2, 3, 4, 6, 8):
    for want_bg in (1, 0):
        p = tb.MixtureOfGaussianV2BGS(hostBands=bands)
        fv, bv = C.c_int(0), C.c_int(0)
        def run(n):
            for i in range(n):
                L.bgsb_process(p._h, C.c_void_p(h_in[i % NF].data_ptr()), W, H, W * 3, C.c_void_p(h_fg.data_ptr()), W,
                               C.c_void_p(h_bg.data_ptr()) if want_bg else None, W * 3, C.byref(fv), C.byref(bv))
        run(64); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(256); dt = (time.perf_counter() - t0) / 256
        print("bands", bands, "bg" if want_bg else "no-bg", "us/frame %.1f  Gpx/s %.2f" % (dt * 1e6, W * H / dt / 1e9), flush=True)
        p.close()
# stage split of the synchronous call (parameter "trace"): event spans of the uploads, kernels, downloads + wall clock
for bands in (1, 2, 3, 4):
    p = tb.MixtureOfGaussianV2BGS(hostBands=bands, trace=1)
    fv, bv = C.c_int(0), C.c_int(0)
    acc = {}
    for i in range(96):
        L.bgsb_process(p._h, C.c_void_p(h_in[i % NF].data_ptr()), W, H, W * 3, C.c_void_p(h_fg.data_ptr()), W,
                       C.c_void_p(h_bg.data_ptr()), W * 3, C.byref(fv), C.byref(bv))
        if i >= 32:
            for k, v in p.trace_last().items():
                acc[k] = acc.get(k, 0.0) + v / 64
    print("trace bands", bands, {k: round(v * 1e3, 1) for k, v in acc.items() if k.endswith("_ms")}, "(us)", flush=True)
    p.close()
