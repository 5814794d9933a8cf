#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the headline bench line
(config 1 = parity case, 3 = ABL/WMV x16 streams, 4 = full pipeline x64 streams, 5 = 4K temporal batches),
plus kernel-level figures for FD / morphology / CC and a PCIe probe.  One JSON object per line.

    python tools/bench_configs.py [--quick]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import tracking_b200 as tb                      # noqa: E402
from tracking_b200 import blobs, capi, synth    # noqa: E402

PEAK = 6539.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def out(**kw):
    print(json.dumps(kw), flush=True)


def simple_streams(algo_cls, name, bpp, S=16, w=1920, h=1080, NT=6, iters=40, retain=False, warm=4):
    st = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
    for t in range(NT):        # layout per time step: [S][1][h][w][3]
        synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
    fg = torch.empty((S, h, w), dtype=torch.uint8, device="cuda")
    bg = torch.empty((S, h, w, 3), dtype=torch.uint8, device="cuda")
    p = algo_cls(nstreams=S, **({"retainInput": 1} if retain else {}))
    k = [0]

    def step():
        p.process_dev(frames[k[0] % NT].data_ptr(), w, h, fg.data_ptr(), bg.data_ptr(), stream=st)
        k[0] += 1
    dt = timed(step, iters, warm=warm)
    px = S * w * h
    out(config={"FD": "fd", "ABL": "3", "WMV": "3"}.get(name.split()[0], "sibling"), algo=name, streams=S, resolution=[w, h], ms_per_step=dt * 1e3,
        mpixel_s=px / dt / 1e6, algorithmic_bytes_per_px=bpp, achieved_gbs=px * bpp / dt / 1e9,
        frac_of_measured_peak=px * bpp / dt / 1e9 / PEAK)
    p.close()


def mog2_batches(S, w, h, Ts, label, NF=48, iters=2):
    """Temporal batches on one continuous stream: NF consecutive frames are cycled, every T divides NF, so every
    setting sees exactly the same video (same share of generic-path pixels)."""
    st = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((S, NF, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(frames.data_ptr(), S, NF, w, h, t0=0, stream=st)
    for T in Ts:
        p = tb.MixtureOfGaussianV2BGS(nstreams=S)
        nwin = NF // T
        wins = [frames[:, i * T:(i + 1) * T].contiguous() for i in range(nwin)]
        fg = torch.empty((S, T, h, w), dtype=torch.uint8, device="cuda")
        bg = torch.empty((S, T, h, w, 3), dtype=torch.uint8, device="cuda")

        def sweep():
            for i in range(nwin):
                p.process_batch_dev(wins[i].data_ptr(), T, w, h, fg.data_ptr(), bg.data_ptr(), stream=st)
        dt = timed(sweep, iters, warm=2)              # seconds per NF frames
        px = S * NF * w * h
        nm = np.concatenate([p.export_state(s)[1] for s in range(min(S, 2))])
        out(config=label, algo="MOG2", streams=S, T=T, resolution=[w, h], us_per_frame_per_stream=dt / NF / S * 1e6,
            mpixel_s=px / dt / 1e6, mean_live_modes=float(nm.mean()),
            live_bytes_per_px_frame=(2 + 40 * float(nm.mean())) / T + 7,
            dense_model_bytes_per_px_frame=202.0 / T + 7)
        p.close()
        del fg, bg, wins


def pipeline(S=64, w=1920, h=1080, NT=24, iters=48):
    """Config 4 through the pipeline object (packed mask between the stages); stage split: tools/pipeline_probe.py."""
    from tracking_b200.pipeline import ForegroundPipeline
    st = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
    for t in range(NT):
        synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
    pipe = ForegroundPipeline(5, nstreams=S)
    k = [0]

    def step():
        pipe.process_dev(frames[k[0] % NT].data_ptr(), w, h, None, None, None, stream=st)
        k[0] += 1
    for _ in range(3 * NT):     # let the models settle on the moving scene
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    pipe.join_dev(st)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / iters * 1e-3
    ncomp = len(pipe.components(0))
    nm = pipe.export_mog2_state(0)[1]
    live = 5 + 40 * float(nm.mean()) + 0.5
    px = S * w * h
    out(config="4", algo="MOG2+OPEN+CC (bgsb_pipeline, packed mask)", streams=S, resolution=[w, h], ms_per_step=dt * 1e3,
        mpixel_s=px / dt / 1e6, components_stream0=ncomp, mean_live_modes=float(nm.mean()), live_bytes_per_px=live,
        live_gbs=px * live / dt / 1e9, frac_of_peak_live=px * live / dt / 1e9 / PEAK, dense_model_bytes_per_px=218,
        dense_equiv_gbs=px * 218 / dt / 1e9)
    pipe.close()


def config1(clip):
    """FD + OPEN + CC on the reference video clip: parity case, timed per frame through the host API."""
    fd = tb.FrameDifferenceBGS()
    bd = blobs.CvBlobDetectorCC()
    for f in clip[:4]:                      # first calls pay CUDA module loading and allocations
        fg, _ = fd.process(f)
        if fg is not None:
            bd.DetectNewBlob(blobs.morph(fg, [("erode", 1), ("dilate", 1)]), [])
    t0 = time.perf_counter()
    n = 0
    for f in clip[4:]:
        fg, _ = fd.process(f)
        if fg is None:
            continue
        m = blobs.morph(fg, [("erode", 1), ("dilate", 1)])
        bd.DetectNewBlob(m, [])
        n += 1
    dt = time.perf_counter() - t0
    out(config="1", algo="FD+OPEN+CvBlobDetectorCC (host API, per frame sync)", frames=n, resolution=list(clip.shape[2:0:-1]),
        ms_per_frame=dt / n * 1e3)


def pcie_probe(w=1920, h=1080):
    hin = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
    hout = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    din = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    dout = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
    t_up = timed(lambda: din.copy_(hin, non_blocking=True), 50)
    t_dn = timed(lambda: hout.copy_(dout, non_blocking=True), 50)
    s2 = torch.cuda.Stream()

    def both():
        din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    t_b = timed(both, 50)
    torch.cuda.synchronize()
    out(probe="pcie", h2d_gbs=hin.numel() / t_up / 1e9, d2h_gbs=hout.numel() / t_dn / 1e9,
        frame_up_us=t_up * 1e6, frame_down_us=t_dn * 1e6, both_us=t_b * 1e6,
        e2e_ceiling_mpixel_s=w * h / max(t_up, t_dn) / 1e6)


def ccl_kernel_probe(w=1920, h=1080):
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    for name, dens in (("sparse blobs", None), ("salt noise 0.2%", 0.002), ("random 30%", 0.3)):
        if dens is None:
            m = np.zeros((h, w), np.uint8)
            for r in range(12):
                m[60 + 83 * r % (h - 90):60 + 83 * r % (h - 90) + 80, (100 + 150 * r) % (w - 120):(100 + 150 * r) % (w - 120) + 100] = 255
        else:
            m = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        d = torch.from_numpy(m).cuda()
        lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
        cc = blobs.ConnectedComponents(w, h)
        t1 = timed(lambda: cc.label_dev(d.data_ptr(), w, h, True, lab.data_ptr(), stream=st), 30)
        t0 = timed(lambda: cc.label_dev(d.data_ptr(), w, h, True, None, stream=st), 30)
        n = len(cc.components())
        out(probe="ccl", mask=name, components=n, us_with_labels=t1 * 1e6, us_table_only=t0 * 1e6,
            gbs_5B_per_px=w * h * 5 / t1 / 1e9)
        cc.close()


def main():
    quick = "--quick" in sys.argv
    torch.cuda.set_device(0)
    if "--mog2t" in sys.argv:                                 # just the temporal-batch lines
        mog2_batches(1, 1920, 1080, [1, 8, 16], "2-T")
        mog2_batches(16, 1920, 1080, [1, 16], "16x1080p-T")
        mog2_batches(4, 3840, 2160, [1, 16], "5", NF=16)
        return
    if "--pipeline" in sys.argv:                              # just config 4 (16 streams): ncu launch-list target
        pipeline(S=16)
        return
    if "--abl" in sys.argv:
        simple_streams(tb.AdaptiveBackgroundLearning, "ABL", 10)
        return
    if "--wmv" in sys.argv:                                   # just the WMV line
        simple_streams(tb.WeightedMovingVarianceBGS, "WMV", 10 + 6)
        return
    if "--sfd" in sys.argv:                                   # StaticFD: 3 in + 3 bg read + 1 mask + 3 bg image
        simple_streams(tb.StaticFrameDifferenceBGS, "StaticFD", 10)
        return
    if "--wmm" in sys.argv:                                   # just the WMM line
        simple_streams(tb.WeightedMovingMeanBGS, "WMM", 10 + 6 + 3)
        return
    if "--dp" in sys.argv:                                    # just the DP package's simple models
        simple_streams(tb.DPAdaptiveMedianBGS, "DPAdaptiveMedian (3 in + 3 model + 1 mask + 3/7 model write)", 7 + 3 / 7)
        simple_streams(tb.DPMeanBGS, "DPMean (3 in + 12 + 12 mean + 1 mask)", 28)
        simple_streams(tb.DPWrenGABGS, "DPWrenGA (3 in + 16 + 16 model + 1 mask)", 36)
        # ring of 16 samples full after 76 frames; then one frame in 5 runs the update pass (16 x (3 sample + 2 + 2 sum) + 11 B/px)
        simple_streams(tb.DPPratiMediodBGS, "DPPratiMediod (7 B/px every frame + 123 B/px on one frame in 5, ring full)", 7 + 123 / 5, S=8, warm=80)
        simple_streams(tb.SigmaDeltaBGS, "SigmaDelta (3 in + 3 + 3 Mt + 3 + 3 Vt + 1 mask)", 16)
        return
    if "--asbl" in sys.argv:                                  # just the ASBL line
        simple_streams(tb.AdaptiveSelectiveBackgroundLearning, "ASBL", 3 + 2 + 1 + 1)
        return
    z = np.load(os.path.join(ROOT, "tests", "golden", "clips.npz"))
    config1(z["video_clip"])
    pcie_probe()
    simple_streams(tb.FrameDifferenceBGS, "FD retainInput (SURVEY 8d bytes: 7 B/px)", 7, retain=True)
    simple_streams(tb.WeightedMovingVarianceBGS, "WMV retainInput (SURVEY 8d bytes: 10 B/px)", 10, retain=True)
    simple_streams(tb.WeightedMovingMeanBGS, "WMM retainInput (9 in + 1 mask + 3 bg)", 13, retain=True)
    simple_streams(tb.FrameDifferenceBGS, "FD", 7 + 3)        # device path also writes the 3 B/px history
    simple_streams(tb.AdaptiveBackgroundLearning, "ABL", 10)
    simple_streams(tb.WeightedMovingVarianceBGS, "WMV", 10 + 6)  # device path writes both history images
    simple_streams(tb.AdaptiveSelectiveBackgroundLearning, "ASBL", 3 + 2 + 1 + 1)   # in, gray model r/w, mask, gray bg image
    simple_streams(tb.DPZivkovicAGMMBGS, "DPZivkovicAGMM (3 modes; bytes = 3 in + 1 mask + 2 counts + 40 per live mode, 1.4 live modes assumed)", 62)
    # the DP package's simple models: bytes = 3 in + model read + model written + 1 mask (AdaptiveMedian writes its model on
    # one frame in samplingRate = 7)
    simple_streams(tb.DPAdaptiveMedianBGS, "DPAdaptiveMedian (3 in + 3 model + 1 mask + 3/7 model write)", 7 + 3 / 7)
    simple_streams(tb.DPMeanBGS, "DPMean (3 in + 12 + 12 mean + 1 mask)", 28)
    simple_streams(tb.DPWrenGABGS, "DPWrenGA (3 in + 16 + 16 model + 1 mask)", 36)
    # ring of 16 samples full after 76 frames; then one frame in 5 runs the update pass (16 x (3 sample + 2 + 2 sum) + 11 B/px)
    simple_streams(tb.DPPratiMediodBGS, "DPPratiMediod (7 B/px every frame + 123 B/px on one frame in 5, ring full)", 7 + 123 / 5, S=8, warm=80)
    simple_streams(tb.SigmaDeltaBGS, "SigmaDelta (3 in + 3 + 3 Mt + 3 + 3 Vt + 1 mask)", 16)
    ccl_kernel_probe()
    import fanout_probe                                   # tools/fanout_probe.py: FrameProcessor fan-out vs four uploads
    fanout_probe.main()
    pipeline(S=16 if quick else 64)
    mog2_batches(1, 1920, 1080, [1, 2, 4, 8, 16], "2-T")
    mog2_batches(16, 1920, 1080, [1, 4, 16], "16x1080p-T")
    mog2_batches(4 if quick else 16, 3840, 2160, [1, 4, 8, 16], "5", NF=16 if quick else 32)


if __name__ == "__main__":
    main()
