#!/usr/bin/env python
"""Headline benchmark: MOG2 Mpixel/s at 1080p (BASELINE.json metric), one camera stream per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: FRAMES_PER_STEP
consecutive 1920x1080 BGR frames of one camera stream, advanced frame by frame (T = 1, the judged
roofline configuration: 209 algorithmic bytes per pixel per frame) through
MixtureOfGaussianV2BGS -- mask + background image produced for every frame, exactly what
`IBGS::process` returns.  Frames are resident in HBM when the timed region starts (K-GEN).

Multi-GPU: independent camera streams, one stream per GPU, no data-path collective (weak scaling);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the timed region.

The JSON line also carries
  e2e          the same metric through the host-buffer C-ABI call bgsb_process (pinned host frames,
               H2D + kernel + D2H of mask and background inside the timed region)
  roofline     achieved algorithmic HBM GB/s of the MOG2 kernel vs the measured copy peak
  cpu_baseline the reference's CPU path (OpenCV calls replayed call-for-call by oracle/cv2_chain.py,
               the reference C++ itself cannot be built in this image) timed on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
NPX = W * H
FRAMES_PER_STEP = 1024        # 34 s of 30 fps video per step; ~36 ms of GPU time, so that K >= 10 steps span several
                               # nvidia-smi clock samples
NFRAMES_RESIDENT = 128          # distinct synthetic frames kept in HBM (796 MB) and cycled
MOG2_BYTES_PER_PX = 209         # dense model: in 3 + state 101 read + 101 write + mask 1 + bg 3  
MOG2_FIXED_BYTES_PER_PX = 9     # in 3 + nmodes 1 read + 1 write + mask 1 + bg 3
MOG2_BYTES_PER_LIVE_MODE = 40   # weight, variance, 3-vector mean: 20 B read + 20 B written
METRIC = "MOG2 Mpixel/s at 1080p"
UNIT = "Mpixel/s"
WORKLOAD = "MixtureOfGaussianV2 (MOG2, K=5) on one synthetic 1920x1080 BGR stream per GPU, T=1"


_STDOUT_FD = None


def emit(text):
    """The contract's one line, on the process's original stdout."""
    sys.stdout.flush()
    if _STDOUT_FD is None:
        print(text, flush=True)
    else:
        os.write(_STDOUT_FD, (text + "\n").encode())


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        rows = [r for (ts, r) in self.rows if t_begin is None or (t_begin <= ts <= t_end + 0.15)]
        window = "timed region"
        if len(rows) < 2:            # region shorter than the sampling period: widen to the load phase around it
            rows = [r for (ts, r) in self.rows if t_begin is None or (t_begin - 1.0 <= ts <= t_end + 1.0)]
            window = "timed region +-1 s (region shorter than 2 samples)"
        for r in rows:
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def cpu_reference_run(steps, warmup, frames_per_step):
    """The reference's CPU path on this box: MixtureOfGaussianV2BGS::process replayed with OpenCV
    (mog(in, fg, 0.05) + getBackgroundImage + threshold), all host threads, same synthetic stream."""
    import cv2
    from oracle import cv2_chain
    from tracking_b200 import synth
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    nuniq = min(16, max(2, frames_per_step))
    frames = [synth.frame(W, H, t) for t in range(nuniq)]
    bgs = cv2_chain.MixtureOfGaussianV2BGS()
    for i in range(max(1, warmup) * frames_per_step):
        bgs.process(frames[i % nuniq])
    t0 = time.perf_counter()
    for i in range(steps * frames_per_step):
        bgs.process(frames[i % nuniq])
    dt = time.perf_counter() - t0
    return steps * frames_per_step * NPX / dt / 1e6, dt, cores, cv2.getNumThreads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: 4 frames per step keeps `--steps 20 --warmup 5` within ~1 minute on 8 vCPUs
    fps = 4
    val, dt, cores, threads = cpu_reference_run(args.steps, args.warmup, fps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": fps, "resolution": [W, H]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d frames/step x %d steps of the same synthetic 1080p stream; OpenCV %s "
                                       "calls of MixtureOfGaussianV2BGS::process replayed call-for-call "
                                       "(reference C++ not buildable here: needs OpenCV 2.4 headers), %d host cores"
                                       % (fps, args.steps, __import__("cv2").__version__, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))
    return 0


def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import tracking_b200 as tb
    from tracking_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation, and rank 0's stdout is the ONE JSON line
        # of the contract: from here on file descriptor 1 is stderr, the line goes to the saved descriptor (emit()).
        global _STDOUT_FD
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    F, K, Wm = args.frames_per_step, args.steps, args.warmup
    stream = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((NFRAMES_RESIDENT, H, W, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(frames.data_ptr(), 1, NFRAMES_RESIDENT, W, H, t0=0, seed0=synth.SEED0 + rank, stream=stream)
    d_fg = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    d_bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    bgs = tb.MixtureOfGaussianV2BGS(device=local, kernelVariant=args.kernel_variant)
    fptr = [frames[i].data_ptr() for i in range(NFRAMES_RESIDENT)]
    fgp, bgp = d_fg.data_ptr(), d_bg.data_ptr()

    def step(i):
        base = (i * F) % NFRAMES_RESIDENT
        for t in range(F):
            bgs.process_dev(fptr[(base + t) % NFRAMES_RESIDENT], W, H, fgp, bgp, stream=stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.3 s before its first sample: start it ahead of the warm-up
    for i in range(Wm):
        step(i)
    barrier()
    t_begin = time.perf_counter()
    launches0 = capi.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        step(Wm + i)
    ev1.record()
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = capi.kernel_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    total_px = world * K * F * NPX
    value = total_px / (ms_max * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (mog2_t1_kernel, csrc/mog2_t1.cu): per-launch figures, this rank ----
    # Algorithmic bytes: SURVEY 8(d) quotes 209 B/px for a DENSE model (all K=5 modes live).  Like the
    # reference's CPU loop (`for mode < nmodes`), the kernel only touches LIVE modes, so the bytes the
    # algorithm has to move depend on the stream: 9 + 40 * (mean live modes per pixel).  `achieved` uses
    # that live figure, measured on the model state right after the timed region (it cannot overstate
    # the kernel); the dense-model equivalent is reported next to it.
    peak, peak_src = measured_peak_gbs()
    launch_ms = ms / max(launches, 1)
    _, nm_host = bgs.export_state()
    mean_modes = float(nm_host.mean())
    live_bpp = MOG2_FIXED_BYTES_PER_PX + MOG2_BYTES_PER_LIVE_MODE * mean_modes
    alg_bytes = live_bpp * NPX
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    dense_equiv = MOG2_BYTES_PER_PX * NPX / (launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "mog2_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, tb, capi, frames, F, K, world, local, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cdt, cores, threads = cpu_reference_run(8, 1, 4)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "32 frames (8 steps x 4) of the same synthetic 1080p stream after 4 warm-up frames; OpenCV "
                         "call-for-call replay of MixtureOfGaussianV2BGS::process on %d host cores, %.1f s" % (cores, cdt)}
    extra = {"mean_live_modes_per_px": mean_modes, "algorithmic_bytes_per_px_live": live_bpp,
             "algorithmic_bytes_per_px_dense_model": MOG2_BYTES_PER_PX, "achieved_dense_model_equiv_gbs": dense_equiv,
             "note": "achieved = (9 + 40*mean live modes) B/px * px per launch / CUDA-event launch time; dead modes are "
                     "never touched (same as the reference's `mode < nmodes` loop); the kernel is issue/latency bound "
                     "on this stream, not HBM bound (see DESIGN.md, profiles/)"}
    return finish(args, rank, world, value, K, Wm, ms_max, F, clocks, e2e, launches, achieved, peak, traffic, peak_src,
                  alg_bytes, launch_ms, cpu, extra)


def run_e2e(args, tb, capi, frames, F, K, world, local, barrier):
    import ctypes as C
    import torch
    import torch.distributed as dist
    nh = 16
    h_in = torch.empty((nh, H, W, 3), dtype=torch.uint8).pin_memory()
    h_in.copy_(frames[:nh].cpu())
    h_fg = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    h_bg = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    e2e_bgs = tb.MixtureOfGaussianV2BGS(device=local)
    L = capi.lib()
    fv, bv = C.c_int(0), C.c_int(0)
    inp = [C.c_void_p(h_in[i].data_ptr()) for i in range(nh)]
    ofg, obg = C.c_void_p(h_fg.data_ptr()), C.c_void_p(h_bg.data_ptr())

    def e2e_step(i):
        for t_ in range(F):
            rc = L.bgsb_process(e2e_bgs._h, inp[(i * F + t_) % nh], W, H, W * 3, ofg, W, obg, W * 3,
                                C.byref(fv), C.byref(bv))
            if rc:
                capi.check(rc)

    Ke = max(2, min(K, 10))
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_step(2 + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * Ke * F * NPX / te.item() / 1e6
    assert int(h_fg.max()) in (0, 255)

    # the same frames through the queued form of the call (bgsb_submit / bgsb_wait, the capture-loop ingest): the
    # upload of frame t+1 overlaps the download of frame t; one wait per step
    nq = 4
    q_fg = torch.empty((nq, H, W), dtype=torch.uint8).pin_memory()
    q_bg = torch.empty((nq, H, W, 3), dtype=torch.uint8).pin_memory()
    qfg = [C.c_void_p(q_fg[i].data_ptr()) for i in range(nq)]
    qbg = [C.c_void_p(q_bg[i].data_ptr()) for i in range(nq)]

    def queued_step(i):
        for t_ in range(F):
            rc = L.bgsb_submit(e2e_bgs._h, inp[(i * F + t_) % nh], W, H, W * 3, qfg[t_ % nq], W, qbg[t_ % nq], W * 3,
                               C.byref(fv), C.byref(bv))
            if rc:
                capi.check(rc)
        capi.check(L.bgsb_wait(e2e_bgs._h))

    queued_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        queued_step(1 + i)
    torch.cuda.synchronize()
    tq = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
    q_val = world * Ke * F * NPX / tq.item() / 1e6
    assert int(q_fg.max()) in (0, 255)
    return {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": F * NPX * 3, "d2h_bytes_per_step": F * NPX * 4,
            "steps": Ke,
            "queued": {"value": q_val, "unit": UNIT,
                       "note": "bgsb_submit / bgsb_wait: same frames, same copies per frame, frames queued so that "
                               "the next upload overlaps this frame's download; one wait per step"},
            "note": "bgsb_process (IBGS::process boundary): pinned host BGR frame in, mask + background image out, "
                    "synchronous per frame; upload/kernel/download of 2 row bands overlap inside the call"}


def finish(args, rank, world, value, K, Wm, ms_max, F, clocks, e2e, launches, achieved, peak, traffic, peak_src,
           alg_bytes, launch_ms, cpu, extra):
    import torch.distributed as dist
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_step": F, "resolution": [W, H], "temporal_batch": 1,
                           "streams_per_gpu": 1, "parallelism": "independent camera streams, %d GPU(s), no collective" % world,
                           "l2": "inputs larger than L2: 209 MB model state + 796 MB of resident frames cycled"},
                "clocks": clocks,
                "e2e": e2e,
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "mog2_t1_kernel<SHADOWS,0,GROUP>", "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                             "frac_of_8TBs_nominal": achieved / 8000.0, **extra},
                "cpu_baseline": cpu}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--frames-per-step", type=int, default=FRAMES_PER_STEP)
    ap.add_argument("--kernel-variant", type=int, default=0, help="0 production MOG2 kernel, 1 straight restatement (A/B)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
