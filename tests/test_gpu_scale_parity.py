"""Parity at the geometries and kernels the benchmarks run (VERDICT r1 "next 1"): every test name says which
kernel it executes, and asserts it through the launch routing rules of csrc/capi.cu (launch_range):

    one stream,  T == 1            mog2_t1_kernel<.,0,false>
    stream group, T == 1           mog2_t1_kernel<.,0,true>   (evict-last policy on the state rows)
    one stream,  2 <= T < 6        T launches of the T == 1 kernel
    one stream,  T >= 6; any group with T > 1      mog2_fused_kernel

All comparisons are bit-exact against the C oracle (oracle/c/bgs_oracle.c, pinned to OpenCV in tests/test_oracle_pin.py)
on the synthetic video of SURVEY 8(d), generated on the device and downloaded, so both sides see the same bytes.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import stress_sequence

pytestmark = pytest.mark.gpu


def synth_resident(S, n, w, h, t0=0):
    """[S][n][h][w][3] device tensor of the K-GEN video + its host copy."""
    import torch
    from tracking_b200 import synth
    d = torch.empty((S, n, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), S, n, w, h, t0=t0)
    torch.cuda.synchronize()
    return d, d.cpu().numpy()


def assert_state_equal(p, o, stream_index=0):
    planes, nm = p.export_state(stream_index)
    assert np.array_equal(nm, o.nmodes), "mode counts"
    for m in range(5):
        live = o.nmodes > m
        assert np.array_equal(planes[m * 5 + 0][live], o.gmm[:, m, 0][live]), "weight, mode %d" % m
        assert np.array_equal(planes[m * 5 + 1][live], o.gmm[:, m, 1][live]), "variance, mode %d" % m
        for c in range(3):
            assert np.array_equal(planes[m * 5 + 2 + c][live], o.mean[:, m, c][live]), "mean, mode %d" % m


def test_t1_kernel_1080p_80_frames_through_the_wrap(oracle):
    """mog2_t1_kernel<.,0,false>: bench.py's loop (one 1080p stream, T = 1, resident frames cycled, so the scene jumps at
    every wrap) for 80 frames over a 32-frame ring: masks, background images and the exported mixture state."""
    import torch
    import tracking_b200 as tb
    w, h, nres, n = 1920, 1080, 32, 80
    d, host = synth_resident(1, nres, w, h)
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    slow_px = 0
    for t in range(n):
        p.process_dev(d[0, t % nres].data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
        ofg, obg = o.process(host[0, t % nres])
        assert np.array_equal(d_fg.cpu().numpy(), ofg), "mask, frame %d" % t
        assert np.array_equal(d_bg.cpu().numpy(), obg), "background, frame %d" % t
        slow_px += int((ofg != 0).sum())
    assert tb.kernel_launch_count() - before == n            # one launch per frame: the T == 1 kernel
    assert slow_px > 0
    assert_state_equal(p, o)
    assert float(o.nmodes.mean()) > 1.05                      # the steady state has multi-mode pixels
    p.close()


def test_group_t1_kernel_4x1080p_40_frames(oracle):
    """mog2_t1_kernel<.,0,true> (the config-4 / config-5 kernel): 4 streams x 1080p advanced by one launch per frame,
    40 frames over a 20-frame ring, against four independent oracles."""
    import torch
    import tracking_b200 as tb
    S, w, h, nres, n = 4, 1920, 1080, 20, 40
    d, host = synth_resident(S, nres, w, h)
    p = tb.MixtureOfGaussianV2BGS(nstreams=S)
    os_ = [oracle.MixtureOfGaussianV2BGS() for _ in range(S)]
    d_in = torch.empty((S, h, w, 3), dtype=torch.uint8, device="cuda")
    d_fg = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((S, h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    with ThreadPoolExecutor(S) as ex:
        for t in range(n):
            d_in.copy_(d[:, t % nres])
            p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
            exp = list(ex.map(lambda s: os_[s].process(host[s, t % nres]), range(S)))
            fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
            for s in range(S):
                assert np.array_equal(fg[s], exp[s][0]), "mask, stream %d frame %d" % (s, t)
                assert np.array_equal(bg[s], exp[s][1]), "background, stream %d frame %d" % (s, t)
    assert tb.kernel_launch_count() - before == n
    for s in range(S):
        assert_state_equal(p, os_[s], s)
    p.close()


@pytest.mark.parametrize("bg_last_only", [False, True])
def test_fused_kernel_1080p_T16_64_frames(oracle, bg_last_only):
    """mog2_fused_kernel at 1080p: one stream, four batches of T = 16 over a 32-frame ring, with a background image per
    frame and with bg_last_only (only frame T-1's model image is written)."""
    import torch
    import tracking_b200 as tb
    w, h, nres, T, nb = 1920, 1080, 32, 16, 4
    d, host = synth_resident(1, nres, w, h)
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    d_fg = torch.zeros((T, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((1 if bg_last_only else T, h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    for b in range(nb):
        t0 = (b * T) % nres
        p.process_batch_dev(d[0, t0:t0 + T].data_ptr(), T, w, h, d_fg.data_ptr(), d_bg.data_ptr(),
                            bg_last_only=bg_last_only)
        fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
        for t in range(T):
            ofg, obg = o.process(host[0, t0 + t])
            assert np.array_equal(fg[t], ofg), "mask, batch %d frame %d" % (b, t)
            if not bg_last_only:
                assert np.array_equal(bg[t], obg), "background, batch %d frame %d" % (b, t)
        if bg_last_only:
            assert np.array_equal(bg[0], obg), "last background, batch %d" % b
    assert tb.kernel_launch_count() - before == nb           # ONE launch per batch: the temporal-fusion kernel
    assert_state_equal(p, o)
    p.close()


def test_fused_kernel_2160p_T8_and_bg_last_only(oracle):
    """mog2_fused_kernel at the config-5 geometry (3840 x 2160): T = 8 with per-frame background images, then T = 8 with
    bg_last_only, on one stream; and a 2-stream group with T = 3 (groups always take the fused kernel for T > 1)."""
    import torch
    import tracking_b200 as tb
    w, h, T = 3840, 2160, 8
    d, host = synth_resident(1, 2 * T, w, h, t0=5)
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    d_fg = torch.zeros((T, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((T, h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    p.process_batch_dev(d[0, :T].data_ptr(), T, w, h, d_fg.data_ptr(), d_bg.data_ptr())
    fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
    for t in range(T):
        ofg, obg = o.process(host[0, t])
        assert np.array_equal(fg[t], ofg) and np.array_equal(bg[t], obg), t
    p.process_batch_dev(d[0, T:].data_ptr(), T, w, h, d_fg.data_ptr(), d_bg[0].data_ptr(), bg_last_only=True)
    fg, bg = d_fg.cpu().numpy(), d_bg[0].cpu().numpy()
    for t in range(T):
        ofg, obg = o.process(host[0, T + t])
        assert np.array_equal(fg[t], ofg), T + t
    assert np.array_equal(bg, obg)
    assert tb.kernel_launch_count() - before == 2
    assert_state_equal(p, o)
    p.close()
    del d_fg, d_bg
    # 2-stream group, T = 3
    S, T = 2, 3
    d2, host2 = synth_resident(S, T, w, h, t0=1)
    q = tb.MixtureOfGaussianV2BGS(nstreams=S)
    os_ = [oracle.MixtureOfGaussianV2BGS() for _ in range(S)]
    g_fg = torch.zeros((S, T, h, w), dtype=torch.uint8, device="cuda")
    g_bg = torch.zeros((S, h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    q.process_batch_dev(d2.data_ptr(), T, w, h, g_fg.data_ptr(), g_bg.data_ptr(), bg_last_only=True)
    assert tb.kernel_launch_count() - before == 1
    fg, bg = g_fg.cpu().numpy(), g_bg.cpu().numpy()
    for s in range(S):
        for t in range(T):
            ofg, obg = os_[s].process(host2[s, t])
            assert np.array_equal(fg[s, t], ofg), (s, t)
        assert np.array_equal(bg[s], obg), s
    q.close()


@pytest.mark.parametrize("mode", ["t1", "group_t1", "fused"])
def test_mode_churn_stress_512x512(oracle, mode):
    """SURVEY A.4 mode-churn stress (7-colour palette, 15 % switch probability) at 512 x 512: 1024 CTAs take the generic
    queue at once in every frame.  t1: single-stream T == 1 kernel; group_t1: 2-stream group kernel; fused: T = 8."""
    import torch
    import tracking_b200 as tb
    h = w = 512
    n = 48
    seqs = [stress_sequence(n, h, w, seed=11), stress_sequence(n, h, w, seed=12)]
    S = 2 if mode == "group_t1" else 1
    p = tb.MixtureOfGaussianV2BGS(nstreams=S, enableThreshold=0)          # raw {0,127,255}: the shadow test runs too
    os_ = [oracle.MixtureOfGaussianV2BGS(enableThreshold=False) for _ in range(S)]
    T = 8 if mode == "fused" else 1
    for t0 in range(0, n, T):
        host = np.stack([np.stack(seqs[s][t0:t0 + T]) for s in range(S)])          # S,T,h,w,3
        d_in = torch.from_numpy(host).cuda()
        d_fg = torch.zeros((S, T, h, w), dtype=torch.uint8, device="cuda")
        d_bg = torch.zeros((S, T, h, w, 3), dtype=torch.uint8, device="cuda")
        before = tb.kernel_launch_count()
        p.process_batch_dev(d_in.data_ptr(), T, w, h, d_fg.data_ptr(), d_bg.data_ptr())
        assert tb.kernel_launch_count() - before == 1
        fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
        for s in range(S):
            for t in range(T):
                ofg, obg = os_[s].process(seqs[s][t0 + t])
                assert np.array_equal(fg[s, t], ofg), (mode, s, t0 + t)
                assert np.array_equal(bg[s, t], obg), (mode, s, t0 + t)
    for s in range(S):
        assert_state_equal(p, os_[s], s)
        assert float(os_[s].nmodes.mean()) > 3.0 and int((os_[s].nmodes == 5).sum()) > 0     # the dense case is exercised
    p.close()


def test_config4_shape_8x1080p_mog2_open_ccl_labels(oracle):
    """BASELINE config 4 at its own shape: MOG2 (group kernel) -> OPEN 3x3 -> batched labelling on 8 x 1080p streams for
    12 frames; the cleaned masks, the canonical label images and the component tables against the oracle."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import blobs
    S, w, h, n = 8, 1920, 1080, 12
    d, host = synth_resident(S, n, w, h)
    p = tb.MixtureOfGaussianV2BGS(nstreams=S)
    os_ = [oracle.MixtureOfGaussianV2BGS() for _ in range(S)]
    cc = blobs.ConnectedComponents(w, h, max_images=S)
    d_in = torch.empty((S, h, w, 3), dtype=torch.uint8, device="cuda")
    d_fg = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_clean = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_lab = torch.zeros((S, h, w), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def expect(s, t):
        ofg, _ = os_[s].process(host[s, t])
        clean = oracle.morph(oracle.morph(ofg, "erode", 1), "dilate", 1)
        return (clean,) + tuple(oracle.ccl8(clean, True))

    with ThreadPoolExecutor(S) as ex:
        for t in range(n):
            d_in.copy_(d[:, t])
            p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None, stream=st)
            blobs.morph_dev(d_fg.data_ptr(), w, h, S, [("erode", 1), ("dilate", 1)], d_clean.data_ptr(), stream=st)
            cc.label_batch_dev(d_clean.data_ptr(), w, h, S, True, d_lab.data_ptr(), stream=st)
            exp = list(ex.map(lambda s: expect(s, t), range(S)))
            clean, lab = d_clean.cpu().numpy(), d_lab.cpu().numpy()
            for s in range(S):
                eclean, en, elab, est, eext = exp[s]
                assert np.array_equal(clean[s], eclean), (s, t)
                assert np.array_equal(lab[s], elab), (s, t)
                comps = cc.components(s)
                assert len(comps) == en
                for c, st_, e in zip(comps, est, eext):
                    assert (c["x"], c["y"], c["x"] + c["w"] - 1, c["y"] + c["h"] - 1, c["area"], c["first_index"]) == \
                        tuple(int(v) for v in st_)
                    assert c["external"] == int(e)
    cc.close()
    p.close()


@pytest.mark.parametrize("aid", [0, 2, 3])
def test_retain_input_no_history_writeback(oracle, clips, aid):
    """"retainInput": FD / WMV / WMM on the device path read their history from the caller's previous frame buffers
    (no write-back); same masks (and WMM background) as the default path and the oracle, one stream and a group."""
    import torch
    import tracking_b200 as tb
    clip = clips["video_clip"]
    n = 20
    h, w = clip.shape[1:3]
    for S in (1, 3):
        streams = [clip[:n], clip[::-1][:n], clip[4:4 + n]][:S]
        host = np.stack([np.stack([streams[s][t] for s in range(S)]) for t in range(n)])       # n,S,h,w,3
        d_all = torch.from_numpy(host).cuda()                                                 # every frame stays valid
        p = tb.ALGOS[aid](nstreams=S, retainInput=1)
        os_ = [oracle.ALGOS[aid]() for _ in range(S)]
        d_fg = torch.full((S, h, w), 9, dtype=torch.uint8, device="cuda")
        d_bg = torch.full((S, h, w, 3), 9, dtype=torch.uint8, device="cuda")
        before = tb.kernel_launch_count()
        for t in range(n):
            fv, bv = p.process_dev(d_all[t].data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
            fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
            for s in range(S):
                ofg, obg = os_[s].process(streams[s][t])
                assert fv == (ofg is not None)
                if ofg is not None:
                    assert np.array_equal(fg[s], ofg), (aid, S, s, t)
                if obg is not None:
                    assert bv and np.array_equal(bg[s], obg), (aid, S, s, t)
        warm = 1 if aid == 0 else 2
        extra = 1 if aid == 3 else 0          # WMV builds its quiet-group bound table once per context (one launch)
        assert tb.kernel_launch_count() - before == n - warm + extra      # warm-up frames are only remembered, not launched
        assert p.frame_count == n
        p.close()
    # a single-stream batch with retainInput: history = the tail of the previous batch
    p, o = tb.ALGOS[aid](retainInput=1), oracle.ALGOS[aid]()
    d_clip = torch.from_numpy(np.ascontiguousarray(clip[:24])).cuda()
    for t0 in range(0, 24, 4):
        d_fg = torch.full((4, h, w), 9, dtype=torch.uint8, device="cuda")
        first, _ = p.process_batch_dev(d_clip[t0:t0 + 4].data_ptr(), 4, w, h, d_fg.data_ptr(), None)
        fg = d_fg.cpu().numpy()
        for t in range(4):
            ofg, _ = o.process(clip[t0 + t])
            assert (ofg is None) == (t < first)
            if ofg is not None:
                assert np.array_equal(fg[t], ofg), (aid, t0 + t)
    p.close()


def test_abl_opencv24_fp32_blend_variant_all_byte_pairs(oracle):
    """ "ablBlend" 1 (OpenCV 2.4's fp32 addWeighted; parity UNPINNED -- no OpenCV 2.4 in this image):
    the new background byte for all 65 536 (input, background) pairs against the numpy fp32 restatement, and the
    documented fact that it differs from the pinned 4.x blend on some pairs."""
    import tracking_b200 as tb
    x = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)         # input byte per row
    y = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)           # background byte per column
    first = np.stack([y, y, y], -1)            # frame 0: bg <- in
    second = np.stack([x, x, x], -1)
    outs = {}
    for variant in (0, 1):
        p = tb.AdaptiveBackgroundLearning(ablBlend=variant)
        p.process(first)
        _, bg = p.process(second)
        outs[variant] = bg[..., 0].copy()
        assert np.array_equal(bg[..., 0], bg[..., 1]) and np.array_equal(bg[..., 0], bg[..., 2])
        p.close()
    exp24 = oracle.abl_blend_table_24(0.05)
    assert np.array_equal(outs[1], exp24)
    o = oracle.AdaptiveBackgroundLearning()
    o.process(first)
    _, obg = o.process(second)
    assert np.array_equal(outs[0], obg[..., 0])
    ndiff = int((outs[0] != outs[1]).sum())
    assert 0 < ndiff < 65536 * 0.1
    with pytest.raises(tb.BgsbError):
        tb.AdaptiveBackgroundLearning(ablTable=0, ablBlend=1)


@pytest.mark.parametrize("alpha", [0.05, 0.0, 1.0, 0.4, 0.004])
def test_abl_quiet_radius_shortcut_matches_plain_lookups(oracle, clips, alpha):
    """ABL's warp-coalesced table kernel skips the lookups of words whose bytes all lie within the table's quiet radius
    (read off the table on the device): same masks and models as with "quietGroups" 0 and as the oracle, on a noisy
    clip tiled to a multiple of 512 pixels (so the coalesced kernel runs), alpha from 0 (radius 255) to 1 (radius 0)."""
    import tracking_b200 as tb
    clip = clips["video_clip"]
    frames = [np.ascontiguousarray(np.tile(f, (2, 2, 1))[:128, :256]) for f in clip[:12]]
    assert frames[0].shape[0] * frames[0].shape[1] % 512 == 0
    a, b = tb.AdaptiveBackgroundLearning(alpha=alpha), tb.AdaptiveBackgroundLearning(alpha=alpha, quietGroups=0)
    o = oracle.AdaptiveBackgroundLearning(alpha=alpha)
    for i, f in enumerate(frames):
        fa, ba = a.process(f)
        fb, bb = b.process(f)
        fo, bo = o.process(f)
        assert np.array_equal(fa, fb) and np.array_equal(ba, bb), i
        assert np.array_equal(fa, fo) and np.array_equal(ba, bo), i
    a.close(); b.close()


@pytest.mark.parametrize("shape", [(64, 512), (48, 32), (96, 1024)])
@pytest.mark.parametrize("retain", [0, 1])
def test_wmv_bulk_copy_kernel_tiles(oracle, shape, retain):
    """Frames of whole 512-pixel tiles with the threshold on take WMV's bulk-copy kernel (persistent warps, cp.async.bulk
    prefetch; own history written back by bulk stores) once the two history frames exist: one stream and a group of three,
    caller-retained and library-owned history, both weightings, a threshold change mid-stream, noisy moving content so
    that quiet and busy groups mix.  (Ragged sizes -- the clips of the other tests -- take the per-thread kernel.)"""
    import torch
    import tracking_b200 as tb
    h, w = shape
    assert (h * w) % 512 == 0
    rng = np.random.default_rng(h + w + retain)
    n = 9

    def video(seed):
        r = np.random.default_rng(seed)
        base = r.integers(30, 200, (h, w, 3)).astype(np.int16)
        out = []
        for t in range(n):
            f = base + r.integers(-5, 6, (h, w, 3))
            x0 = (7 * t + seed) % max(1, w - 8)
            f[h // 4:h // 2, x0:x0 + 8] = r.integers(0, 256, 3)
            out.append(np.clip(f, 0, 255).astype(np.uint8))
        return out

    for S in (1, 3):
        for ew in (1, 0):
            vids = [video(10 * S + s + ew) for s in range(S)]
            host = np.stack([np.stack([vids[s][t] for s in range(S)]) for t in range(n)])          # n,S,h,w,3
            d_all = torch.from_numpy(host).cuda()
            p = tb.WeightedMovingVarianceBGS(nstreams=S, enableWeight=ew, **({"retainInput": 1} if retain else {}))
            os_ = [oracle.WeightedMovingVarianceBGS(enableWeight=bool(ew)) for _ in range(S)]
            d_fg = torch.full((S, h, w), 9, dtype=torch.uint8, device="cuda")
            for t in range(n):
                if t == 6:
                    p.set("threshold", 4)
                    for o in os_:
                        o.threshold = 4
                fv, _ = p.process_dev(d_all[t].data_ptr(), w, h, d_fg.data_ptr(), None)
                fg = d_fg.cpu().numpy()
                for s in range(S):
                    ofg, _ = os_[s].process(vids[s][t])
                    assert fv == (ofg is not None)
                    if ofg is not None:
                        assert np.array_equal(fg[s], ofg), (shape, retain, S, ew, s, t)
                        assert 0 < int((ofg > 0).sum()) < ofg.size or t < 3
            p.close()


@pytest.mark.parametrize("shape", [(64, 512), (48, 32), (96, 1024)])
@pytest.mark.parametrize("alpha", [0.05, 1.0, 0.0, 0.3])
def test_abl_bulk_copy_kernel_tiles(oracle, shape, alpha):
    """ABL's bulk-copy kernel ("ablTable" 3 forces it on frames this small; by default it takes groups of at least ~3.6
    Mpx): persistent warps, input and model tiles by cp.async.bulk, changed model pieces written back into the ring and
    stored by bulk copies.  One stream and a group of three, with and without a background image, raw and thresholded
    output, a threshold change mid-stream; alpha 0 (nothing ever changes), 1 (model = input) and in between.  The model
    is checked through the following frames and, at the end, by a call that asks for the background image."""
    import torch
    import tracking_b200 as tb
    h, w = shape
    assert (h * w) % 512 == 0
    n = 8

    def video(seed):
        r = np.random.default_rng(seed)
        base = r.integers(30, 200, (h, w, 3)).astype(np.int16)
        out = []
        for t in range(n):
            f = base + r.integers(-7, 8, (h, w, 3))
            x0 = (7 * t + seed) % max(1, w - 8)
            f[h // 4:h // 2, x0:x0 + 8] = r.integers(0, 256, 3)
            out.append(np.clip(f, 0, 255).astype(np.uint8))
        return out

    for S in (1, 3):
        for thr_on, want_bg in ((1, 1), (1, 0), (0, 1)):
            vids = [video(100 * S + 10 * thr_on + s + want_bg) for s in range(S)]
            host = np.stack([np.stack([vids[s][t] for s in range(S)]) for t in range(n)])          # n,S,h,w,3
            d_all = torch.from_numpy(host).cuda()
            p = tb.AdaptiveBackgroundLearning(nstreams=S, alpha=alpha, ablTable=3, enableThreshold=thr_on)
            os_ = [oracle.AdaptiveBackgroundLearning(alpha=alpha, enableThreshold=bool(thr_on)) for _ in range(S)]
            d_fg = torch.full((S, h, w), 9, dtype=torch.uint8, device="cuda")
            d_bg = torch.full((S, h, w, 3), 9, dtype=torch.uint8, device="cuda")
            for t in range(n):
                if t == 5:
                    p.set("threshold", 3)
                    for o in os_:
                        o.threshold = 3
                ask = want_bg or t == n - 1
                p.process_dev(d_all[t].data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr() if ask else None)
                fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
                for s in range(S):
                    ofg, obg = os_[s].process(vids[s][t])
                    assert np.array_equal(fg[s], ofg), (shape, alpha, S, thr_on, want_bg, s, t)
                    if ask:
                        assert np.array_equal(bg[s], obg), (shape, alpha, S, thr_on, want_bg, s, t)
            p.close()


def test_abl_and_wmv_bulk_kernels_at_bench_geometry(oracle):
    """The geometry tools/bench_configs.py times (1080p stream groups on the K-GEN video): with default parameters a
    group of 3 x 1080p streams is routed to abl_bulk_kernel (>= ~3.6 Mpx per launch) and wmv_bulk_kernel; 30 frames over a
    10-frame ring -- long enough for the rectangles' trails to sit at the blend table's quiet radius -- against
    independent oracles per stream: masks every frame, background images every frame (ABL)."""
    import torch
    import tracking_b200 as tb
    S, w, h, nres, n = 3, 1920, 1080, 10, 30
    d, host = synth_resident(S, nres, w, h)
    d_in = torch.empty((S, h, w, 3), dtype=torch.uint8, device="cuda")
    d_fg = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((S, h, w, 3), dtype=torch.uint8, device="cuda")
    for name in ("AdaptiveBackgroundLearning", "WeightedMovingVarianceBGS"):
        p = getattr(tb, name)(nstreams=S)
        os_ = [getattr(oracle, name)() for _ in range(S)]
        with ThreadPoolExecutor(max_workers=S) as ex:
            for t in range(n):
                d_in.copy_(d[:, t % nres])
                fv, bv = p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
                res = list(ex.map(lambda s: os_[s].process(host[s, t % nres]), range(S)))
                fg = d_fg.cpu().numpy()
                bg = d_bg.cpu().numpy() if bv else None
                for s in range(S):
                    ofg, obg = res[s]
                    assert fv == (ofg is not None), (name, t)
                    if ofg is not None:
                        assert np.array_equal(fg[s], ofg), (name, s, t)
                    if bv:
                        assert np.array_equal(bg[s], obg), (name, s, t)
        p.close()
