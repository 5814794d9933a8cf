"""Parity of the CUDA plugins against the oracle -- every call goes through the C ABI."""
import hashlib

import numpy as np
import pytest

from conftest import stress_sequence

pytestmark = pytest.mark.gpu

NAMES = {0: "FrameDifferenceBGS", 1: "StaticFrameDifferenceBGS", 2: "WeightedMovingMeanBGS",
         3: "WeightedMovingVarianceBGS", 5: "MixtureOfGaussianV2BGS", 6: "AdaptiveBackgroundLearning"}
# north_star tolerances: FD bit-exact; MOG2/ABL/WMV masks <= 0.1 % disagreement, bg <= 1e-4 relative.
# The kernels are in fact bit-exact on every fixture, so the tests assert 0 mismatches and would
# start failing long before the contractual tolerance is reached.
MASK_TOL = 0.0
BG_TOL = 0.0


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_host(plugin, frames):
    fgs, bgs = [], []
    for f in frames:
        fg, bg = plugin.process(f)
        if fg is not None:
            fgs.append(fg.copy())
        if bg is not None:
            bgs.append(bg.copy())
    return fgs, bgs


def assert_same(a, b, what):
    assert len(a) == len(b), "%s: %d vs %d outputs" % (what, len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        bad = int((x != y).sum())
        assert bad <= MASK_TOL * x.size, "%s frame %d: %d / %d differ" % (what, i, bad, x.size)


@pytest.mark.parametrize("seq", ["video_clip", "png_clip"])
@pytest.mark.parametrize("aid", [0, 1, 2, 3, 5, 6])
@pytest.mark.parametrize("thr", [True, False])
def test_host_path_matches_golden_and_oracle(oracle, clips, golden, seq, aid, thr):
    import tracking_b200 as tb
    frames = list(clips[seq])
    p = tb.ALGOS[aid](enableThreshold=int(thr))
    fgs, bgs = run_host(p, frames)
    ofg, obg = [], []
    o = oracle.ALGOS[aid](enableThreshold=thr)
    for f in frames:
        fg, bg = o.process(f)
        if fg is not None:
            ofg.append(fg)
        if bg is not None:
            obg.append(bg)
    assert_same(fgs, ofg, "fg")
    assert_same(bgs, obg, "bg")
    exp = golden["sequences"][seq]["algos"][NAMES[aid] + ("" if thr else ":enableThreshold=0")]
    assert sha(fgs) == exp["fg_sha256"]
    if exp["bg_sha256"]:
        assert sha(bgs) == exp["bg_sha256"]
    p.close()


def test_warmup_contract_outputs_untouched():
    """FD frame 0 and WMV frames 0-1 return without writing (reference early returns)."""
    import tracking_b200 as tb
    f = np.zeros((40, 50, 3), np.uint8)
    fd, wmv, abl, mog = tb.FrameDifferenceBGS(), tb.WeightedMovingVarianceBGS(), tb.AdaptiveBackgroundLearning(), tb.MixtureOfGaussianV2BGS()
    assert fd.process(f) == (None, None)
    fg, bg = fd.process(f)
    assert fg is not None and bg is None
    assert wmv.process(f) == (None, None) and wmv.process(f) == (None, None)
    fg, bg = wmv.process(f)
    assert fg is not None and bg is None
    fg, bg = abl.process(f)
    assert fg is not None and bg is not None and not fg.any()
    fg, bg = mog.process(f)
    assert (fg == 255).all() and np.array_equal(bg, f)          # frame 1: all foreground, bg = frame
    assert fd.process(None) == (None, None) and fd.process(np.zeros((0, 0, 3), np.uint8)) == (None, None)
    assert fd.frame_count == 2


@pytest.mark.parametrize("variant", [0, 1])
def test_mog2_stress_sequence_state_bit_exact(oracle, variant):
    """Mode churn (prune / replace / re-sort paths) incl. the exported mixture state."""
    import tracking_b200 as tb
    frames = stress_sequence(200, 40, 52)
    for thr, kw in ((True, {}), (False, {}), (True, {"threshold": 200})):
        p = tb.MixtureOfGaussianV2BGS(enableThreshold=int(thr), kernelVariant=variant, **kw)
        o = oracle.MixtureOfGaussianV2BGS(enableThreshold=thr, **kw)
        for i, f in enumerate(frames):
            fg, bg = p.process(f)
            ofg, obg = o.process(f)
            assert np.array_equal(fg, ofg), "mask frame %d" % i
            assert np.array_equal(bg, obg), "bg frame %d" % i
        planes, nm = p.export_state()
        assert np.array_equal(nm, o.nmodes)
        K = 5
        for m in range(K):
            live = o.nmodes > m
            assert np.array_equal(planes[m * 5 + 0][live], o.gmm[:, m, 0][live])
            assert np.array_equal(planes[m * 5 + 1][live], o.gmm[:, m, 1][live])
            for c in range(3):
                assert np.array_equal(planes[m * 5 + 2 + c][live], o.mean[:, m, c][live])
        p.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_mog2_variants_on_reference_clip_temporal_batches(oracle, clips, variant):
    """Both kernels, T in {1, 5, 32}, raw {0,127,255} output (shadow path) on the reference video clip."""
    import torch
    import tracking_b200 as tb
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    for T in (1, 5, 32):
        p = tb.MixtureOfGaussianV2BGS(kernelVariant=variant, enableThreshold=0)
        o = oracle.MixtureOfGaussianV2BGS(enableThreshold=False)
        for t0 in range(0, 30, T):
            tt = min(T, 32 - t0)
            d_in = torch.from_numpy(clip[t0:t0 + tt]).cuda()
            d_fg = torch.zeros((tt, h, w), dtype=torch.uint8, device="cuda")
            d_bg = torch.zeros((tt, h, w, 3), dtype=torch.uint8, device="cuda")
            p.process_batch_dev(d_in.data_ptr(), tt, w, h, d_fg.data_ptr(), d_bg.data_ptr())
            torch.cuda.synchronize()
            for t in range(tt):
                ofg, obg = o.process(clip[t0 + t])
                assert np.array_equal(d_fg[t].cpu().numpy(), ofg), (T, t0 + t)
                assert np.array_equal(d_bg[t].cpu().numpy(), obg), (T, t0 + t)
        planes, nm = p.export_state()
        assert np.array_equal(nm, o.nmodes)
        for m in range(5):
            live = o.nmodes > m
            assert np.array_equal(planes[m * 5][live], o.gmm[:, m, 0][live])
            assert np.array_equal(planes[m * 5 + 1][live], o.gmm[:, m, 1][live])
            for c in range(3):
                assert np.array_equal(planes[m * 5 + 2 + c][live], o.mean[:, m, c][live])
        p.close()


def test_mog2_auto_learning_rate_and_params(oracle):
    """alpha < 0 -> 1/min(2*nframes, history); non-default MOG2 properties."""
    import tracking_b200 as tb
    frames = stress_sequence(60, 33, 47, seed=11)
    p = tb.MixtureOfGaussianV2BGS(alpha=-1, history=40, varThreshold=10, backgroundRatio=0.7)
    o = oracle.MixtureOfGaussianV2BGS(alpha=-1)
    o.params.history = 40; o.params.Tb = 10; o.params.TB = 0.7
    for i, f in enumerate(frames):
        fg, bg = p.process(f)
        ofg, obg = o.process(f)
        assert np.array_equal(fg, ofg) and np.array_equal(bg, obg), i


@pytest.mark.parametrize("table", [1, 2, 0, 3])
def test_abl_exhaustive_byte_pairs(oracle, table):
    """Every (input, background) byte pair, lookup-table kernel and arithmetic kernel; alpha changes mid-stream
    (the table is rebuilt)."""
    import tracking_b200 as tb
    inp, bg = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    f0 = np.repeat(bg[:, :, None], 3, 2).copy()
    f1 = np.repeat(inp[:, :, None], 3, 2).copy()
    for alpha in (0.05, 0.3, 0.001):
        p = tb.AdaptiveBackgroundLearning(alpha=alpha, ablTable=table)
        o = oracle.AdaptiveBackgroundLearning(alpha=alpha)
        for f in (f0, f1):
            fa, ba = p.process(f)
            fb, bb = o.process(f)
        assert np.array_equal(fa, fb) and np.array_equal(ba, bb)
        p.set("alpha", 0.5); o.alpha = 0.5
        fa, ba = p.process(f0)
        fb, bb = o.process(f0)
        assert np.array_equal(fa, fb) and np.array_equal(ba, bb)


@pytest.mark.parametrize("shape", [(64, 512), (96, 1024), (37, 53)])
def test_fd_coalesced_and_per_thread_kernels(oracle, shape):
    """FrameDifference: frames whose pixel count is a multiple of 512 take the warp-coalesced kernel (device path,
    single stream, two-stream group, temporal batch), anything else the per-thread one; raw and thresholded."""
    import torch
    import tracking_b200 as tb
    h, w = shape
    rng = np.random.default_rng(17)
    frames = [rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8) for _ in range(7)]
    for thr in (1, 0):
        g = tb.FrameDifferenceBGS(nstreams=2, enableThreshold=thr)
        oa, ob = oracle.FrameDifferenceBGS(enableThreshold=bool(thr)), oracle.FrameDifferenceBGS(enableThreshold=bool(thr))
        d_fg = torch.zeros((2, h, w), dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames[:4]):
            d_in = torch.from_numpy(f).cuda()
            fv, _ = g.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None)
            torch.cuda.synchronize()
            ra, rb = oa.process(f[0])[0], ob.process(f[1])[0]
            assert fv == (ra is not None), i
            if ra is not None:
                out = d_fg.cpu().numpy()
                assert np.array_equal(out[0], ra) and np.array_equal(out[1], rb), i
        # temporal batch of 3 frames per stream: [S][T][h][w][3]
        batch = np.stack([np.stack([frames[4 + t][s_] for t in range(3)]) for s_ in range(2)])
        d_b = torch.from_numpy(batch).cuda()
        d_fgb = torch.zeros((2, 3, h, w), dtype=torch.uint8, device="cuda")
        first, _ = g.process_batch_dev(d_b.data_ptr(), 3, w, h, d_fgb.data_ptr())
        torch.cuda.synchronize()
        assert first == 0
        out = d_fgb.cpu().numpy()
        for t in range(3):
            assert np.array_equal(out[0, t], oa.process(frames[4 + t][0])[0]), t
            assert np.array_equal(out[1, t], ob.process(frames[4 + t][1])[0]), t
        g.close()


def test_wmm_unweighted_and_raw(oracle):
    """WeightedMovingMean sibling plugin: (x0+x1+x2)/3.0 branch and un-thresholded output."""
    import tracking_b200 as tb
    rng = np.random.default_rng(12)
    frames = [rng.integers(0, 256, (131, 177, 3), dtype=np.uint8) for _ in range(6)]
    for ew, thr in ((1, 1), (0, 1), (1, 0), (0, 0)):
        p = tb.WeightedMovingMeanBGS(enableWeight=ew, enableThreshold=thr)
        o = oracle.WeightedMovingMeanBGS(enableWeight=bool(ew), enableThreshold=bool(thr))
        for f in frames:
            fa, ba = p.process(f)
            fb, bb = o.process(f)
            assert (fa is None) == (fb is None) and (ba is None) == (bb is None)
            if fa is not None:
                assert np.array_equal(fa, fb) and np.array_equal(ba, bb)


def test_wmv_small_differences_and_weight_change(oracle):
    """Frames that differ by camera-noise amounts (many exact rounding ties: equal previous frames and an odd
    difference give 255*sd = k + 0.5) with a moving high-contrast patch, both weightings, raw and thresholded
    output, and a weighting change mid-stream."""
    import tracking_b200 as tb
    rng = np.random.default_rng(21)
    h, w = 203, 317
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames = []
    for t in range(9):
        f = np.clip(base.astype(np.int16) + rng.integers(-14, 15, (h, w, 3)), 0, 255).astype(np.uint8)
        f[20 + 9 * t:60 + 9 * t, 30 + 11 * t:90 + 11 * t] = (255 * (t & 1), 128, 255 - 255 * (t & 1))
        frames.append(f)
    for ew in (1, 0):
        for thr in (1, 0):
            p = tb.WeightedMovingVarianceBGS(enableWeight=ew, enableThreshold=thr)
            o = oracle.WeightedMovingVarianceBGS(enableWeight=bool(ew), enableThreshold=bool(thr))
            for i, f in enumerate(frames):
                if i == 6:
                    p.set("enableWeight", 1 - ew); o.enableWeight = not bool(ew)
                fa, _ = p.process(f)
                fb, _ = o.process(f)
                assert (fa is None) == (fb is None)
                if fa is not None:
                    assert np.array_equal(fa, fb), (ew, thr, i)
            p.close()


@pytest.mark.parametrize("ew", [1, 0])
def test_wmv_all_byte_triples(oracle, ew):
    """Every (current, previous, previous-previous) byte triple, both weightings: 4096 x 4096 pixels whose three
    channels carry the same triple, raw output (gray(v, v, v) == v, so the mask byte IS the per-channel value)."""
    import torch
    import tracking_b200 as tb
    n = 4096
    idx = np.arange(n * n, dtype=np.uint32).reshape(n, n)
    planes = [(idx & 0xff).astype(np.uint8), ((idx >> 8) & 0xff).astype(np.uint8), ((idx >> 16) & 0xff).astype(np.uint8)]
    frames = [np.repeat(pl[:, :, None], 3, 2) for pl in (planes[2], planes[1], planes[0])]      # oldest first
    p = tb.WeightedMovingVarianceBGS(enableWeight=ew, enableThreshold=0)
    o = oracle.WeightedMovingVarianceBGS(enableWeight=bool(ew), enableThreshold=False)
    d_fg = torch.zeros((n, n), dtype=torch.uint8, device="cuda")
    for f in frames:
        d_in = torch.from_numpy(f).cuda()
        fv, _ = p.process_dev(d_in.data_ptr(), n, n, d_fg.data_ptr(), None)
        torch.cuda.synchronize()
        ofg, _ = o.process(f)
    assert fv and ofg is not None
    got = d_fg.cpu().numpy()
    bad = int((got != ofg).sum())
    assert bad == 0, "%d of 16.7 M triples differ" % bad
    p.close()


@pytest.mark.parametrize("ew", [1, 0])
def test_wmm_all_byte_triples(oracle, ew):
    """WeightedMovingMean: every byte triple, both weightings; the background image carries the per-channel mean."""
    import torch
    import tracking_b200 as tb
    n = 4096
    idx = np.arange(n * n, dtype=np.uint32).reshape(n, n)
    planes = [(idx & 0xff).astype(np.uint8), ((idx >> 8) & 0xff).astype(np.uint8), ((idx >> 16) & 0xff).astype(np.uint8)]
    frames = [np.repeat(pl[:, :, None], 3, 2) for pl in (planes[2], planes[1], planes[0])]
    p = tb.WeightedMovingMeanBGS(enableWeight=ew, enableThreshold=0)
    o = oracle.WeightedMovingMeanBGS(enableWeight=bool(ew), enableThreshold=False)
    d_fg = torch.zeros((n, n), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((n, n, 3), dtype=torch.uint8, device="cuda")
    for f in frames:
        d_in = torch.from_numpy(f).cuda()
        fv, bv = p.process_dev(d_in.data_ptr(), n, n, d_fg.data_ptr(), d_bg.data_ptr())
        torch.cuda.synchronize()
        ofg, obg = o.process(f)
    assert fv and bv and obg is not None
    assert int((d_bg.cpu().numpy() != obg).sum()) == 0 and int((d_fg.cpu().numpy() != ofg).sum()) == 0
    p.close()


def test_wmv_exhaustive_triples_sample(oracle):
    """Random byte triples incl. unweighted variant and raw (un-thresholded) output."""
    import tracking_b200 as tb
    rng = np.random.default_rng(2)
    frames = [rng.integers(0, 256, (257, 301, 3), dtype=np.uint8) for _ in range(6)]
    for ew in (1, 0):
        for thr in (1, 0):
            p = tb.WeightedMovingVarianceBGS(enableWeight=ew, enableThreshold=thr)
            o = oracle.WeightedMovingVarianceBGS(enableWeight=bool(ew), enableThreshold=bool(thr))
            for f in frames:
                fa, _ = p.process(f)
                fb, _ = o.process(f)
                assert (fa is None) == (fb is None)
                if fa is not None:
                    assert np.array_equal(fa, fb)


@pytest.mark.parametrize("aid", [0, 1, 2, 3, 5, 6])
def test_gray_variant_24_constants(oracle, clips, aid):
    import tracking_b200 as tb
    if aid == 5:
        pytest.skip("MOG2 has no gray conversion")
    frames = list(clips["png_clip"])
    p = tb.ALGOS[aid](grayVariant=1)
    o = oracle.ALGOS[aid](gray_variant=1)
    a, _ = run_host(p, frames)
    b = [x for x in (o.process(f)[0] for f in frames) if x is not None]
    assert_same(a, b, "gray 2.4")


@pytest.mark.parametrize("aid", [0, 1, 2, 3, 5, 6])
@pytest.mark.parametrize("T", [1, 3, 8])
def test_device_batch_and_stream_group(oracle, clips, aid, T):
    """Temporal batches of T frames and a group of 3 streams advanced by one launch equal the
    frame-by-frame oracle (state carried in registers across the batch must not change results)."""
    import torch
    import tracking_b200 as tb
    clip = clips["video_clip"]
    S, n = 3, 24
    # three different streams: the clip, the clip reversed, the clip shifted by 5 frames
    streams = [clip[:n], clip[::-1][:n], clip[5:5 + n]]
    h, w = clip.shape[1:3]
    p = tb.ALGOS[aid](nstreams=S)
    oracles = [oracle.ALGOS[aid]() for _ in range(S)]
    has_bg = aid in (1, 2, 5, 6)
    for t0 in range(0, n, T):
        host = np.stack([np.stack([streams[s][t0 + t] for t in range(T)]) for s in range(S)])   # S,T,h,w,3
        d_in = torch.from_numpy(host).cuda()
        d_fg = torch.full((S, T, h, w), 77, dtype=torch.uint8, device="cuda")
        d_bg = torch.full((S, T, h, w, 3), 77, dtype=torch.uint8, device="cuda")
        first, bgv = p.process_batch_dev(d_in.data_ptr(), T, w, h, d_fg.data_ptr(), d_bg.data_ptr(),
                                         stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        fg, bg = d_fg.cpu().numpy(), d_bg.cpu().numpy()
        assert bgv == (has_bg and first < T)
        for s in range(S):
            for t in range(T):
                ofg, obg = oracles[s].process(streams[s][t0 + t])
                if ofg is None:
                    assert t < first and (fg[s, t] == 77).all()          # untouched
                else:
                    assert t >= first
                    assert np.array_equal(fg[s, t], ofg), (s, t0 + t)
                if obg is not None:
                    assert np.array_equal(bg[s, t], obg), (s, t0 + t)
    p.close()


def test_bg_last_only_and_no_bg(oracle, clips):
    import torch
    import tracking_b200 as tb
    clip = clips["png_clip"]
    h, w = clip.shape[1:3]
    T = 4
    for aid in (5, 6):
        p, o = tb.ALGOS[aid](), oracle.ALGOS[aid]()
        q = tb.ALGOS[aid]()
        for t0 in range(0, 16, T):
            d_in = torch.from_numpy(clip[t0:t0 + T]).cuda()
            d_fg = torch.zeros((T, h, w), dtype=torch.uint8, device="cuda")
            d_bg = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
            p.process_batch_dev(d_in.data_ptr(), T, w, h, d_fg.data_ptr(), d_bg.data_ptr(), bg_last_only=True)
            d_fg2 = torch.zeros((T, h, w), dtype=torch.uint8, device="cuda")
            q.process_batch_dev(d_in.data_ptr(), T, w, h, d_fg2.data_ptr(), None)
            torch.cuda.synchronize()
            for t in range(T):
                ofg, obg = o.process(clip[t0 + t])
                assert np.array_equal(d_fg[t].cpu().numpy(), ofg)
                assert np.array_equal(d_fg2[t].cpu().numpy(), ofg)
            assert np.array_equal(d_bg.cpu().numpy(), obg)


def test_strided_host_input_and_reset(oracle):
    """cv::Mat rows from cvQueryFrame are 4-byte aligned (padded stride); reset() restarts the model."""
    import tracking_b200 as tb
    rng = np.random.default_rng(4)
    h, w = 37, 101
    padded = rng.integers(0, 256, (5, h, w * 3 + 5), dtype=np.uint8)
    frames = [np.lib.stride_tricks.as_strided(padded[i], (h, w, 3), (w * 3 + 5, 3, 1)) for i in range(5)]
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    for f in frames:
        fg, bg = p.process(f)
        ofg, obg = o.process(np.ascontiguousarray(f))
        assert np.array_equal(fg, ofg) and np.array_equal(bg, obg)
    p.reset()
    o = oracle.MixtureOfGaussianV2BGS()
    for f in frames[:2]:
        fg, bg = p.process(f)
        ofg, obg = o.process(np.ascontiguousarray(f))
        assert np.array_equal(fg, ofg) and np.array_equal(bg, obg)
    # geometry change re-initialises (cv::BackgroundSubtractorMOG2::operator() does the same)
    f2 = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    fg, bg = p.process(f2)
    assert (fg == 255).all() and np.array_equal(bg, f2)


@pytest.mark.parametrize("aid", [0, 2, 5, 6])
def test_host_path_band_pipeline_large_frames(oracle, aid):
    """bgsb_process on frames >= 1 MB cuts the image into row bands (upload / kernel / download overlap);
    results must not depend on the band count (incl. history kept in the upload ring for FD / WMM)."""
    import tracking_b200 as tb
    from tracking_b200 import synth
    w, h = 1280, 720
    frames = [synth.frame(w, h, t) for t in range(5)]
    ref = None
    for bands in (1, 4, 7):
        p = tb.ALGOS[aid](hostBands=bands)
        outs = [p.process(f) for f in frames]
        if ref is None:
            o = oracle.ALGOS[aid]()
            ref = [o.process(f) for f in frames]
        for (fa, ba), (fb, bb) in zip(outs, ref):
            assert (fa is None) == (fb is None) and (ba is None) == (bb is None)
            if fa is not None:
                assert np.array_equal(fa, fb)
            if ba is not None:
                assert np.array_equal(ba, bb)
        p.close()


@pytest.mark.parametrize("shape", [(97, 131), (600, 700)])
def test_fanout_one_upload_matches_separate_plugins(oracle, shape):
    """FrameProcessor-style fan-out (FrameProcessor.cpp:169-215): FD, StaticFD, WMM, WMV, MOG2, ABL on one uploaded
    frame == the oracle's plugins run one by one, incl. warm-up frames; (600, 700) takes the row-band pipeline.
    Afterwards every context continues correctly through its own process()."""
    import tracking_b200 as tb
    h, w = shape
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames = [np.clip(base.astype(np.int16) + rng.integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8) for _ in range(7)]
    frames[4][h // 4:h // 2, w // 4:w // 2] = 255 - frames[4][h // 4:h // 2, w // 4:w // 2]
    names = ["FrameDifferenceBGS", "StaticFrameDifferenceBGS", "WeightedMovingMeanBGS", "WeightedMovingVarianceBGS",
             "MixtureOfGaussianV2BGS", "AdaptiveBackgroundLearning"]
    ps = [getattr(tb, nm)() for nm in names]
    os_ = [getattr(oracle, nm)() for nm in names]
    for i, f in enumerate(frames[:6]):
        outs = tb.process_fanout(ps, f)
        for nm, (fa, ba), o in zip(names, outs, os_):
            fb, bb = o.process(f)
            assert (fa is None) == (fb is None), (nm, i)
            assert (ba is None) == (bb is None), (nm, i)
            if fa is not None:
                assert np.array_equal(fa, fb), (nm, i)
            if ba is not None:
                assert np.array_equal(ba, bb), (nm, i)
    for nm, p, o in zip(names, ps, os_):               # the per-context state is the ordinary one
        fa, ba = p.process(frames[6])
        fb, bb = o.process(frames[6])
        assert np.array_equal(fa, fb), nm
        if bb is not None:
            assert np.array_equal(ba, bb), nm
        p.close()
    with pytest.raises(tb.BgsbError):
        q = tb.FrameDifferenceBGS()
        tb.process_fanout([q, q], frames[0])


@pytest.mark.parametrize("pinned", [True, False])
def test_submit_wait_pipeline_matches_oracle(oracle, pinned):
    """Pipelined ingest (bgsb_submit / bgsb_wait): nine queued frames per plugin give exactly the
    outputs the oracle's plugins give frame by frame -- warm-up frames leave their buffers untouched -- and the model
    continues correctly through a synchronous process() (which drains the queue first) and through reset()."""
    import tracking_b200 as tb
    h, w = 203, 311
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames = [np.clip(base.astype(np.int16) + rng.integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8) for _ in range(10)]
    frames[5][h // 4:h // 2, w // 4:w // 2] = 255 - frames[5][h // 4:h // 2, w // 4:w // 2]
    alloc = tb.pinned_empty if pinned else (lambda shape: np.empty(shape, np.uint8))
    names = ["FrameDifferenceBGS", "StaticFrameDifferenceBGS", "WeightedMovingMeanBGS", "WeightedMovingVarianceBGS",
             "MixtureOfGaussianV2BGS", "AdaptiveBackgroundLearning", "AdaptiveSelectiveBackgroundLearning",
             "DPZivkovicAGMMBGS", "DPAdaptiveMedianBGS", "DPMeanBGS", "DPWrenGABGS", "DPPratiMediodBGS", "SigmaDeltaBGS"]
    for nm in names:
        p, o = getattr(tb, nm)(), getattr(oracle, nm)()
        bgshape = (h, w, 3) if p.BG_CHANNELS == 3 else (h, w)
        ins, fgs, bgs, flags = [], [], [], []
        for f in frames[:9]:
            a = alloc((h, w, 3)); a[...] = f
            fg = alloc((h, w)); fg[...] = 77
            bg = alloc(bgshape); bg[...] = 77
            flags.append(p.submit(a, fg, bg))
            ins.append(a); fgs.append(fg); bgs.append(bg)
        p.wait()
        for i, f in enumerate(frames[:9]):
            fb, bb = o.process(f)
            assert flags[i] == (fb is not None, bb is not None), (nm, i)
            assert np.array_equal(fgs[i], fb) if fb is not None else (fgs[i] == 77).all(), (nm, i)
            assert np.array_equal(bgs[i], bb) if bb is not None else (bgs[i] == 77).all(), (nm, i)
        # queue two more and go straight into a synchronous call: it waits for the queue
        a = alloc((h, w, 3)); a[...] = frames[9]
        fg = alloc((h, w)); bg = alloc(bgshape)
        p.submit(a, fg, bg)
        fa, ba = p.process(frames[3])
        fb, bb = o.process(frames[9])
        assert np.array_equal(fg, fb), nm
        fb, bb = o.process(frames[3])
        assert np.array_equal(fa, fb), nm
        if bb is not None:
            assert np.array_equal(ba, bb), nm
        p.wait()                                         # nothing queued: returns at once
        p.close()


@pytest.mark.parametrize("shape", [(1, 1), (1, 67), (3, 2), (2, 33), (65, 1), (5, 129)])
def test_tiny_and_ragged_frames_all_plugins(oracle, shape):
    """Degenerate geometries (one pixel, one row, one column, widths that are not multiples of any vector or tile
    size): every plugin, synchronous and queued host calls, against the oracle."""
    import tracking_b200 as tb
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames = [np.clip(base.astype(np.int16) + rng.integers(-20, 21, (h, w, 3)), 0, 255).astype(np.uint8) for _ in range(6)]
    frames[3] = 255 - frames[3]
    for nm in ["FrameDifferenceBGS", "StaticFrameDifferenceBGS", "WeightedMovingMeanBGS", "WeightedMovingVarianceBGS",
               "MixtureOfGaussianV2BGS", "AdaptiveBackgroundLearning", "AdaptiveSelectiveBackgroundLearning",
               "DPZivkovicAGMMBGS", "DPAdaptiveMedianBGS", "DPMeanBGS", "DPWrenGABGS", "DPPratiMediodBGS", "SigmaDeltaBGS"]:
        p, q, o = getattr(tb, nm)(), getattr(tb, nm)(), getattr(oracle, nm)()
        bgshape = (h, w, 3) if q.BG_CHANNELS == 3 else (h, w)
        outs = []
        for f in frames:
            fg, bg = np.full((h, w), 77, np.uint8), np.full(bgshape, 77, np.uint8)
            q.submit(np.ascontiguousarray(f), fg, bg)
            outs.append((fg, bg))
        q.wait()
        for i, f in enumerate(frames):
            fa, ba = p.process(f)
            fb, bb = o.process(f)
            assert (fa is None) == (fb is None) and (ba is None) == (bb is None), (nm, i)
            if fb is not None:
                assert np.array_equal(fa, fb) and np.array_equal(outs[i][0], fb), (nm, i)
            if bb is not None:
                assert np.array_equal(ba, bb) and np.array_equal(outs[i][1], bb), (nm, i)
        p.close(); q.close()


def test_submit_geometry_change_and_errors():
    """A new frame size re-initialises the model exactly as process() does; bad buffers are refused."""
    import tracking_b200 as tb
    rng = np.random.default_rng(3)
    p, q = tb.MixtureOfGaussianV2BGS(), tb.MixtureOfGaussianV2BGS()
    for (h, w) in [(64, 80), (64, 80), (50, 33), (50, 33)]:
        f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        fg, bg = np.empty((h, w), np.uint8), np.empty((h, w, 3), np.uint8)
        p.submit(f, fg, bg)
        p.wait()
        fb, bb = q.process(f)
        assert np.array_equal(fg, fb) and np.array_equal(bg, bb)
    with pytest.raises(ValueError):
        p.submit(f, np.empty((5, 5), np.uint8))
    with pytest.raises(ValueError):
        p.submit(f[:, ::2], np.empty((50, 17), np.uint8))
    p.close(); q.close()


def _asbl_frames(h, w, n, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frames = []
    for t in range(n):
        f = np.clip(base.astype(np.int16) + rng.integers(-30, 31, (h, w, 3)), 0, 255).astype(np.uint8)
        f[10 + 4 * t:40 + 4 * t, 20 + 5 * t:60 + 5 * t] = 255 - base[10 + 4 * t:40 + 4 * t, 20 + 5 * t:60 + 5 * t]
        frames.append(f)
    return frames


@pytest.mark.parametrize("kw", [{}, {"learningFrames": 3}, {"learningFrames": -1, "alphaDetection": 0.3, "threshold": 10},
                                {"learningFrames": 5, "alphaLearn": 0.5, "threshold": 40},
                                {"learningFrames": 2, "threshold": 150}, {"learningFrames": 2, "alphaDetection": 0.0, "threshold": 0}])
def test_asbl_quad_kernel_geometries(oracle, kw):
    """asbl_fused16_kernel (frame width a multiple of 16): tiles cut by the right / bottom image border (400 px = 100
    words: the last tile holds 4; 70 and 33 rows), a single quad, a one-row image, an exact tile grid; thresholds on both
    sides of 127 (the word-wise and the byte-wise compare), alpha 0 (quiet radius 255: no lookups at all); host path,
    then the device path on the same model, then a three-stream group."""
    import torch
    import tracking_b200 as tb
    for (h, w) in ((70, 400), (33, 16), (1, 48), (64, 256), (40, 144)):
        frames = _asbl_frames(h, w, 10, 5)
        p, o = tb.AdaptiveSelectiveBackgroundLearning(**kw), oracle.AdaptiveSelectiveBackgroundLearning(**kw)
        for i, f in enumerate(frames[:6]):
            fa, ba = p.process(f)
            fb, bb = o.process(f)
            assert np.array_equal(fa, fb) and np.array_equal(ba, bb), (h, w, i)
        d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        d_bg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames[6:]):
            d_in = torch.from_numpy(f).cuda()
            p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
            fb, bb = o.process(f)
            assert np.array_equal(d_fg.cpu().numpy(), fb) and np.array_equal(d_bg.cpu().numpy(), bb), (h, w, i)
        p.close()
    h, w = 48, 160
    vids = [_asbl_frames(h, w, 6, 20 + s) for s in range(3)]
    g = tb.AdaptiveSelectiveBackgroundLearning(nstreams=3, **kw)
    os_ = [oracle.AdaptiveSelectiveBackgroundLearning(**kw) for _ in range(3)]
    for i in range(6):
        fg, bg = g.process(np.stack([v[i] for v in vids]))
        for s in range(3):
            fb, bb = os_[s].process(vids[s][i])
            assert np.array_equal(fg[s], fb) and np.array_equal(bg[s], bb), (s, i)
    g.close()


@pytest.mark.parametrize("kw", [{}, {"learningFrames": 3}, {"learningFrames": -1, "alphaDetection": 0.3, "threshold": 10},
                                {"learningFrames": 5, "alphaLearn": 0.5, "threshold": 40}])
def test_asbl_sibling_plugin(oracle, kw):
    """AdaptiveSelectiveBackgroundLearning (USTC_BGS type 7): gray model, 3x3 median with replicated border, learning
    phase then selective update; single-channel background image; host path, device path and reset."""
    import torch
    import tracking_b200 as tb
    # (600, 700) and the two small multiples of 4 take the single-pass tile kernel (partial and exact tiles)
    for (h, w) in ((97, 131), (600, 700), (64, 256), (35, 132)):
        frames = _asbl_frames(h, w, 12, 3)
        p, o = tb.AdaptiveSelectiveBackgroundLearning(**kw), oracle.AdaptiveSelectiveBackgroundLearning(**kw)
        for i, f in enumerate(frames[:8]):
            fa, ba = p.process(f)
            fb, bb = o.process(f)
            assert ba.shape == (h, w)
            assert np.array_equal(fa, fb) and np.array_equal(ba, bb), (h, w, i)
        d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        d_bg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames[8:]):                       # device buffers continue the same model
            d_in = torch.from_numpy(f).cuda()
            fv, bv = p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
            torch.cuda.synchronize()
            fb, bb = o.process(f)
            assert fv and bv
            assert np.array_equal(d_fg.cpu().numpy(), fb) and np.array_equal(d_bg.cpu().numpy(), bb), (h, w, i)
        p.reset()
        o2 = oracle.AdaptiveSelectiveBackgroundLearning(**kw)
        fa, ba = p.process(frames[0]); fb, bb = o2.process(frames[0])
        assert np.array_equal(fa, fb) and np.array_equal(ba, bb)
        p.close()
    u = tb.USTC_BGS(7)
    u.Process(frames[0]); u.Process(frames[1])
    assert u.GetMask().shape == (h, w)
    u.Release()


def test_asbl_stream_group_and_fanout(oracle):
    """Two streams advanced by one launch pair, and ASBL next to banded plugins in a fan-out (it runs once on the
    whole frame after the last band)."""
    import tracking_b200 as tb
    h, w = 600, 700
    fa_ = _asbl_frames(h, w, 5, 7); fb_ = _asbl_frames(h, w, 5, 8)
    g = tb.AdaptiveSelectiveBackgroundLearning(nstreams=2, learningFrames=2)
    oa, ob = oracle.AdaptiveSelectiveBackgroundLearning(learningFrames=2), oracle.AdaptiveSelectiveBackgroundLearning(learningFrames=2)
    for i in range(5):
        fg, bg = g.process(np.stack([fa_[i], fb_[i]]))
        ra, rb = oa.process(fa_[i]), ob.process(fb_[i])
        assert np.array_equal(fg[0], ra[0]) and np.array_equal(fg[1], rb[0]), i
        assert np.array_equal(bg[0], ra[1]) and np.array_equal(bg[1], rb[1]), i
    g.close()
    ps = [tb.FrameDifferenceBGS(), tb.AdaptiveSelectiveBackgroundLearning(), tb.MixtureOfGaussianV2BGS()]
    os_ = [oracle.FrameDifferenceBGS(), oracle.AdaptiveSelectiveBackgroundLearning(), oracle.MixtureOfGaussianV2BGS()]
    for i in range(5):
        outs = tb.process_fanout(ps, fa_[i])
        for k, o in enumerate(os_):
            fb, bb = o.process(fa_[i])
            assert (outs[k][0] is None) == (fb is None), (k, i)
            if fb is not None:
                assert np.array_equal(outs[k][0], fb), (k, i)
            if bb is not None:
                assert np.array_equal(outs[k][1], bb), (k, i)
    for p in ps:
        p.close()


@pytest.mark.parametrize("kw", [{}, {"alpha": 0.05, "threshold": 9.0, "gaussians": 5}, {"alpha": 0.3, "gaussians": 2},
                                {"alpha": 0.01, "threshold": 12.5, "gaussians": 4}])
def test_dpzivkovic_sibling_plugin(oracle, clips, kw):
    """DPZivkovicAGMMBGS (USTC_BGS type 11): bit-exact masks against the restatement that is pinned to a build of the
    reference's own sources; reference clip, mode-churn stress sequence, large frame through the row-band host path,
    device path, two-stream group; img_bgmodel is never produced."""
    import torch
    import tracking_b200 as tb
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, (600, 700, 3), dtype=np.uint8)
    bigs = []
    for t in range(10):
        f = np.clip(big.astype(np.int16) + rng.integers(-8, 9, big.shape), 0, 255).astype(np.uint8)
        f[50 + 10 * t:200 + 10 * t, 100 + 20 * t:300 + 20 * t] = rng.integers(0, 256, 3)
        bigs.append(f)
    for frames in (list(clips["video_clip"]), stress_sequence(120, 40, 52), bigs):
        p, o = tb.DPZivkovicAGMMBGS(**kw), oracle.DPZivkovicAGMMBGS(**kw)
        for i, f in enumerate(frames):
            fa, ba = p.process(f)
            fb, _ = o.process(f)
            assert ba is None
            assert np.array_equal(fa, fb), i
        p.close()
    # device path + stream group
    h, w = bigs[0].shape[:2]
    g = tb.DPZivkovicAGMMBGS(nstreams=2, **kw)
    oa, ob = oracle.DPZivkovicAGMMBGS(**kw), oracle.DPZivkovicAGMMBGS(**kw)
    d_fg = torch.zeros((2, h, w), dtype=torch.uint8, device="cuda")
    for i in range(6):
        d_in = torch.from_numpy(np.stack([bigs[i], bigs[9 - i]])).cuda()
        fv, bv = g.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None)
        torch.cuda.synchronize()
        assert fv and not bv
        out = d_fg.cpu().numpy()
        assert np.array_equal(out[0], oa.process(bigs[i])[0]) and np.array_equal(out[1], ob.process(bigs[9 - i])[0]), i
    g.close()


DP_SIMPLE_CASES = [("DPAdaptiveMedianBGS", {}), ("DPAdaptiveMedianBGS", {"threshold": 10, "samplingRate": 2}),
                   ("DPAdaptiveMedianBGS", {"threshold": 200, "samplingRate": 3}), ("DPAdaptiveMedianBGS", {"threshold": 25, "samplingRate": 1}),
                   ("DPMeanBGS", {}), ("DPMeanBGS", {"threshold": 300, "alpha": 0.9}), ("DPMeanBGS", {"threshold": 50, "alpha": 0.999}),
                   ("DPWrenGABGS", {}), ("DPWrenGABGS", {"threshold": 3.0, "alpha": 0.2}), ("DPWrenGABGS", {"threshold": 0.5, "alpha": 0.9}),
                   ("DPPratiMediodBGS", {}), ("DPPratiMediodBGS", {"threshold": 10, "samplingRate": 1, "historySize": 4}),
                   ("DPPratiMediodBGS", {"threshold": 20, "samplingRate": 3, "historySize": 7, "weight": 1}),
                   ("DPPratiMediodBGS", {"threshold": 5, "samplingRate": 20})]


@pytest.mark.parametrize("name,kw", DP_SIMPLE_CASES)
def test_dp_simple_sibling_plugins(oracle, clips, name, kw):
    """DPAdaptiveMedianBGS / DPMeanBGS / DPWrenGABGS / DPPratiMediodBGS (USTC_BGS types 9, 12, 13, 14): bit-exact masks
    against the restatements that are pinned to a build of the reference's own sources (tests/test_oracle_pin.py,
    golden_dp.json).  The reference clip and the stress sequence (ragged frames: 40 x 52) through the host path, then the
    device path on the same model, a two-stream group, a temporal batch, parameters latched on the first frame,
    and reset."""
    import torch
    import tracking_b200 as tb
    from conftest import stress_sequence
    cls, ocls = getattr(tb, name), getattr(oracle, name)
    for frames in (list(clips["video_clip"][:40]), stress_sequence(60, 40, 52)):
        h, w = frames[0].shape[:2]
        p, o = cls(**kw), ocls(**kw)
        half = len(frames) // 2
        for i, f in enumerate(frames[:half]):
            fa, ba = p.process(f)
            fb, bb = o.process(f)
            assert ba is None and bb is None
            assert np.array_equal(fa, fb), (name, kw, i)
        # a parameter changed after the first frame has no effect until reset (the wrappers hand them over once)
        p.set("threshold", 1)
        d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames[half:]):
            d_in = torch.from_numpy(np.ascontiguousarray(f)).cuda()
            fv, bv = p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None)
            fb, _ = o.process(f)
            assert fv and not bv
            assert np.array_equal(d_fg.cpu().numpy(), fb), (name, kw, half + i)
        p.reset()
        for k, v in kw.items():
            p.set(k, v)
        if "threshold" not in kw:
            p.set("threshold", ocls().threshold)
        o2 = ocls(**kw)
        for i, f in enumerate(frames[:4]):
            assert np.array_equal(p.process(f)[0], o2.process(f)[0]), (name, kw, "after reset", i)
        p.close()
    # two streams advanced by one launch; then a temporal batch of four frames per stream
    fa_, fb_ = list(clips["video_clip"][:12]), list(clips["video_clip"][20:32])
    h, w = fa_[0].shape[:2]
    g = cls(nstreams=2, **kw)
    oa, ob = ocls(**kw), ocls(**kw)
    for i in range(8):
        fg, bg = g.process(np.stack([fa_[i], fb_[i]]))
        assert np.array_equal(fg[0], oa.process(fa_[i])[0]) and np.array_equal(fg[1], ob.process(fb_[i])[0]), (name, kw, i)
    batch = np.stack([np.stack(fa_[8:12]), np.stack(fb_[8:12])])                 # [S][T][h][w][3]
    d_b = torch.from_numpy(batch).cuda()
    d_fg = torch.zeros((2, 4, h, w), dtype=torch.uint8, device="cuda")
    g.process_batch_dev(d_b.data_ptr(), 4, w, h, d_fg.data_ptr(), None)
    out = d_fg.cpu().numpy()
    for t in range(4):
        assert np.array_equal(out[0, t], oa.process(fa_[8 + t])[0]) and np.array_equal(out[1, t], ob.process(fb_[8 + t])[0]), (name, kw, t)
    g.close()


@pytest.mark.parametrize("kw", [{}, {"ampFactor": 3, "minVar": 2, "maxVar": 40}, {"ampFactor": 2, "minVar": 300, "maxVar": 700},
                                {"ampFactor": 5, "minVar": 1}])
def test_sigma_delta_sibling_plugin(oracle, clips, kw):
    """SigmaDeltaBGS (USTC_BGS type 35): bit-exact against the restatement pinned to a build of the reference's
    sdLaMa091.cpp -- including the signed-char wrap of large differences (an inverted frame mid-sequence), the uint8_t wrap
    of the variance with N > 1, parameters truncated to a byte, and the initialiser's first-third-of-every-row rule.  First
    frame: no outputs.  Host path, device path, live parameter change, a two-stream group, a temporal batch that starts on
    the warm-up frame, a 1080p frame through the row-banded host path."""
    import torch
    import tracking_b200 as tb
    from conftest import stress_sequence
    for frames in (list(clips["video_clip"][:24]), stress_sequence(40, 40, 52)):
        frames = [f.copy() for f in frames]
        frames[7] = 255 - frames[7]
        h, w = frames[0].shape[:2]
        p, o = tb.SigmaDeltaBGS(**kw), oracle.SigmaDeltaBGS(**kw)
        d_fg = torch.full((h, w), 9, dtype=torch.uint8, device="cuda")
        for i, f in enumerate(frames):
            if i == 12:
                p.set("minVar", 7); o.minVar = 7                  # parameters are applied before every frame
            fb, bb = o.process(f)
            if i % 2 == 0:
                fa, ba = p.process(f)
                assert ba is None and (fa is None) == (fb is None), i
                if fb is not None:
                    assert np.array_equal(fa, fb), (kw, i)
            else:
                d_in = torch.from_numpy(np.ascontiguousarray(f)).cuda()
                fv, bv = p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None)
                assert fv == (fb is not None) and not bv
                if fb is not None:
                    assert np.array_equal(d_fg.cpu().numpy(), fb), (kw, i)
        p.close()
    fa_, fb_ = list(clips["video_clip"][:8]), list(clips["video_clip"][16:24])
    h, w = fa_[0].shape[:2]
    g = tb.SigmaDeltaBGS(nstreams=2, **kw)
    oa, ob = oracle.SigmaDeltaBGS(**kw), oracle.SigmaDeltaBGS(**kw)
    batch = np.stack([np.stack(fa_[:4]), np.stack(fb_[:4])])                     # [S][T][h][w][3], frame 0 = warm-up
    d_b = torch.from_numpy(batch).cuda()
    d_fg = torch.full((2, 4, h, w), 9, dtype=torch.uint8, device="cuda")
    first, _ = g.process_batch_dev(d_b.data_ptr(), 4, w, h, d_fg.data_ptr(), None)
    assert first == 1
    out = d_fg.cpu().numpy()
    for t in range(4):
        ra, rb = oa.process(fa_[t])[0], ob.process(fb_[t])[0]
        if t == 0:
            assert ra is None and (out[:, 0] == 9).all()                         # untouched
        else:
            assert np.array_equal(out[0, t], ra) and np.array_equal(out[1, t], rb), (kw, t)
    for i in range(4, 8):
        fg, _ = g.process(np.stack([fa_[i], fb_[i]]))
        assert np.array_equal(fg[0], oa.process(fa_[i])[0]) and np.array_equal(fg[1], ob.process(fb_[i])[0]), (kw, i)
    g.close()
    from tracking_b200 import synth
    big = [synth.frame(1920, 1080, t) for t in range(4)]
    p, o = tb.SigmaDeltaBGS(**kw), oracle.SigmaDeltaBGS(**kw)
    for i, f in enumerate(big):
        fa, fb = p.process(f)[0], o.process(f)[0]
        assert (fa is None) == (fb is None) and (fb is None or np.array_equal(fa, fb)), (kw, "1080p", i)
    p.close()


def test_dp_simple_plugins_at_1080p(oracle):
    """The three DP models on a 1080p K-GEN stream through the banded host path (row-band sub-launches) and on a
    two-stream device group: 12 frames against the restatements."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import synth
    w, h, n = 1920, 1080, 12
    d = torch.empty((2, n, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), 2, n, w, h)
    torch.cuda.synchronize()
    host = d.cpu().numpy()
    for name in ("DPAdaptiveMedianBGS", "DPMeanBGS", "DPWrenGABGS", "DPPratiMediodBGS"):
        kw = {"DPAdaptiveMedianBGS": {"threshold": 8, "samplingRate": 2}, "DPMeanBGS": {"threshold": 60, "alpha": 0.9},
              "DPWrenGABGS": {"threshold": 2.0, "alpha": 0.05}, "DPPratiMediodBGS": {"threshold": 4, "samplingRate": 2, "historySize": 3}}[name]
        p, g = getattr(tb, name)(**kw), getattr(tb, name)(nstreams=2, **kw)
        os_ = [getattr(oracle, name)(**kw) for _ in range(2)]
        d_in = torch.empty((2, h, w, 3), dtype=torch.uint8, device="cuda")
        d_fg = torch.zeros((2, h, w), dtype=torch.uint8, device="cuda")
        total = 0
        for t in range(n):
            ref = [os_[s].process(host[s, t])[0] for s in range(2)]
            fa, _ = p.process(host[0, t])
            assert np.array_equal(fa, ref[0]), (name, t)
            d_in.copy_(d[:, t])
            g.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None)
            out = d_fg.cpu().numpy()
            assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]), (name, t)
            total += int((ref[0] != 0).sum())
        assert 0 < total < n * w * h
        p.close(); g.close()


def test_second_device_in_one_process(oracle, clips):
    """Contexts on two devices of one process (per-device function attributes, stream and buffer ownership).
    Skipped on single-GPU boxes."""
    import torch
    import tracking_b200 as tb
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    frames = _asbl_frames(600, 700, 4, 11)
    for aid in (6, 5, 3, 0, 7, 11, 9, 12, 13, 14, 35):
        p0, p1 = tb.ALGOS[aid](device=0), tb.ALGOS[aid](device=1)
        o = oracle.ALGOS[aid]()
        for f in frames:
            a, b = p0.process(f), p1.process(f)
            r = o.process(f)
            for got in (a, b):
                assert (got[0] is None) == (r[0] is None)
                if r[0] is not None:
                    assert np.array_equal(got[0], r[0]), aid
                if r[1] is not None:
                    assert np.array_equal(got[1], r[1]), aid
        p0.close(); p1.close()


def test_mog2_state_export_import_roundtrip(clips):
    import tracking_b200 as tb
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    a = tb.MixtureOfGaussianV2BGS()
    for f in clip[:10]:
        a.process(f)
    planes, nm = a.export_state()
    b = tb.MixtureOfGaussianV2BGS()
    b.import_state(planes, nm, w, h, a.frame_count)
    for f in clip[10:16]:
        fa, ba = a.process(f)
        fb, bb = b.process(f)
        assert np.array_equal(fa, fb) and np.array_equal(ba, bb)


def test_ustc_bgs_adapter_contract(clips):
    import tracking_b200 as tb
    with pytest.raises(ValueError):
        tb.USTC_BGS(36)
    d = tb.USTC_BGS(0)
    assert d.GetMask() is None                     # frameNum == 0 -> NULL (ustc_bgs.cpp:81)
    d.Process(clips["png_clip"][0])
    assert d.GetMask() is None                     # FD warm-up frame: still no mask
    d.Process(clips["png_clip"][1])
    assert d.GetMask() is not None and d.GetMask().shape == clips["png_clip"][0].shape[:2]
    d.Release()


def test_synth_generator_matches_numpy_twin():
    import torch
    from tracking_b200 import synth
    S, T, w, h = 2, 3, 333, 250
    d = torch.zeros((S, T, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), S, T, w, h, t0=7)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    for s in range(S):
        for t in range(T):
            assert np.array_equal(got[s, t], synth.frame(w, h, 7 + t, synth.SEED0 + s))


def test_full_size_1080p_mog2_parity_and_launch_counter(oracle):
    """BASELINE config 2 geometry against the C oracle on a few frames + size-independent properties."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import synth
    w, h, n = 1920, 1080, 5
    d = torch.zeros((1, n, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), 1, n, w, h, t0=0)
    torch.cuda.synchronize()
    frames = d.cpu().numpy()[0]
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    before = tb.kernel_launch_count()
    for t in range(n):
        p.process_dev(d[0, t].data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
        torch.cuda.synchronize()
        ofg, obg = o.process(frames[t])
        assert np.array_equal(d_fg.cpu().numpy(), ofg), t
        assert np.array_equal(d_bg.cpu().numpy(), obg), t
    assert tb.kernel_launch_count() - before == n
    # property: a static scene converges to "no foreground" and bg == scene
    q = tb.MixtureOfGaussianV2BGS()
    for _ in range(30):
        q.process_dev(d[0, 0].data_ptr(), w, h, d_fg.data_ptr(), d_bg.data_ptr())
    torch.cuda.synchronize()
    assert not d_fg.any().item()
    assert torch.equal(d_bg, d[0, 0])


def test_max_size_2160p_parity(oracle):
    """BASELINE config 5 geometry (3840x2160): MOG2 (T = 1 and a temporal batch) and FD against the C oracle."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import synth
    w, h, n = 3840, 2160, 4
    d = torch.zeros((1, n, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.frames_dev(d.data_ptr(), 1, n, w, h, t0=3)
    torch.cuda.synchronize()
    frames = d.cpu().numpy()[0]
    assert np.array_equal(frames[1], synth.frame(w, h, 4))          # the 2x-scaled rectangles of the 2160p generator
    d_fg = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    exp = []
    o = oracle.MixtureOfGaussianV2BGS()
    for t in range(n):
        exp.append(o.process(frames[t]))
    # frame by frame
    p = tb.MixtureOfGaussianV2BGS()
    for t in range(n):
        p.process_dev(d[0, t].data_ptr(), w, h, d_fg[t].data_ptr(), d_bg[t].data_ptr())
    torch.cuda.synchronize()
    for t in range(n):
        assert np.array_equal(d_fg[t].cpu().numpy(), exp[t][0]) and np.array_equal(d_bg[t].cpu().numpy(), exp[t][1]), t
    # one temporal batch of all 4 frames
    q = tb.MixtureOfGaussianV2BGS()
    d_fg.zero_(); d_bg.zero_()
    q.process_batch_dev(d.data_ptr(), n, w, h, d_fg.data_ptr(), d_bg.data_ptr())
    torch.cuda.synchronize()
    for t in range(n):
        assert np.array_equal(d_fg[t].cpu().numpy(), exp[t][0]) and np.array_equal(d_bg[t].cpu().numpy(), exp[t][1]), t
    # FD
    fd, ofd = tb.FrameDifferenceBGS(), oracle.FrameDifferenceBGS()
    first, _ = fd.process_batch_dev(d.data_ptr(), n, w, h, d_fg.data_ptr(), None)
    torch.cuda.synchronize()
    assert first == 1
    for t in range(n):
        e, _ = ofd.process(frames[t])
        if e is not None:
            assert np.array_equal(d_fg[t].cpu().numpy(), e), t


def test_trace_stage_times_and_toc_line(tmp_path):
    """Tracing (FrameProcessor::tic / toc equivalent, per stage): the "trace" parameter fills bgsb_trace_last;
    BGSB_TRACE=1 prints toc's line with the stage split on stderr."""
    import os
    import subprocess
    import sys
    import tracking_b200 as tb
    from tracking_b200 import synth
    f = [synth.frame(1920, 1080, t) for t in range(3)]
    p = tb.MixtureOfGaussianV2BGS()
    with pytest.raises(tb.BgsbError):
        p.trace_last()                                   # nothing traced yet
    p.set("trace", 1)
    for x in f:
        p.process(x)
    t = p.trace_last()
    assert t["frame"] == 2 and t["bands"] == 3           # mask + background image go back: 3 row bands by default
    assert t["upload_ms"] > 0.02 and t["download_ms"] > 0.02 and t["kernel_ms"] > 0.005
    assert t["wall_ms"] >= max(t["upload_ms"], t["download_ms"]) * 0.9
    p.close()
    big_fd = tb.FrameDifferenceBGS(trace=1)              # no background image: 2 bands; an explicit "hostBands" wins
    for x in f:
        big_fd.process(x)
    assert big_fd.trace_last()["bands"] == 2
    big_fd.set("hostBands", 4)
    big_fd.process(f[0])
    assert big_fd.trace_last()["bands"] == 4
    big_fd.close()
    small = tb.FrameDifferenceBGS(trace=1)
    small.process(np.zeros((40, 50, 3), np.uint8))       # warm-up frame: uploaded, no kernel output
    small.process(np.zeros((40, 50, 3), np.uint8))
    assert small.trace_last()["bands"] == 1
    small.close()
    code = ("import numpy as np, tracking_b200 as tb\n"
            "p = tb.AdaptiveBackgroundLearning()\n"
            "for _ in range(2): p.process(np.zeros((64, 64, 3), np.uint8))\n")
    env = dict(os.environ, BGSB_TRACE="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stderr.splitlines() if l.startswith("AdaptiveBackgroundLearning\ttime(sec):")]
    assert len(lines) == 2 and "upload" in lines[0] and "download" in lines[0]


def test_churn_generator_matches_numpy_twin_and_keeps_modes_live(oracle):
    import torch
    import tracking_b200 as tb
    from tracking_b200 import synth
    S, T, h, w = 2, 3, 97, 131
    d = torch.zeros((S, T, h, w, 3), dtype=torch.uint8, device="cuda")
    synth.churn_frames_dev(d.data_ptr(), S, T, w, h, t0=5, seed0=synth.SEED0)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    for s in range(S):
        for t in range(T):
            assert np.array_equal(got[s, t], synth.churn_frame(w, h, 5 + t, synth.SEED0 + s))
    # MOG2 on the churn stream: the dense case (most pixels at 5 modes), masks / background / state against the oracle
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    for t in range(40):
        f = synth.churn_frame(w, h, t)
        fg, bg = p.process(f)
        ofg, obg = o.process(f)
        assert np.array_equal(fg, ofg) and np.array_equal(bg, obg), t
    _, nm = p.export_state()
    assert np.array_equal(nm, o.nmodes) and (nm == 5).mean() > 0.6
    p.close()


def test_copy_probe_reports_plausible_pcie_rates():
    import ctypes as C
    from tracking_b200 import capi
    up, dn, both = C.c_double(0), C.c_double(0), C.c_double(0)
    capi.check(capi.lib().bgsb_copy_probe(0, 6 << 20, 8 << 20, 20, C.byref(up), C.byref(dn), C.byref(both)))
    for t in (up.value, dn.value, both.value):
        assert 1e-5 < t < 1e-2                       # 6-8 MB at somewhere between 1 and 800 GB/s
    assert both.value >= max(up.value, dn.value) * 0.8
    assert capi.lib().bgsb_copy_probe(0, 0, 1, 1, C.byref(up), C.byref(dn), C.byref(both)) == capi.ERR_ARG


def test_gray_variant_24_golden_hashes(clips):
    """`grayVariant` 1 (OpenCV 2.4 BGR2GRAY constants, what the C++ adapters select in a 2.4 build) against hashes of a
    pure-numpy restatement committed in tests/golden/golden_gray24.json (tests/golden/make_golden_gray24.py)."""
    import json
    import os
    import tracking_b200 as tb
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(gdir, "golden_gray24.json")))
    sys.path.insert(0, gdir)
    import make_golden_gray24 as mk
    ties, n = mk.tie_frames()                    # every colour on which the two generations of constants disagree at threshold 15
    assert n == gold["gray_ties"]["colours"]
    for variant in (0, 1):
        p = tb.FrameDifferenceBGS(grayVariant=variant)
        fgs, _ = run_host(p, list(ties))
        assert sha(fgs) == gold["gray_ties"]["FrameDifferenceBGS:grayVariant=%d" % variant]
        p.close()
    for name, clip in clips.items():
        for aid, key in ((0, "FrameDifferenceBGS:grayVariant=1"), (1, "StaticFrameDifferenceBGS:grayVariant=1")):
            p = tb.ALGOS[aid](grayVariant=1)
            fgs, _ = run_host(p, list(clip))
            assert sha(fgs) == gold[name][key], (name, key)
            p.close()
