"""The device-resident foreground pipeline (bgsb_pipeline_*: plugin -> erode/dilate chain -> labelling with the mask
bit-packed in between) against the oracle chain: plugin restatement -> orc_morph3x3 -> orc_ccl8 / rect moments.
Bit-exact: cleaned masks, canonical label images, component tables, external flags, moments."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def oracle_chain(oracle, fg, chain, zero_border):
    clean = fg
    for op, it in chain:
        clean = oracle.morph(clean, op, it)
    # DetectNewBlob thresholds a clone at 128 before cvFindContours
    return (clean,) + tuple(oracle.ccl8(clean, zero_border))


def check_stream(pipe, s, clean, lab, exp, oracle):
    eclean, en, elab, est, eext = exp
    if clean is not None:
        assert np.array_equal(clean, eclean), "cleaned mask"
    if lab is not None:
        assert np.array_equal(lab, elab), "label image"
    comps = pipe.components(s)
    assert len(comps) == en
    for c, st_, e in zip(comps, est, eext):
        assert (c["x"], c["y"], c["x"] + c["w"] - 1, c["y"] + c["h"] - 1, c["area"], c["first_index"]) == \
            tuple(int(v) for v in st_)
        assert c["external"] == int(e)
    ih, iw = eclean.shape
    x0, y0 = min(3, iw - 1), min(2, ih - 1)            # cvGetSubRect needs rectangles inside the image
    rects = [(c["x"], c["y"], c["w"], c["h"]) for c in comps[:6]] + [(0, 0, iw, ih), (x0, y0, min(45, iw - x0), min(30, ih - y0))]
    got = pipe.rect_moments(rects, s)
    for r, g in zip(rects, got):
        assert g == oracle.rect_moments(eclean, r), r


@pytest.mark.parametrize("aid", [5, 0, 3, 6])
@pytest.mark.parametrize("chain", [(("erode", 1), ("dilate", 1)), (), (("dilate", 2), ("erode", 2)), (("erode", 9), ("dilate", 9))])
def test_pipeline_matches_oracle_chain_on_reference_clip(oracle, clips, aid, chain):
    """3 streams of the reference clip (320 x 176: 10 words per row), every output requested."""
    import torch
    from tracking_b200.pipeline import ForegroundPipeline
    clip = clips["video_clip"]
    S, n = 3, 14
    streams = [clip[:n], clip[::-1][:n], clip[6:6 + n]]
    h, w = clip.shape[1:3]
    pipe = ForegroundPipeline(aid, nstreams=S, morph=chain)
    os_ = [oracle.ALGOS[aid]() for _ in range(S)]
    d_mask = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_bg = torch.zeros((S, h, w, 3), dtype=torch.uint8, device="cuda")
    d_lab = torch.zeros((S, h, w), dtype=torch.int32, device="cuda")
    for t in range(n):
        d_in = torch.from_numpy(np.stack([streams[s][t] for s in range(S)])).cuda()
        valid, bgv = pipe.process_dev(d_in.data_ptr(), w, h, d_mask.data_ptr(), d_bg.data_ptr(), d_lab.data_ptr())
        torch.cuda.synchronize()
        mask, bg, lab = d_mask.cpu().numpy(), d_bg.cpu().numpy(), d_lab.cpu().numpy()
        for s in range(S):
            ofg, obg = os_[s].process(streams[s][t])
            assert valid == (ofg is not None)
            if obg is not None:
                assert bgv and np.array_equal(bg[s], obg)
            if ofg is None:
                continue
            check_stream(pipe, s, mask[s], lab[s], oracle_chain(oracle, ofg, chain, True), oracle)
    pipe.close()


@pytest.mark.parametrize("shape", [(97, 131), (40, 33), (64, 1), (3, 70)])
@pytest.mark.parametrize("zb", [1, 0])
def test_pipeline_ragged_widths_packed_mask(oracle, shape, zb):
    """Widths that are not multiples of 32 (the packed rows end in a partial word), MOG2, table only and full outputs."""
    import torch
    from tracking_b200 import synth
    from tracking_b200.pipeline import ForegroundPipeline
    h, w = shape
    n = 8
    rng = np.random.default_rng(5)
    frames = [np.clip(synth.frame(w, h, t).astype(np.int16) + (rng.random((h, w, 1)) < 0.1) * 90, 0, 255).astype(np.uint8)
              for t in range(n)]
    chain = (("erode", 1), ("dilate", 1))
    for full in (True, False):
        pipe = ForegroundPipeline(5, nstreams=1, morph=chain, zeroBorder=zb)
        o = oracle.MixtureOfGaussianV2BGS()
        d_mask = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
        d_lab = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        for t in range(n):
            d_in = torch.from_numpy(frames[t]).cuda()
            pipe.process_dev(d_in.data_ptr(), w, h, d_mask.data_ptr() if full else None, None,
                             d_lab.data_ptr() if full else None)
            torch.cuda.synchronize()
            ofg, _ = o.process(frames[t])
            exp = oracle_chain(oracle, ofg, chain, bool(zb))
            check_stream(pipe, 0, d_mask.cpu().numpy() if full else None, d_lab.cpu().numpy() if full else None, exp, oracle)
        pipe.close()


@pytest.mark.parametrize("chain_ctas", [0, 200])
def test_pipeline_background_pass_forms(oracle, chain_ctas):
    """RETR_EXTERNAL's background pass inside the pipeline, forced on every frame: as one cooperative launch (default)
    and as four plain launches ("chainCtas"); external flags and tables against the oracle on 3 streams."""
    import torch
    from tracking_b200.pipeline import ForegroundPipeline
    S, n = 3, 8
    h, w = 120, 200
    yy, xx = np.mgrid[0:h, 0:w]

    def scene(s, t):
        """rings with islands, blinking: FrameDifference sees ring + island, the island sits in the ring's hole"""
        img = np.zeros((h, w, 3), np.uint8)
        if t % 2 == 0:
            return img                                   # every other frame is empty: the difference is the whole drawing
        for cy, cx, r in ((40 + t % 5, 50 + 2 * s, 30), (70, 140 - t % 7, 38 + s)):
            d = (yy - cy) ** 2 + (xx - cx) ** 2
            img[(d < r * r) & (d >= (r - 6) ** 2)] = 220
            img[d < 25] = 180
        return img

    chain = (("dilate", 1),)
    pipe = ForegroundPipeline(0, nstreams=S, morph=chain, forceBackgroundPass=1, chainCtas=chain_ctas)
    os_ = [oracle.FrameDifferenceBGS() for _ in range(S)]
    seen_nested = False
    for t in range(n):
        frames = np.stack([scene(s, t) for s in range(S)])
        d_in = torch.from_numpy(frames).cuda()
        valid, _ = pipe.process_dev(d_in.data_ptr(), w, h)
        for s in range(S):
            ofg, _ = os_[s].process(frames[s])
            assert valid == (ofg is not None)
            if ofg is None:
                continue
            exp = oracle_chain(oracle, ofg, chain, True)
            seen_nested |= bool((exp[4] == 0).any())
            check_stream(pipe, s, None, None, exp, oracle)
    assert seen_nested            # some component really sits in a hole of another one on these masks
    pipe.close()


def test_pipeline_raw_mask_without_threshold_and_chain(oracle, clips):
    """enableThreshold = 0 (raw {0,127,255}) and no chain: DetectNewBlob's own threshold (> 128) decides what is
    labelled, and the moments weigh the raw mask values (shadow pixels = 127)."""
    import torch
    from tracking_b200.pipeline import ForegroundPipeline
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    pipe = ForegroundPipeline(5, nstreams=1, morph=(), enableThreshold=0)
    o = oracle.MixtureOfGaussianV2BGS(enableThreshold=False)
    d_mask = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    d_lab = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    seen127 = False
    for t in range(12):
        d_in = torch.from_numpy(clip[t]).cuda()
        pipe.process_dev(d_in.data_ptr(), w, h, d_mask.data_ptr(), None, d_lab.data_ptr())
        torch.cuda.synchronize()
        ofg, _ = o.process(clip[t])
        seen127 |= bool((ofg == 127).any())
        assert np.array_equal(d_mask.cpu().numpy(), ofg)
        n, elab, est, eext = oracle.ccl8(ofg, True)
        assert np.array_equal(d_lab.cpu().numpy(), elab)
        comps = pipe.components(0)
        assert len(comps) == n
        rects = [(c["x"], c["y"], c["w"], c["h"]) for c in comps[:5]] + [(0, 0, w, h)]
        for r, g in zip(rects, pipe.rect_moments(rects, 0)):
            assert g == oracle.rect_moments(ofg, r)
    assert seen127
    pipe.close()


def test_pipeline_config4_shape_8x1080p(oracle):
    """BASELINE config 4 at its own shape through the pipeline object: 8 x 1080p streams, MOG2 -> OPEN -> labelling,
    12 frames, all outputs, against eight oracle chains; then table-only frames (the benchmarked form)."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import synth
    from tracking_b200.pipeline import ForegroundPipeline
    S, w, h, n = 8, 1920, 1080, 12
    d = torch.empty((n, S, h, w, 3), dtype=torch.uint8, device="cuda")
    for t in range(n):
        synth.frames_dev(d[t].data_ptr(), S, 1, w, h, t0=t)
    torch.cuda.synchronize()
    host = d.cpu().numpy()
    chain = (("erode", 1), ("dilate", 1))
    pipe = ForegroundPipeline(5, nstreams=S, morph=chain)
    os_ = [oracle.MixtureOfGaussianV2BGS() for _ in range(S)]
    d_mask = torch.zeros((S, h, w), dtype=torch.uint8, device="cuda")
    d_lab = torch.zeros((S, h, w), dtype=torch.int32, device="cuda")

    def expect(s, t):
        ofg, _ = os_[s].process(host[t, s])
        return oracle_chain(oracle, ofg, chain, True)

    with ThreadPoolExecutor(S) as ex:
        for t in range(n):
            full = t < 8
            before = tb.kernel_launch_count()
            pipe.process_dev(d[t].data_ptr(), w, h, d_mask.data_ptr() if full else None, None,
                             d_lab.data_ptr() if full else None)
            assert tb.kernel_launch_count() - before == 6         # plugin, morphology, merge, roots, label, background
            exp = list(ex.map(lambda s: expect(s, t), range(S)))
            torch.cuda.synchronize()
            mask, lab = d_mask.cpu().numpy(), d_lab.cpu().numpy()
            for s in range(S):
                check_stream(pipe, s, mask[s] if full else None, lab[s] if full else None, exp[s], oracle)
    planes, nm = pipe.export_mog2_state(3)
    assert np.array_equal(nm, os_[3].nmodes)
    pipe.close()


def test_ccl_noisy_1080p_batch_with_nesting(oracle):
    """The labeller alone on a batch of 1080p masks that stress every path: thousands of components (salt noise), rings
    with islands (background pass taken), an empty and a full image; with and without the label image."""
    import torch
    from tracking_b200 import blobs
    rng = np.random.default_rng(9)
    h, w = 1080, 1920
    yy, xx = np.mgrid[0:h, 0:w]
    masks = []
    m = (rng.random((h, w)) < 0.002).astype(np.uint8) * 255
    masks.append(m)
    m = np.zeros((h, w), np.uint8)
    for cy, cx, r in ((300, 400, 200), (700, 1400, 300), (540, 960, 60)):
        dd = (yy - cy) ** 2 + (xx - cx) ** 2
        m[(dd < r * r) & (dd >= (r - 9) ** 2)] = 255
        m[dd < 100] = 255
    masks.append(m)
    masks.append(np.zeros((h, w), np.uint8))
    masks.append(np.full((h, w), 255, np.uint8))
    m = (rng.random((h, w)) < 0.45).astype(np.uint8) * 255          # long merge chains
    masks.append(m)
    m = np.zeros((h, w), np.uint8); m[::2, ::2] = 255               # the maximum component count
    masks.append(m)
    S = len(masks)
    d = torch.from_numpy(np.stack(masks)).cuda()
    d_lab = torch.zeros((S, h, w), dtype=torch.int32, device="cuda")
    cc = blobs.ConnectedComponents(w, h, max_images=S)
    for zb in (True, False):
        exp = [oracle.ccl8(mk, zb) for mk in masks]
        for want_labels in (True, False):
            d_lab.zero_()
            cc.label_batch_dev(d.data_ptr(), w, h, S, zb, d_lab.data_ptr() if want_labels else None)
            torch.cuda.synchronize()
            lab = d_lab.cpu().numpy()
            for i in range(S):
                n, elab, est, eext = exp[i]
                comps = cc.components(i)
                assert len(comps) == n, (i, zb)
                if want_labels:
                    assert np.array_equal(lab[i], elab), (i, zb)
                got = np.array([(c["x"], c["y"], c["x"] + c["w"] - 1, c["y"] + c["h"] - 1, c["area"], c["first_index"]) for c in comps],
                               np.int64).reshape(-1, 6)
                assert np.array_equal(got, est.astype(np.int64)), (i, zb)
                assert np.array_equal(np.array([c["external"] for c in comps], np.uint8), eext), (i, zb)
    # repeated calls on one labeller: the look-back state is clean between calls
    for _ in range(3):
        cc.label_batch_dev(d.data_ptr(), w, h, S, True, None)
    for i in (0, 1, 4):
        assert len(cc.components(i)) == oracle.ccl8(masks[i], True)[0]
    cc.close()
