"""Parity of K-MORPH, K-CCL and the CvBlobDetectorCC replacement against the oracle (through the C ABI)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def random_masks(rng, n, big=False):
    for _ in range(n):
        if big:
            h, w = int(rng.integers(200, 700)), int(rng.integers(1900, 2300))
        else:
            h, w = int(rng.integers(1, 90)), int(rng.integers(1, 200))
        p = rng.choice([0.02, 0.2, 0.45, 0.6, 0.9])
        m = (rng.random((h, w)) < p).astype(np.uint8) * 255
        if rng.random() < 0.5 and h > 8 and w > 8:      # blobs with holes and islands
            yy, xx = np.mgrid[0:h, 0:w]
            for _ in range(6):
                cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(2, max(3, min(h, w) // 2))
                d = (yy - cy) ** 2 + (xx - cx) ** 2
                m[(d < r * r) & (d >= (r * 0.6) ** 2)] = 255
                m[d < (r * 0.25) ** 2] = 255
        yield m


def test_morph_bit_exact(oracle):
    from tracking_b200 import blobs
    rng = np.random.default_rng(0)
    chains = [[("erode", 1)], [("dilate", 1)], [("erode", 1), ("dilate", 1)], [("dilate", 2), ("erode", 2)],
              [("erode", 0)], [("erode", 3), ("dilate", 5)], [("dilate", 11)], [("erode", 9), ("dilate", 9), ("erode", 1)]]
    for m in list(random_masks(rng, 24)) + list(random_masks(rng, 2, big=True)):
        for ch in chains:
            exp = m
            for op, it in ch:
                exp = oracle.morph(exp, op, it)
            got = blobs.morph(m, ch)
            assert np.array_equal(got, exp), (m.shape, ch)


def test_morph_device_batch_and_inplace(oracle):
    import torch
    from tracking_b200 import blobs
    rng = np.random.default_rng(1)
    ms = np.stack([(rng.random((123, 517)) < 0.5).astype(np.uint8) * 255 for _ in range(5)])
    d = torch.from_numpy(ms).cuda()
    out = torch.zeros_like(d)
    ch = [("erode", 1), ("dilate", 1)]
    blobs.morph_dev(d.data_ptr(), 517, 123, 5, ch, out.data_ptr())
    blobs.morph_dev(d.data_ptr(), 517, 123, 5, ch, d.data_ptr())          # in place
    torch.cuda.synchronize()
    for i in range(5):
        exp = oracle.morph(oracle.morph(ms[i], "erode", 1), "dilate", 1)
        assert np.array_equal(out[i].cpu().numpy(), exp)
        assert np.array_equal(d[i].cpu().numpy(), exp)


def test_ccl_labels_stats_external_bit_exact(oracle):
    from tracking_b200 import blobs
    rng = np.random.default_rng(2)
    cc = blobs.ConnectedComponents(2304, 704)
    masks = list(random_masks(rng, 40)) + list(random_masks(rng, 3, big=True))
    masks += [np.zeros((5, 7), np.uint8), np.full((9, 33), 255, np.uint8), np.full((1, 1), 255, np.uint8)]
    chk = np.zeros((40, 70), np.uint8); chk[::2, ::2] = 255          # maximum component count
    masks.append(chk)
    # values around the 128 threshold: foreground is strictly > 128
    masks.append(rng.integers(120, 136, (50, 60)).astype(np.uint8))
    for m in masks:
        for zb in (False, True):
            n, lab, comps = cc.label(m, zero_border=zb)
            on, olab, ost, oext = oracle.ccl8(m, zb)
            assert n == on, (m.shape, zb)
            assert np.array_equal(lab, olab), (m.shape, zb)           # canonical labels, bit-exact
            for c, s, e in zip(comps, ost, oext):
                assert (c["x"], c["y"], c["x"] + c["w"] - 1, c["y"] + c["h"] - 1, c["area"], c["first_index"]) == tuple(int(v) for v in s)
                assert c["external"] == int(e)
                assert c["label"] == comps.index(c) + 1
    cc.close()


def test_ccl_nesting_paths(oracle):
    """RETR_EXTERNAL nesting: rings with islands (background pass needed) and plain blobs (skipped), both with the
    pass forced and with the bounding-box pre-check deciding."""
    from tracking_b200 import blobs
    h, w = 120, 200
    yy, xx = np.mgrid[0:h, 0:w]
    nested = np.zeros((h, w), np.uint8)
    for cy, cx, r in ((40, 50, 30), (70, 140, 40)):
        d = (yy - cy) ** 2 + (xx - cx) ** 2
        nested[(d < r * r) & (d >= (r - 4) ** 2)] = 255       # ring
        nested[d < 16] = 255                                   # island inside the ring
        nested[(np.abs(yy - cy) < 8) & (np.abs(xx - cx - 12) < 2)] = 255
    plain = np.zeros((h, w), np.uint8)
    plain[10:30, 10:50] = 255; plain[60:100, 90:130] = 255
    # bbox strictly inside another bbox but NOT in a hole (an L-shaped blob around a small one)
    tricky = np.zeros((h, w), np.uint8)
    tricky[10:100, 10:20] = 255; tricky[90:100, 10:150] = 255; tricky[30:40, 60:70] = 255
    cc = blobs.ConnectedComponents(w, h)
    for force in (0, 1):
        cc.set("forceBackgroundPass", force)
        for m in (nested, plain, tricky):
            for zb in (False, True):
                n, lab, comps = cc.label(m, zero_border=zb)
                on, olab, ost, oext = oracle.ccl8(m, zb)
                assert n == on and np.array_equal(lab, olab)
                assert [c["external"] for c in comps] == [int(e) for e in oext], (force, zb)
    assert sum(1 - int(e) for e in oracle.ccl8(nested)[3]) >= 2          # the fixture really has nested components
    cc.close()


def test_ccl_batch_of_images(oracle):
    """One launch sequence for a group of masks (one per camera stream) == per-image results."""
    import torch
    from tracking_b200 import blobs
    rng = np.random.default_rng(9)
    S, h, w = 5, 97, 203
    ms = np.stack([(rng.random((h, w)) < p).astype(np.uint8) * 255 for p in (0.02, 0.3, 0.5, 0.7, 0.0)])
    d = torch.from_numpy(ms).cuda()
    lab = torch.zeros((S, h, w), dtype=torch.int32, device="cuda")
    cc = blobs.ConnectedComponents(w, h, max_images=S)
    for zb in (False, True):
        cc.label_batch_dev(d.data_ptr(), w, h, S, zb, lab.data_ptr())
        torch.cuda.synchronize()
        for i in range(S):
            on, olab, ost, oext = oracle.ccl8(ms[i], zb)
            comps = cc.components(i)
            assert len(comps) == on and np.array_equal(lab[i].cpu().numpy(), olab)
            assert [c["external"] for c in comps] == [int(e) for e in oext]
            assert [c["area"] for c in comps] == [int(s[4]) for s in ost]
    cc.close()


def test_ccl_golden_tables(oracle, clips, golden):
    """BASELINE config 1 shape: FD -> OPEN -> CC on the reference video clip vs committed OpenCV results."""
    import tracking_b200 as tb
    from tracking_b200 import blobs
    frames = list(clips["video_clip"])
    fd = tb.FrameDifferenceBGS()
    fgs = [m for m in (fd.process(f)[0] for f in frames) if m is not None]
    h, w = fgs[0].shape
    cc = blobs.ConnectedComponents(w, h)
    for zb in (0, 1):
        for row in golden["sequences"]["video_clip"]["fd_open_ccl"]["zero_border_%d" % zb]:
            m = blobs.morph(fgs[row["frame"]], [("erode", 1), ("dilate", 1)])
            assert sha([m]) == row["open_sha256"]
            n, lab, comps = cc.label(m, zero_border=bool(zb))
            assert n == row["n_components"] and sha([lab]) == row["labels_sha256"]
            ext = [[c["x"], c["y"], c["w"], c["h"]] for c in comps if c["external"]]
            assert list(reversed(ext)) == row["external_rects_findcontours_order"]


def test_rect_moments_exact(oracle):
    from tracking_b200 import blobs
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (300, 500), dtype=np.uint8)
    cc = blobs.ConnectedComponents(500, 300)
    cc.label(img, want_labels=False)
    rects = [(0, 0, 500, 300), (3, 5, 40, 33), (498, 299, 2, 1), (100, 100, 1, 150), (7, 9, 333, 2)]
    got = cc.rect_moments(rects)
    for r, g in zip(rects, got):
        assert g == oracle.rect_moments(img, r)


def test_blobdetector_matches_restated_oracle(clips, golden):
    """Whole DetectNewBlob sequence (cluster, filter, sort, top-10, trajectories) vs oracle/blobdetect.py.
    The list logic is 'restated, unpinned' on the oracle side (no OpenCV legacy module in this image)."""
    import tracking_b200 as tb
    from tracking_b200 import blobs
    from oracle import blobdetect, cv2_chain
    frames = list(clips["video_clip"])
    fd = tb.FrameDifferenceBGS()
    bd = blobs.CvBlobDetectorCC()
    ob = blobdetect.CvBlobDetectorCC(zero_border=True)
    exp = golden["sequences"]["video_clip"]["fd_open_blobdetector_restated_unpinned"]
    k = 0
    tracked = []
    for f in frames:
        fg, _ = fd.process(f)
        if fg is None:
            continue
        m = blobs.morph(fg, [("erode", 1), ("dilate", 1)])
        res, nb = bd.DetectNewBlob(m, tracked)
        ores, onb = ob.DetectNewBlob(cv2_chain.morph(cv2_chain.morph(fg, "erode"), "dilate"),
                                     [blobdetect.Blob(*t) for t in tracked])
        assert res == ores
        assert len(bd.frame_blobs) == len(ob.lists[0])
        for a, b in zip(bd.frame_blobs, ob.lists[0]):
            assert a == b.tuple()                       # fp32 fields, bit-exact
        if res:
            assert nb == onb.tuple()
        if not tracked:
            e = exp[k]
            assert res == e["result"]
            assert [[round(v, 4) for v in b] for b in bd.frame_blobs] == e["frame_blobs"]
        k += 1
    # with tracked blobs present, overlapping detections are suppressed
    bd2, ob2 = blobs.CvBlobDetectorCC(), blobdetect.CvBlobDetectorCC(zero_border=True)
    fd2 = tb.FrameDifferenceBGS()
    tracked = [(150.0, 80.0, 40.0, 40.0)]
    for f in frames:
        fg, _ = fd2.process(f)
        if fg is None:
            continue
        m = blobs.morph(fg, [("erode", 1), ("dilate", 1)])
        res, nb = bd2.DetectNewBlob(m, tracked)
        ores, onb = ob2.DetectNewBlob(m, [blobdetect.Blob(*t) for t in tracked])
        assert res == ores and [a for a in bd2.frame_blobs] == [b.tuple() for b in ob2.lists[0]]


def test_pipeline_on_device_mog2_open_cc(oracle, clips):
    """BASELINE config 4 shape on the clip: MOG2 -> OPEN -> CC with every buffer resident in HBM."""
    import torch
    import tracking_b200 as tb
    from tracking_b200 import blobs
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    p, o = tb.MixtureOfGaussianV2BGS(), oracle.MixtureOfGaussianV2BGS()
    cc = blobs.ConnectedComponents(w, h)
    d_fg = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    d_open = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    d_lab = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for f in clip[:20]:
        d_in = torch.from_numpy(f).cuda()
        p.process_dev(d_in.data_ptr(), w, h, d_fg.data_ptr(), None, stream=st)
        blobs.morph_dev(d_fg.data_ptr(), w, h, 1, [("erode", 1), ("dilate", 1)], d_open.data_ptr(), stream=st)
        cc.label_dev(d_open.data_ptr(), w, h, True, d_lab.data_ptr(), stream=st)
        comps = cc.components()
        ofg, _ = o.process(f)
        om = oracle.morph(oracle.morph(ofg, "erode"), "dilate")
        on, olab, ost, oext = oracle.ccl8(om, True)
        assert np.array_equal(d_open.cpu().numpy(), om)
        assert len(comps) == on and np.array_equal(d_lab.cpu().numpy(), olab)


@pytest.mark.parametrize("shape", [(70, 128), (65, 160), (130, 96), (257, 416), (64, 2080), (300, 1920)])
def test_ccl_label_tiles_word_aligned_widths(oracle, shape):
    """Widths that are multiples of 32 take the label kernel's staged 512-byte row stores where a tile column is whole and
    the per-thread stores where it is not (wpr % 4 != 0), tile rows past the image bottom, single image and a batch big
    enough for the four-words-per-thread form; blobs larger than a tile, specks, and a noise image (per-run path)."""
    import torch
    from tracking_b200 import blobs
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    yy, xx = np.mgrid[0:h, 0:w]
    blobby = np.zeros((h, w), np.uint8)
    for _ in range(5):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(5, 90)
        blobby[((yy - cy) ** 2 + ((xx - cx) // 2) ** 2) < r * r] = 255
    blobby[rng.random((h, w)) < 0.003] = 255
    noise = (rng.random((h, w)) < 0.35).astype(np.uint8) * 255
    cc1 = blobs.ConnectedComponents(w, h)
    for m in (blobby, noise):
        for zb in (False, True):
            n, lab, comps = cc1.label(m, zero_border=zb)
            on, olab, ost, oext = oracle.ccl8(m, zb)
            assert n == on and np.array_equal(lab, olab), (shape, zb)
            assert [(c["x"], c["y"], c["w"], c["h"], c["area"]) for c in comps] == \
                [(int(s[0]), int(s[1]), int(s[2] - s[0] + 1), int(s[3] - s[1] + 1), int(s[4])) for s in ost]
    cc1.close()
    S = max(2, 2400 // max(1, ((w + 31) // 32 * h + 255) // 256))            # enough 256-word chunks for the batch form
    S = min(S, 48)
    ms = np.stack([np.roll(blobby if s % 3 else noise, 5 * s, axis=1) for s in range(S)])
    d = torch.from_numpy(ms).cuda()
    lab = torch.full((S, h, w), -7, dtype=torch.int32, device="cuda")
    ccS = blobs.ConnectedComponents(w, h, max_images=S)
    for with_labels in (True, False):
        ccS.label_batch_dev(d.data_ptr(), w, h, S, False, lab.data_ptr() if with_labels else None)
        torch.cuda.synchronize()
        for i in (0, 1, S // 2, S - 1):
            on, olab, ost, oext = oracle.ccl8(ms[i], False)
            comps = ccS.components(i)
            assert len(comps) == on
            assert [c["area"] for c in comps] == [int(s[4]) for s in ost]
            assert [c["external"] for c in comps] == [int(e) for e in oext]
            if with_labels:
                assert np.array_equal(lab[i].cpu().numpy(), olab), (shape, i)
    ccS.close()
