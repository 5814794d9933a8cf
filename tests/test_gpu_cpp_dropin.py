"""The C++ drop-in, EXECUTED: tracking_b200/adapters/dropin_test.cpp drives the adapter classes with the reference's own
call patterns (new <Plugin>; bgs->process(in, fg, bg); USTC_BGS::Process / GetMask; DetectNewBlob; the tracker's second
look at the mask), compiled against the functional OpenCV stand-in (adapters/stub_opencv: containers only) and linked
against libbgsb200.so.  What it writes is compared with the cv2-generated golden hashes (stand-in claiming OpenCV 4: the
arithmetic the goldens were produced with) and with the oracle's OpenCV 2.4 variants (stand-in claiming 2.4, like the
reference's real build)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AD = os.path.join(ROOT, "tracking_b200", "adapters")
NAMES = {0: "FrameDifferenceBGS", 1: "StaticFrameDifferenceBGS", 2: "WeightedMovingMeanBGS",
         3: "WeightedMovingVarianceBGS", 5: "MixtureOfGaussianV2BGS", 6: "AdaptiveBackgroundLearning"}


def build(tmp, cv_major):
    exe = os.path.join(tmp, "dropin_test_cv%d" % cv_major)
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-DBGSB_STUB_CV_MAJOR=%d" % cv_major,
                           "-I", os.path.join(AD, "stub_opencv"), "-I", AD, "-I", os.path.join(ROOT, "include"),
                           os.path.join(AD, "dropin_test.cpp"), "-L", os.path.join(ROOT, "tracking_b200"), "-lbgsb200",
                           "-Wl,-rpath," + os.path.join(ROOT, "tracking_b200"), "-o", exe])
    return exe


def run(exe, clip, outdir, config=None):
    os.makedirs(os.path.join(outdir, "config"), exist_ok=True)
    for name, text in (config or {}).items():
        with open(os.path.join(outdir, "config", name + ".xml"), "w") as f:
            f.write(text)
    raw = os.path.join(outdir, "clip.raw")
    with open(raw, "wb") as f:
        f.write(np.array([clip.shape[0], clip.shape[1], clip.shape[2]], np.int32).tobytes())
        f.write(np.ascontiguousarray(clip).tobytes())
    r = subprocess.run([exe, raw, outdir], cwd=outdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def read(outdir, name, shape):
    p = os.path.join(outdir, name)
    if not os.path.exists(p):
        return np.zeros((0,) + shape, np.uint8)
    return np.fromfile(p, np.uint8).reshape((-1,) + shape)


def sha_file(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def oracle_outputs(o, clip):
    fgs, bgs = [], []
    for f in clip:
        fg, bg = o.process(f)
        if fg is not None:
            fgs.append(fg)
        if bg is not None:
            bgs.append(bg)
    return fgs, bgs


def test_cpp_dropin_opencv4_arithmetic_matches_golden_hashes(tmp_path, clips, golden, oracle):
    import cv2
    from tracking_b200 import blobs
    exe = build(str(tmp_path), 4)
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    out = str(tmp_path / "run4")
    stdout = run(exe, clip, out)
    for ctor in ("FrameDifferenceBGS()", "~FrameDifferenceBGS()", "MixtureOfGaussianV2BGS()", "~AdaptiveBackgroundLearning()"):
        assert ctor in stdout                                   # the reference's stdout banners (e.g. FrameDifferenceBGS.cpp:21,26)
    algos = golden["sequences"]["video_clip"]["algos"]
    for aid, name in NAMES.items():
        exp = algos[name]
        assert sha_file(os.path.join(out, name + ".fg")) == exp["fg_sha256"], name
        if exp["bg_sha256"]:
            assert sha_file(os.path.join(out, name + ".bg")) == exp["bg_sha256"], name
        else:
            assert not os.path.exists(os.path.join(out, name + ".bg")), name      # FD / WMV never write img_bgmodel
    # the DP package's plugins (USTC_BGS types 9 / 12 / 13) against the masks a build of the reference's own sources produced
    gdp = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_dp.json")))["plugins"]
    for name in ("DPAdaptiveMedianBGS", "DPMeanBGS", "DPWrenGABGS", "DPPratiMediodBGS", "SigmaDeltaBGS"):
        assert name + "()" in stdout and "~" + name + "()" in stdout
        assert sha_file(os.path.join(out, name + ".fg")) == gdp[name]["video_clip"]["params"]["{}"]["masks_sha256"], name
        assert not os.path.exists(os.path.join(out, name + ".bg")), name
        xml = open(os.path.join(out, "config", name + ".xml")).read()
        keys = {"DPAdaptiveMedianBGS": ("threshold", "samplingRate", "learningFrames"), "DPPratiMediodBGS": ("threshold", "samplingRate", "historySize", "weight"),
                "SigmaDeltaBGS": ("ampFactor", "minVar", "maxVar")}
        for key in ("showOutput",) + keys.get(name, ("threshold", "alpha", "learningFrames")):
            assert "<%s>" % key in xml, (name, key)
    # fan-out: the same masks with one upload per frame
    for name in ("FrameDifferenceBGS", "WeightedMovingVarianceBGS", "MixtureOfGaussianV2BGS", "AdaptiveBackgroundLearning"):
        assert sha_file(os.path.join(out, "fan_" + name + ".fg")) == algos[name]["fg_sha256"], name
    # saveConfig wrote the reference's XML keys on the first frame
    xml = open(os.path.join(out, "config", "MixtureOfGaussianV2BGS.xml")).read()
    for key in ("alpha", "enableThreshold", "threshold", "showOutput"):
        assert "<%s>" % key in xml
    # USTC_BGS (FrameDifference): GetMask() is NULL on the warm-up frame, then the plugin's mask
    masks = read(out, "ustc.fg", (h, w))
    assert hashlib.sha256(masks.tobytes()).hexdigest() == algos["FrameDifferenceBGS"]["fg_sha256"]
    opened = read(out, "ustc_open.fg", (h, w))
    lines = open(os.path.join(out, "blobs.txt")).read().split("\n")
    assert lines[0] == "0 nomask"
    # DetectNewBlob through the C++ class == the same detector through the Python mirror (zeroBorder 0: OpenCV >= 3.2)
    bd = blobs.CvBlobDetectorCC(zeroBorder=0)
    feed = open(os.path.join(out, "feed.txt")).read().strip().split("\n")
    assert len(feed) == len(opened)
    for i, m in enumerate(opened):
        assert np.array_equal(m, oracle.morph(oracle.morph(masks[i], "erode", 1), "dilate", 1))
        res, nb = bd.DetectNewBlob(m, [])
        parts = lines[i + 1].split()
        assert int(parts[0]) == i + 1 and int(parts[1]) == res
        if res:
            assert [np.float32(v) for v in parts[2:6]] == [np.float32(v) for v in nb]
        # the tracker's second look (BgsbTrackerFeed) against OpenCV itself: cvFindContours(RETR_EXTERNAL) rectangles
        # in its order, cvSum of the mask under each, and of the whole mask
        head, sums, total = feed[i].split("|")
        toks = head.split()
        contours, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        assert int(toks[1]) == len(contours)
        got = [tuple(int(v) for v in t.split(",")) for t in toks[2:]]
        assert got == [tuple(cv2.boundingRect(c)) for c in contours]
        assert [int(v) for v in sums.split()] == [int(cv2.sumElems(m[y:y + hh, x:x + ww])[0]) for (x, y, ww, hh) in got]
        assert int(total) == int(m.sum())


def test_cpp_dropin_opencv24_arithmetic_and_xml_config(tmp_path, clips, oracle):
    """The stand-in claiming OpenCV 2.4 (the reference's real build): BGR2GRAY with the 2.4 constants, the fp32
    addWeighted of AdaptiveBackgroundLearning (parity unpinned), cvFindContours clearing the border -- and a config file
    that is present before the first frame is honoured (loadConfig runs before saveConfig)."""
    exe = build(str(tmp_path), 2)
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    out = str(tmp_path / "run2")
    cfg = {"FrameDifferenceBGS": "<?xml version=\"1.0\"?>\n<opencv_storage>\n<enableThreshold>1</enableThreshold>\n"
                                 "<threshold>40</threshold>\n<showOutput>0</showOutput>\n</opencv_storage>\n"}
    run(exe, clip, out, cfg)
    for aid, kw in ((0, {"threshold": 40}), (1, {}), (2, {}), (3, {})):
        fgs, bgs = oracle_outputs(oracle.ALGOS[aid](gray_variant=1, **kw), clip)
        assert np.array_equal(read(out, NAMES[aid] + ".fg", (h, w)), np.stack(fgs)), NAMES[aid]
        if bgs:
            assert np.array_equal(read(out, NAMES[aid] + ".bg", (h, w, 3)), np.stack(bgs)), NAMES[aid]
    fgs, bgs = oracle_outputs(oracle.MixtureOfGaussianV2BGS(), clip)      # MOG2 has no version-dependent arithmetic
    assert np.array_equal(read(out, "MixtureOfGaussianV2BGS.fg", (h, w)), np.stack(fgs))
    assert np.array_equal(read(out, "MixtureOfGaussianV2BGS.bg", (h, w, 3)), np.stack(bgs))
    # AdaptiveBackgroundLearning as OpenCV 2.4 computes it, restated in numpy (UNPINNED): bg <- table[in, bg],
    # mask = gray24(|in - bg_old|) > 15
    table = oracle.abl_blend_table_24(0.05)
    bg = clip[0].copy()
    fg_exp, bg_exp = [], []
    for f in clip:
        d = np.abs(f.astype(np.int16) - bg.astype(np.int16)).astype(np.uint8)
        fg_exp.append(((oracle.gray_bgr(d, 1) > 15) * 255).astype(np.uint8))
        bg = table[f, bg]
        bg_exp.append(bg.copy())
    assert np.array_equal(read(out, "AdaptiveBackgroundLearning.fg", (h, w)), np.stack(fg_exp))
    assert np.array_equal(read(out, "AdaptiveBackgroundLearning.bg", (h, w, 3)), np.stack(bg_exp))
    # ustc.fg: USTC_BGS(0) = FrameDifference, which read the same config file
    fgs, _ = oracle_outputs(oracle.FrameDifferenceBGS(gray_variant=1, threshold=40), clip)
    assert np.array_equal(read(out, "ustc.fg", (h, w)), np.stack(fgs))
