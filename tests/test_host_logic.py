"""CPU-side checks of test infrastructure and host logic added in round 2 (no GPU, no compute calls into the library)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_churn_video_twins_agree(oracle):
    """The mode-churn generator: numpy twin == C oracle twin (the device kernel is checked against them under -m gpu)."""
    from tracking_b200 import synth
    for (w, h, t, seed) in ((200, 120, 7, 1237), (33, 17, 0, 1234), (64, 64, 1001, 99)):
        assert np.array_equal(synth.churn_frame(w, h, t, seed), oracle.synth_churn_frame(w, h, t, seed))
    # what the generator is for: every pixel keeps (nearly) all five modes live
    m = oracle.MixtureOfGaussianV2BGS()
    for t in range(40):
        m.process(oracle.synth_churn_frame(160, 96, t, 1234))
    assert m.nmodes.mean() > 4.5 and (m.nmodes == 5).mean() > 0.6


def test_committed_bench_hashes_match_the_oracle(oracle):
    """tests/golden/bench_hashes.json (what bench.py checks its timed paths against) is what the oracle produces today:
    one MOG2 stream and one pipeline stream are recomputed."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_bench_hashes as mk
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_hashes.json")))
    assert gold["geometry"] == [mk.W, mk.H] and gold["seed0"] == mk.SEED0
    assert len(gold["mog2"]["by_seed"]) == mk.MOG2_SEEDS and len(gold["pipeline"]["by_seed"]) == mk.PIPE_STREAMS
    seed = mk.SEED0 + 3
    assert mk.mog2_stream(seed) == gold["mog2"]["by_seed"][str(seed)]
    seed = mk.SEED0 + 41
    assert mk.pipe_stream(seed) == gold["pipeline"]["by_seed"][str(seed)]


def test_abl_opencv24_table_properties(oracle):
    """The unpinned OpenCV 2.4 blend restatement: identity on equal bytes, monotone, close to the pinned 4.x blend."""
    t = oracle.abl_blend_table_24(0.05)
    assert t.shape == (256, 256) and all(t[i, i] == i for i in range(256))
    assert (np.diff(t.astype(np.int16), axis=0) >= 0).all() and (np.diff(t.astype(np.int16), axis=1) >= 0).all()
    x = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    y = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    o = oracle.AdaptiveBackgroundLearning()
    o.process(np.stack([y, y, y], -1))
    _, bg = o.process(np.stack([x, x, x], -1))
    assert np.abs(bg[..., 0].astype(np.int16) - t.astype(np.int16)).max() <= 1


STUB_TEST = r'''
#include <stdio.h>
#include "opencv2/legacy/blobtrack.hpp"
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)
int main(int argc, char **argv)
{
    const std::string dir = argv[1];
    // cv::Mat: create, shared ownership, copyTo, header over an IplImage and back
    cv::Mat a(3, 5, CV_8UC3);
    CHECK(!a.empty() && a.rows == 3 && a.cols == 5 && a.channels() == 3 && (size_t)a.step == 15 && a.isContinuous());
    for (int i = 0; i < 45; i++) a.data[i] = (unsigned char)i;
    cv::Mat b = a;                       // shares
    b.data[0] = 99; CHECK(a.data[0] == 99);
    cv::Mat c; a.copyTo(c); c.data[0] = 1; CHECK(a.data[0] == 99 && c.data[44] == 44);
    unsigned char *before = c.data; c.create(3, 5, CV_8UC3); CHECK(c.data == before);      // same geometry: kept
    cv::Mat e; CHECK(e.empty()); e.copyTo(c); CHECK(c.empty());
    unsigned char raw[4 * 8];
    for (int i = 0; i < 32; i++) raw[i] = (unsigned char)(i * 3);
    IplImage ipl = {1, IPL_DEPTH_8U, 6, 4, 8, (char *)raw};            // 6 px wide, 8-byte rows
    cv::Mat h(&ipl); CHECK(h.data == raw && (size_t)h.step == 8 && !h.isContinuous() && h.type() == CV_8UC1);
    cv::Mat hc(&ipl, true); CHECK(hc.data != raw && hc.isContinuous() && hc.data[6] == raw[8]);
    IplImage back = hc; CHECK(back.width == 6 && back.height == 4 && back.widthStep == 6 && back.imageData == (char *)hc.data);
    try { CV_Assert(1 == 2); CHECK(false); } catch (const cv::Exception &x) { CHECK(std::string(x.what()).find("1 == 2") != std::string::npos); }
    // CvFileStorage: missing file -> defaults; write -> read back; unknown key -> default
    const std::string path = dir + "/cfg.xml";
    CHECK(cvOpenFileStorage(path.c_str(), 0, CV_STORAGE_READ) == 0);
    CHECK(cvReadIntByName(0, 0, "threshold", 15) == 15 && cvReadRealByName(0, 0, "alpha", 0.05) == 0.05);
    CvFileStorage *fs = cvOpenFileStorage(path.c_str(), 0, CV_STORAGE_WRITE);
    CHECK(fs != 0);
    cvWriteReal(fs, "alpha", 0.0123456789012345); cvWriteInt(fs, "threshold", 40); cvWriteInt(fs, "showOutput", 0);
    cvReleaseFileStorage(&fs); CHECK(fs == 0);
    fs = cvOpenFileStorage(path.c_str(), 0, CV_STORAGE_READ);
    CHECK(fs != 0 && cvReadIntByName(fs, 0, "threshold", 15) == 40 && cvReadIntByName(fs, 0, "showOutput", 1) == 0);
    CHECK(cvReadRealByName(fs, 0, "alpha", 0.05) == 0.0123456789012345 && cvReadIntByName(fs, 0, "nosuchkey", 7) == 7);
    cvReleaseFileStorage(&fs);
    CHECK(cvOpenFileStorage((dir + "/nodir/x.xml").c_str(), 0, CV_STORAGE_WRITE) == 0);
    // CvBlobSeq
    CvBlobSeq seq; CvBlob bl = cvBlob(1, 2, 3, 4); seq.AddBlob(&bl); seq.AddBlob(&bl);
    CHECK(seq.GetBlobNum() == 2 && seq.GetBlob(1)->w == 3 && seq.GetBlob(2) == 0);
    printf("stand-in ok\n");
    return 0;
}
'''


def test_opencv_stand_in_containers(tmp_path):
    """adapters/stub_opencv is functional (the executed C++ drop-in test rests on it): cv::Mat ownership / headers,
    CvFileStorage round trip and defaults, CvBlobSeq."""
    src = tmp_path / "stub_test.cpp"
    src.write_text(STUB_TEST)
    exe = str(tmp_path / "stub_test")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tracking_b200", "adapters", "stub_opencv"),
                           str(src), "-o", exe])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "stand-in ok" in r.stdout, r.stdout + r.stderr
    xml = (tmp_path / "cfg.xml").read_text()
    assert xml.startswith("<?xml") and "<threshold>40</threshold>" in xml


def test_stream_pool_sharding_rule_matches_shard_streams():
    """bgsb_pool puts stream s on devices[s % ndevices] (csrc/pool.cu); shard_streams is the same rule across processes."""
    from tracking_b200 import streams
    src = open(os.path.join(ROOT, "tracking_b200", "csrc", "pool.cu")).read()
    assert "for (int s = g; s < nstreams; s += ng)" in src
    for n, g in ((64, 8), (5, 2), (3, 4)):
        ng = min(n, g)
        for r in range(ng):
            assert streams.shard_streams(n, ng, r) == list(range(r, n, ng))


def test_gray24_golden_matches_the_oracle_variant(oracle, clips):
    """tests/golden/golden_gray24.json (pure-numpy restatement of FD / StaticFD with the OpenCV 2.4 gray constants) ==
    the C oracle's gray_variant = 1, so the GPU test that checks `grayVariant` 1 against the oracle is anchored to an
    independent statement of the formula."""
    import hashlib
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_gray24.json")))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_gray24 as mk
    ties, n = mk.tie_frames()
    assert n == gold["gray_ties"]["colours"] and n >= 1
    for variant in (0, 1):
        o = oracle.FrameDifferenceBGS(gray_variant=variant)
        o.process(ties[0])
        fg, _ = o.process(ties[1])
        assert hashlib.sha256(fg.tobytes()).hexdigest() == gold["gray_ties"]["FrameDifferenceBGS:grayVariant=%d" % variant]
    for name, clip in clips.items():
        for aid, key in ((0, "FrameDifferenceBGS:grayVariant=1"), (1, "StaticFrameDifferenceBGS:grayVariant=1")):
            o = oracle.ALGOS[aid](gray_variant=1)
            h = hashlib.sha256()
            for f in clip:
                fg, _ = o.process(f)
                if fg is not None:
                    h.update(np.ascontiguousarray(fg).tobytes())
            assert h.hexdigest() == gold[name][key], (name, key)
