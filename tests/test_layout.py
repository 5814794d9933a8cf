"""The oracle is test infrastructure: nothing shipped may import, link or execute it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(d, exts):
    for base, _, names in os.walk(os.path.join(ROOT, d)):
        if "_obj" in base or "__pycache__" in base:
            continue
        for n in names:
            if n.endswith(exts):
                yield os.path.join(base, n)


def test_product_never_touches_oracle():
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|bgs_oracle|cv2)")
    bad = []
    for p in _files("tracking_b200", (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
        for i, line in enumerate(open(p, encoding="utf-8", errors="replace"), 1):
            if pat.search(line) and "oracle/c/bgs_oracle.c: orc_synth_frame" not in line \
                    and "C oracle" not in line and "opencv2" not in line:
                bad.append("%s:%d: %s" % (os.path.relpath(p, ROOT), i, line.strip()))
    assert not bad, "product path references the oracle / cv2:\n" + "\n".join(bad)


def test_required_layout():
    for p in ("include/bgsb200.h", "oracle/c/bgs_oracle.c", "oracle/Makefile", "tests/golden/golden.json",
              "tests/golden/clips.npz", "tests/golden/make_golden.py", "bench.py", "__graft_entry__.py",
              "DESIGN.md", "INTEGRATION.md"):
        assert os.path.exists(os.path.join(ROOT, p)), p


def test_no_reference_reads_at_gpu_runtime():
    """-m gpu tests, smoke() and bench.py must not read /root/reference."""
    for p in ["bench.py", "__graft_entry__.py"] + [f for f in _files("tests", (".py",)) if "test_gpu" in f]:
        p = p if os.path.isabs(p) else os.path.join(ROOT, p)
        if os.path.exists(p):
            assert "/root/reference" not in open(p).read(), p
