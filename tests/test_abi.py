"""The C-ABI library loads and exports every symbol include/bgsb200.h declares (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bgsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"BGSB_API\s+[^;{]*?\b(bgsb_\w+)\s*\(", src)))


def test_build_and_exports():
    from tracking_b200 import _build, capi
    _build.build()
    names = _declared()
    assert len(names) >= 30
    lib = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libbgsb200.so does not export %s" % n
    assert sorted(capi.SIGNATURES) == names, "capi.SIGNATURES and the header disagree"


def test_only_abi_symbols_are_exported():
    from tracking_b200 import capi
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    extra = [s for s in syms if not s.startswith("bgsb_") and s not in ("_init", "_fini")]
    assert not extra, extra


def test_argument_errors_without_gpu():
    from tracking_b200 import capi
    L = capi.lib()
    assert L.bgsb_version().startswith(b"bgsb200")
    assert L.bgsb_set_param(None, b"alpha", 0.1) == capi.ERR_ARG
    assert b"null" in L.bgsb_last_error()
    h = C.c_void_p()
    assert L.bgsb_create(C.byref(h), 36, 0) == capi.ERR_ARG       # SuBSENSE id: not on the hot path
    assert L.bgsb_create_group(C.byref(h), 5, 0, 0) == capi.ERR_ARG
    assert L.bgsb_morph_dev(None, 4, 4, 1, None, 0, None, None) == capi.ERR_ARG
    assert L.bgsb_kernel_launch_count() == 0


def test_fails_loudly_when_extension_missing(tmp_path, monkeypatch):
    from tracking_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(ImportError, match="no CPU fallback"):
        capi.lib()


def test_adapters_compile_against_stub_opencv():
    """The header-only C++ adapters keep the reference's class shapes (IBGS, CvFGDetector...)."""
    test_cpp = os.path.join(ROOT, "tracking_b200", "adapters", "compile_check.cpp")
    if not os.path.exists(test_cpp):
        import pytest
        pytest.skip("adapters not present yet")
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra",
                           "-I", os.path.join(ROOT, "tracking_b200", "adapters", "stub_opencv"),
                           "-I", os.path.join(ROOT, "tracking_b200", "adapters"),
                           "-I", os.path.join(ROOT, "include"), test_cpp])


def test_dropin_test_program_links_against_the_library(tmp_path):
    """adapters/dropin_test.cpp (run under -m gpu by tests/test_gpu_cpp_dropin.py) compiles against the functional
    OpenCV stand-in, for both OpenCV generations it can claim, and links against libbgsb200.so."""
    ad = os.path.join(ROOT, "tracking_b200", "adapters")
    for major in (2, 4):
        subprocess.check_call(["g++", "-std=c++11", "-O0", "-Wall", "-Wextra", "-DBGSB_STUB_CV_MAJOR=%d" % major,
                               "-I", os.path.join(ad, "stub_opencv"), "-I", ad, "-I", os.path.join(ROOT, "include"),
                               os.path.join(ad, "dropin_test.cpp"), "-L", os.path.join(ROOT, "tracking_b200"), "-lbgsb200",
                               "-Wl,-rpath," + os.path.join(ROOT, "tracking_b200"), "-o", str(tmp_path / ("dropin%d" % major))])
