"""One 1080p frame up (6.2 MB) / mask + background down (8.3 MB) from page-locked memory, each transfer as k pieces on k
streams (BGSB_PROBE_SPLIT=k): does one copy engine per direction saturate the link?  Also a 4x larger transfer.
GPU box, measurement tooling."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from tracking_b200 import capi
    NPX = 1920 * 1080
    for mult in (1, 4):
        up, dn, both = C.c_double(0), C.c_double(0), C.c_double(0)
        capi.check(capi.lib().bgsb_copy_probe(0, NPX * 3 * mult, NPX * 4 * mult, 200, C.byref(up), C.byref(dn), C.byref(both)))
        print(json.dumps({"split": int(sys.argv[1]), "frames_per_copy": mult, "h2d_gbs": NPX * 3 * mult / up.value / 1e9,
                          "d2h_gbs": NPX * 4 * mult / dn.value / 1e9, "up_us": up.value * 1e6 / mult, "down_us": dn.value * 1e6 / mult,
                          "both_us": both.value * 1e6 / mult}))
else:
    for k in (1, 2, 3, 4):
        subprocess.run([sys.executable, __file__, str(k)], env=dict(os.environ, BGSB_PROBE_SPLIT=str(k)), check=True)
