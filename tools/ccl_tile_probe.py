"""A/B of the labeller's kernel forms (BGSB_CCL_TILE bit 0: tiled label kernel, bit 1: tiled merge kernel): one 1080p mask of
three kinds, and a batch of 64.  GPU box, measurement tooling."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.argv = ["x"]
    import torch
    import bench_configs as b
    torch.cuda.set_device(0)
    b.ccl_kernel_probe()
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ccl_batch_probe.py")], check=True)
else:
    for mode in (3, 1, 2, 0):
        print("BGSB_CCL_TILE=%d" % mode, flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, BGSB_CCL_TILE=str(mode)), check=True)
