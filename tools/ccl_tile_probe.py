"""The labeller alone: one 1080p mask of three kinds (12 blobs, 0.2 % salt noise, 30 % random) and a batch of 64 blob masks,
table only and with the label image.  (Round 2 used it for the A/B of the raster / tiled label and merge kernels; only the
forms that won are in the library.)  GPU box, measurement tooling."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.argv = ["x"]
import torch
import bench_configs as b
torch.cuda.set_device(0)
b.ccl_kernel_probe()
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ccl_batch_probe.py")], check=True)
