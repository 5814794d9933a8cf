"""WMV, 16 x 1080p streams: default device path, retainInput with / without the quiet-group shortcut, short and long frame
rings (the ring wrap makes the scene jump: more busy groups).  GPU box, measurement tooling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = ["x"]
import torch
import bench_configs as b
import tracking_b200 as tb
torch.cuda.set_device(0)
b.simple_streams(tb.WeightedMovingVarianceBGS, "WMV", 16)
b.simple_streams(tb.WeightedMovingVarianceBGS, "WMV retainInput quiet", 10, retain=True)
b.simple_streams(tb.WeightedMovingVarianceBGS, "WMV retainInput quiet NT=24", 10, retain=True, NT=24)


class NoQuiet(tb.WeightedMovingVarianceBGS):
    def __init__(self, **kw):
        super().__init__(quietGroups=0, **kw)


b.simple_streams(NoQuiet, "WMV retainInput noquiet", 10, retain=True)
