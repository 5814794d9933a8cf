"""ABL, 16 x 1080p streams: warp-coalesced table kernel with / without the quiet-radius shortcut, short and long
frame rings.  GPU box, measurement tooling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = ["x"]
import torch
import bench_configs as b
import tracking_b200 as tb
torch.cuda.set_device(0)


class NoQuiet(tb.AdaptiveBackgroundLearning):
    def __init__(self, **kw):
        super().__init__(quietGroups=0, **kw)


b.simple_streams(tb.AdaptiveBackgroundLearning, "ABL quiet", 10)
b.simple_streams(tb.AdaptiveBackgroundLearning, "ABL quiet NT=24", 10, NT=24)
b.simple_streams(NoQuiet, "ABL noquiet", 10)
b.simple_streams(NoQuiet, "ABL noquiet NT=24", 10, NT=24)
