"""ncu target: the config-4 pipeline object on S x 1080p streams -- 30 frames to settle, then a few more.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/pipeline_ncu_target.py 16"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracking_b200 import synth
from tracking_b200.pipeline import ForegroundPipeline
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H, NT = 1920, 1080, 12
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NT, S, H, W, 3), dtype=torch.uint8, device="cuda")
for t in range(NT):
    synth.frames_dev(d[t].data_ptr(), S, 1, W, H, t0=t, stream=st)
pipe = ForegroundPipeline(5, nstreams=S)
for k in range(36):
    pipe.process_dev(d[k % NT].data_ptr(), W, H, None, None, None, stream=st)
pipe.join_dev(st)
torch.cuda.synchronize()
print("components", len(pipe.components(0)))
