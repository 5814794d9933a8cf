"""Pipeline frame-set time vs the cap on the clean-up / labelling grids ("chainCtas").  python tools/chain_ctas_probe.py S"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracking_b200 import synth
from tracking_b200.pipeline import ForegroundPipeline
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w, h, NT = 1920, 1080, 24
st = torch.cuda.current_stream().cuda_stream
frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
for t in range(NT): synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
for cap in (0, 148, 296, 592, 1184, 2368):
    pipe = ForegroundPipeline(5, nstreams=S, chainCtas=cap)
    k = [0]
    def step():
        pipe.process_dev(frames[k[0] % NT].data_ptr(), w, h, None, None, None, stream=st); k[0] += 1
    for _ in range(4 * NT): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(48): step()
    pipe.join_dev(st); e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(streams=S, chainCtas=cap, us_per_step=e0.elapsed_time(e1) / 48 * 1e3)), flush=True)
    pipe.close()
