"""Config 4 stage timings: packed pipeline object vs the byte-mask chain, labeller alone (single image / 64-batch,
with and without the label image).  GPU box, measurement tooling.  python tools/pipeline_probe.py [S]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tracking_b200 as tb
from tracking_b200 import blobs, synth
from tracking_b200.pipeline import ForegroundPipeline

def timed(fn, iters, warm=3, join=None):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    if join: join()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3     # us

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w, h, NT = 1920, 1080, 24
st = torch.cuda.current_stream().cuda_stream
frames = torch.empty((NT, S, h, w, 3), dtype=torch.uint8, device="cuda")
for t in range(NT): synth.frames_dev(frames[t].data_ptr(), S, 1, w, h, t0=t, stream=st)
# every variant: a fresh model, 4 passes over the 24-frame ring to settle, then 2 timed passes (same frames for all)
pipe = ForegroundPipeline(5, nstreams=S)
k = [0]
def step():
    pipe.process_dev(frames[k[0] % NT].data_ptr(), w, h, None, None, None, stream=st); k[0] += 1
for _ in range(4 * NT): step()
t_pipe = timed(step, 48, warm=0, join=lambda: pipe.join_dev(st))
pipe.close()
pipe = ForegroundPipeline(5, nstreams=S)
d_mask = torch.empty((S, h, w), dtype=torch.uint8, device="cuda")
k = [0]
def step_m():
    pipe.process_dev(frames[k[0] % NT].data_ptr(), w, h, d_mask.data_ptr(), None, None, stream=st); k[0] += 1
for _ in range(4 * NT): step_m()
t_pipe_m = timed(step_m, 48, warm=0, join=lambda: pipe.join_dev(st))
print(json.dumps(dict(probe="pipeline packed", streams=S, us_per_step=t_pipe, mpixel_s=S*w*h/t_pipe, us_with_byte_mask=t_pipe_m,
                      components_stream0=(len(pipe.components(0)) if not os.environ.get('BGSB_PIPE_DBG') == '1' else -1))), flush=True)
# byte chain (round-1 form)
p = tb.MixtureOfGaussianV2BGS(nstreams=S); cc = blobs.ConnectedComponents(w, h, max_images=S)
fg = torch.empty((S, h, w), dtype=torch.uint8, device="cuda"); clean = torch.empty_like(fg)
def step_b():
    f = frames[k[0] % NT]; k[0] += 1
    p.process_dev(f.data_ptr(), w, h, fg.data_ptr(), None, stream=st)
    blobs.morph_dev(fg.data_ptr(), w, h, S, [("erode", 1), ("dilate", 1)], clean.data_ptr(), stream=st)
    cc.label_batch_dev(clean.data_ptr(), w, h, S, True, None, stream=st)
k = [0]
for _ in range(4 * NT): step_b()
t_b = timed(step_b, 48, warm=0)
def mog_only():
    f = frames[k[0] % NT]; k[0] += 1
    p.process_dev(f.data_ptr(), w, h, fg.data_ptr(), None, stream=st)
t_m = timed(mog_only, 48, warm=0)
t_o = timed(lambda: blobs.morph_dev(fg.data_ptr(), w, h, S, [("erode", 1), ("dilate", 1)], clean.data_ptr(), stream=st), 48)
t_c = timed(lambda: cc.label_batch_dev(clean.data_ptr(), w, h, S, True, None, stream=st), 48)
lab = torch.empty((S, h, w), dtype=torch.int32, device="cuda")
t_cl = timed(lambda: cc.label_batch_dev(clean.data_ptr(), w, h, S, True, lab.data_ptr(), stream=st), 48)
print(json.dumps(dict(probe="byte chain", streams=S, us_per_step=t_b, mog2_us=t_m, morph_us=t_o, ccl_table_us=t_c, ccl_labels_us=t_cl,
                      ccl_labels_gbs_5B=S*w*h*5/t_cl/1e3)), flush=True)
# labeller alone, single image
c1 = blobs.ConnectedComponents(w, h)
rng = np.random.default_rng(0)
for name, m in (("12 blobs", clean[0].cpu().numpy()), ("salt 0.2%", (rng.random((h, w)) < 0.002).astype(np.uint8) * 255)):
    d = torch.from_numpy(m).cuda(); l1 = torch.empty((h, w), dtype=torch.int32, device="cuda")
    a = timed(lambda: c1.label_dev(d.data_ptr(), w, h, True, None, stream=st), 100)
    b = timed(lambda: c1.label_dev(d.data_ptr(), w, h, True, l1.data_ptr(), stream=st), 100)
    print(json.dumps(dict(probe="ccl single 1080p", mask=name, components=len(c1.components()), us_table=a, us_labels=b)), flush=True)
