"""Which pixels leave the MOG2 fast path?  Classifies the pixels of frame NF+1 against the model exported after NF
frames (numpy restatement of the match test only; GPU box, measurement tooling)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracking_b200 as tb
from tracking_b200 import synth
W, H, NF = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 128
st = torch.cuda.current_stream().cuda_stream
d = torch.empty((NF + 1, H, W, 3), dtype=torch.uint8, device="cuda")
synth.frames_dev(d.data_ptr(), 1, NF + 1, W, H, stream=st)
fg = torch.empty((H, W), dtype=torch.uint8, device="cuda"); bg = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
p = tb.MixtureOfGaussianV2BGS()
for t in range(NF):
    p.process_dev(d[t].data_ptr(), W, H, fg.data_ptr(), bg.data_ptr(), stream=st)
torch.cuda.synchronize()
planes, nm = p.export_state()
x = d[NF].cpu().numpy().reshape(-1, 3).astype(np.float32)
npx = W * H
Tg = 9.0
fit = np.full(npx, -1, np.int32)
for m in range(5):
    w, v, b, g, r = planes[m * 5:(m + 1) * 5]
    d2 = (b - x[:, 0]) ** 2 + (g - x[:, 1]) ** 2 + (r - x[:, 2]) ** 2
    ok = (m < nm) & (d2 < Tg * v) & (fit < 0)
    fit[ok] = m
print("frames", NF, "mean modes", nm.mean(), "hist n", np.bincount(nm, minlength=6) / npx)
print("no fit             ", (fit < 0).mean())
for n in range(1, 6):
    for f in range(n):
        print("n=%d fit slot %d     " % (n, f), ((nm == n) & (fit == f)).mean())
for n in range(1, 6):
    print("n=%d no fit         " % n, ((nm == n) & (fit < 0)).mean())
# --- of the pixels that match slot 0: how many need three or more modes for the background image? ---
aT = np.float32(p.get("alpha")); a1 = np.float32(1) - aT; prune = np.float32(-float(aT) * 0.05)
w = [planes[m * 5].astype(np.float32) for m in range(5)]
wt = [a1 * w[m] + prune for m in range(5)]
wt[0] = wt[0] + aT
live = [(nm > m) for m in range(5)]
tw = sum(np.where(live[m], wt[m], 0) for m in range(5))
wn = [np.where(live[m], wt[m] / tw, 0) for m in range(5)]
fit0 = (fit == 0)
need3 = fit0 & (nm >= 3) & ((wn[0] + wn[1]) <= np.float32(0.9))
pruned = fit0 & np.any([live[m] & (wt[m] < -prune) for m in range(1, 5)], axis=0)
print("fit slot 0 and background image needs >= 3 modes:", need3.mean())
print("fit slot 0 and some weight pruned this frame     :", pruned.mean())
# --- how are the ineligible pixels spread over the 64-pixel state tiles (= warps of the T == 1 kernel)? ---
slowpx = ~fit0 | need3 | pruned
ntile = npx // 64
per_tile = slowpx[:ntile * 64].reshape(ntile, 64).sum(1)
print("ineligible (approx.):", slowpx.mean(), " tiles with none: %.3f" % (per_tile == 0).mean())
print("tiles by count 1-4 / 5-16 / 17-32 / 33-64: %.3f %.3f %.3f %.3f" % (
    ((per_tile >= 1) & (per_tile <= 4)).mean(), ((per_tile >= 5) & (per_tile <= 16)).mean(),
    ((per_tile >= 17) & (per_tile <= 32)).mean(), (per_tile >= 33).mean()))
print("share of ineligible pixels in tiles with > 16 of them: %.3f" % (per_tile[per_tile > 16].sum() / max(1, per_tile.sum())))
per_cta = slowpx[:(npx // 256) * 256].reshape(-1, 256).sum(1)
print("per 256-px CTA: mean %.1f, none %.3f, > 32: %.3f, > 64: %.3f" % (per_cta.mean(), (per_cta == 0).mean(), (per_cta > 32).mean(), (per_cta > 64).mean()))
