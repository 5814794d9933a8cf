#!/usr/bin/env python
"""Golden hashes for the OpenCV 2.4 BGR2GRAY constants (`grayVariant` 1):
    gray = (1868*B + 9617*G + 4899*R + 8192) >> 14
cv2 4.13 cannot produce them, so the plugins that only need integer arithmetic are restated here in pure numpy (no C
oracle, no cv2) on the committed clips:
    FrameDifferenceBGS        fg = gray24(|cur - prev|) > 15                       (FrameDifferenceBGS.cpp:45-51)
    StaticFrameDifferenceBGS  fg = gray24(|cur - first|) > 15, bg = first frame   (StaticFrameDifferenceBGS.cpp:29-57)
    python tests/golden/make_golden_gray24.py   -> tests/golden/golden_gray24.json
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def gray24(img):
    b, g, r = (img[..., c].astype(np.int64) for c in range(3))
    return ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8)


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def fd_masks(clip, thr=15):
    out = []
    for t in range(1, len(clip)):
        d = np.abs(clip[t].astype(np.int16) - clip[t - 1].astype(np.int16)).astype(np.uint8)
        out.append(((gray24(d) > thr) * 255).astype(np.uint8))
    return out


def sfd_masks(clip, thr=15):
    return [((gray24(np.abs(f.astype(np.int16) - clip[0].astype(np.int16)).astype(np.uint8)) > thr) * 255).astype(np.uint8) for f in clip]


def tie_frames():
    """Two frames whose difference image holds every colour (b < 144, g < 32, r < 56: all colours whose gray value can be
    15 or 16) on which the 2.4 and the 4.x constants put the gray value on different sides of the plugins' threshold 15 --
    the only inputs where `grayVariant` is observable."""
    b, g, r = np.meshgrid(np.arange(144), np.arange(32), np.arange(56), indexing="ij")
    g24 = (1868 * b + 9617 * g + 4899 * r + 8192) >> 14
    g4x = (3735 * b + 19235 * g + 9798 * r + 16384) >> 15
    sel = (g24 > 15) != (g4x > 15)
    cols = np.stack([b[sel], g[sel], r[sel]], -1).astype(np.uint8)
    n = len(cols)
    w = 64
    h = (n + w - 1) // w
    cur = np.zeros((h * w, 3), np.uint8)
    cur[:n] = cols
    return np.stack([np.zeros((h, w, 3), np.uint8), cur.reshape(h, w, 3)]), n


def main():
    z = np.load(os.path.join(HERE, "clips.npz"))
    out = {}
    ties, n = tie_frames()
    m24, m4x = fd_masks(ties)[0], fd_masks_4x(ties)[0]
    assert n > 0 and int((m24 != m4x).sum()) == n
    out["gray_ties"] = {"colours": n, "FrameDifferenceBGS:grayVariant=1": sha([m24]), "FrameDifferenceBGS:grayVariant=0": sha([m4x])}
    for name in z.files:
        clip = z[name]
        out[name] = {"FrameDifferenceBGS:grayVariant=1": sha(fd_masks(clip)),
                     "StaticFrameDifferenceBGS:grayVariant=1": sha(sfd_masks(clip)),
                     "differs_from_4x_constants_on_px": int(sum((a != b).sum() for a, b in zip(fd_masks(clip), fd_masks_4x(clip))))}
    with open(os.path.join(HERE, "golden_gray24.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


def fd_masks_4x(clip, thr=15):
    out = []
    for t in range(1, len(clip)):
        d = np.abs(clip[t].astype(np.int16) - clip[t - 1].astype(np.int16)).astype(np.int64)
        g = (3735 * d[..., 0] + 19235 * d[..., 1] + 9798 * d[..., 2] + 16384) >> 15
        out.append(((g > thr) * 255).astype(np.uint8))
    return out


if __name__ == "__main__":
    main()
