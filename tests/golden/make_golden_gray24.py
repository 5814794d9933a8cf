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


def main():
    z = np.load(os.path.join(HERE, "clips.npz"))
    out = {}
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
