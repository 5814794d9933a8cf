"""Synthetic camera streams of SURVEY.md 8(d): numpy twin of csrc/synth.cu (byte-identical).

frame(s,t)[y,x,c] = clamp_u8(B[y,x,c] + N(s,t,y,x,c)), overwritten by 12 moving rectangles;
integer-only so that the GPU generator, this module and the C oracle agree exactly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

SEED0 = 1234          # stream s uses SEED0 + s


def _mix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x7FEB352D)
    h ^= h >> np.uint32(15)
    h *= np.uint32(0x846CA68B)
    h ^= h >> np.uint32(16)
    return h


def frame(w, h, t, seed=SEED0):
    """One BGR frame (h, w, 3) uint8."""
    scale = 2 if w >= 3840 else 1
    x = np.arange(w, dtype=np.int64)[None, :, None]
    y = np.arange(h, dtype=np.int64)[:, None, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    B = 32 + (((x * 5 + y * 3 + 64 * c) >> 3) & 127) + 48 * (((x >> 6) ^ (y >> 6)) & 1)
    with np.errstate(over="ignore"):
        hsh = ((x.astype(np.uint32) * np.uint32(73856093)) ^ (y.astype(np.uint32) * np.uint32(19349663)) ^
               np.uint32((t * 83492791) & 0xFFFFFFFF) ^ (c.astype(np.uint32) * np.uint32(2654435761)) ^
               np.uint32(seed & 0xFFFFFFFF))
        N = (_mix32(hsh) % np.uint32(13)).astype(np.int64) - 6
    img = np.clip(B + N, 0, 255).astype(np.uint8)
    for r in range(12):
        rw, rh = 100 * scale, 80 * scale
        x0 = (100 + 150 * r + 17 * t) % (w - 120 * scale)
        y0 = (60 + 83 * r + 5 * t) % (h - 90 * scale)
        img[y0:y0 + rh, x0:x0 + rw] = ((40 * r) & 255, 255 - 20 * r, 128)
    return img


def frames_dev(d_ptr, nstreams, T, w, h, t0=0, seed0=SEED0, stream=0):
    """Fill a device buffer [nstreams][T][h][w][3] with synthetic frames (K-GEN)."""
    capi.check(capi.lib().bgsb_synth_frames_dev(C.c_void_p(d_ptr), nstreams, T, w, h, t0, seed0 & 0xFFFFFFFF,
                                                C.c_void_p(stream)))


def churn_frame(w, h, t, seed=SEED0):
    """One frame of the mode-churn video (numpy twin of synth_churn_kernel): five well-separated colours per pixel,
    redrawn every second frame, noise in [-2, 2] -- keeps all K = 5 mixture modes of every pixel live."""
    x = np.arange(w, dtype=np.int64)[None, :, None]
    y = np.arange(h, dtype=np.int64)[:, None, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    with np.errstate(over="ignore"):
        xy = (x.astype(np.uint32) * np.uint32(73856093)) ^ (y.astype(np.uint32) * np.uint32(19349663))
        hi = xy ^ np.uint32(((t >> 1) * 83492791) & 0xFFFFFFFF) ^ np.uint32(((seed & 0xFFFFFFFF) * 2246822519) & 0xFFFFFFFF)
        i = (_mix32(hi) % np.uint32(5)).astype(np.int64)                       # (h, w, 1)
        hsh = xy ^ np.uint32((t * 83492791) & 0xFFFFFFFF) ^ (c.astype(np.uint32) * np.uint32(2654435761)) ^ np.uint32(seed & 0xFFFFFFFF)
        N = (_mix32(hsh) % np.uint32(5)).astype(np.int64) - 2
    col = np.concatenate([20 + 50 * i, 230 - 45 * i, (90 + 110 * i) & 255], axis=2)
    return np.clip(col + N, 0, 255).astype(np.uint8)


def churn_frames_dev(d_ptr, nstreams, T, w, h, t0=0, seed0=SEED0, stream=0):
    capi.check(capi.lib().bgsb_synth_churn_frames_dev(C.c_void_p(d_ptr), nstreams, T, w, h, t0, seed0 & 0xFFFFFFFF,
                                                      C.c_void_p(stream)))
