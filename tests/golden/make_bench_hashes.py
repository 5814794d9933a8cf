#!/usr/bin/env python
"""Oracle hashes that bench.py checks its own timed code paths against: the benchmark's kernels
run the first frames of the very streams they are then timed on, and the outputs must hash to these values.

    python tests/golden/make_bench_hashes.py        # writes tests/golden/bench_hashes.json (C oracle, ~1 minute)

The C oracle (oracle/c/bgs_oracle.c) is pinned bit-exactly to OpenCV 4.13 by tests/test_oracle_pin.py; the synthetic
video is the integer generator of SURVEY 8(d) (oracle twin: orc_synth_frame).
"""
import hashlib
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import restate as orc            # noqa: E402

W, H, SEED0 = 1920, 1080, 1234
MOG2_FRAMES, MOG2_SEEDS = 6, 8               # bench.py: rank r times stream seed SEED0 + r
PIPE_FRAMES, PIPE_STREAMS = 4, 64            # bench.py config 4: stream s uses seed SEED0 + s


def table_bytes(n, stats, ext):
    """(x, y, w, h, area, first_index, external) per component, int32 -- what bench.py builds from the C-ABI table."""
    rows = np.zeros((n, 7), np.int32)
    for i in range(n):
        xmin, ymin, xmax, ymax, area, first = (int(v) for v in stats[i])
        rows[i] = (xmin, ymin, xmax - xmin + 1, ymax - ymin + 1, area, first, int(ext[i]))
    return np.int32(n).tobytes() + rows.tobytes()


def mog2_stream(seed):
    o = orc.MixtureOfGaussianV2BGS()
    out = []
    for t in range(MOG2_FRAMES):
        fg, bg = o.process(orc.synth_frame(W, H, t, seed))
        out.append(hashlib.sha256(fg.tobytes() + bg.tobytes()).hexdigest())
    return out


def pipe_stream(seed):
    o = orc.MixtureOfGaussianV2BGS()
    h = hashlib.sha256()
    for t in range(PIPE_FRAMES):
        fg, _ = o.process(orc.synth_frame(W, H, t, seed))
        clean = orc.morph(orc.morph(fg, "erode", 1), "dilate", 1)
        n, _, stats, ext = orc.ccl8(clean, True)
        h.update(table_bytes(n, stats, ext))
    return h.hexdigest()


def main():
    orc.build()
    with ThreadPoolExecutor(8) as ex:
        mog2 = list(ex.map(mog2_stream, [SEED0 + r for r in range(MOG2_SEEDS)]))
        pipe = list(ex.map(pipe_stream, [SEED0 + s for s in range(PIPE_STREAMS)]))
    out = {"geometry": [W, H], "seed0": SEED0,
           "mog2": {"frames": MOG2_FRAMES, "what": "sha256(mask bytes + background bytes) per frame, MixtureOfGaussianV2BGS defaults",
                    "by_seed": {str(SEED0 + r): mog2[r] for r in range(MOG2_SEEDS)}},
           "pipeline": {"frames": PIPE_FRAMES, "what": "sha256 over frames of int32 n + n x (x,y,w,h,area,first_index,external): "
                                                      "MOG2 -> OPEN 3x3 -> 8-connected components, zero_border = 1",
                        "by_seed": {str(SEED0 + s): pipe[s] for s in range(PIPE_STREAMS)}}}
    with open(os.path.join(ROOT, "tests", "golden", "bench_hashes.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote bench_hashes.json")


if __name__ == "__main__":
    main()
