"""TEST INFRASTRUCTURE ONLY - restatement of CvBlobDetectorCC::DetectNewBlob.

PARITY UNPINNED for the list logic: OpenCV's `legacy` module (legacy/src/enteringblobdetection.cpp,
OpenCV 2.4.x -- the reference pulls it in through `opencv2/legacy/blobtrack.hpp`,
ustc_src/ustc_bgs.h:53, and creates the detector at ustc_src/trackingMain.cpp:56,626) is neither
vendored in /root/reference nor present in the cv2 4.13 wheel (removed upstream in 3.0), and the
reference holds no test or golden vector for it.  This module restates the published algorithm
(SURVEY.md Appendix A.6).  The image-processing steps ARE executed by the real OpenCV
(cv2.threshold, cv2.findContours(RETR_EXTERNAL), cv2.boundingRect, cv2.moments), so steps 1, 2 and
4 are pinned against the library; steps 3 and 5-7 (cvSeqPartition clustering, filtering, sorting,
track building) are a restatement only.

Arithmetic follows the C types of the original: CvBlob fields and the CompareContour /
trajectory-fit temporaries are fp32 (numpy.float32), moments are fp64.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

f32 = np.float32
SEQ_NUM = 1000


class Blob:
    __slots__ = ("x", "y", "w", "h", "id")

    def __init__(self, x, y, w, h, id=0):
        self.x, self.y, self.w, self.h, self.id = f32(x), f32(y), f32(w), f32(h), id

    def tuple(self):
        return (float(self.x), float(self.y), float(self.w), float(self.h))


def _rects_close(ra, rb):
    """CompareContour."""
    pax = f32(ra[0]) + f32(ra[2]) * f32(0.5); pay = f32(ra[1]) + f32(ra[3]) * f32(0.5)
    pbx = f32(rb[0]) + f32(rb[2]) * f32(0.5); pby = f32(rb[1]) + f32(rb[3]) * f32(0.5)
    w = f32(ra[2] + rb[2]) * f32(0.5); h = f32(ra[3] + rb[3]) * f32(0.5)
    dx = f32(abs(float(pax - pbx)) - float(w))
    dy = f32(abs(float(pay - pby)) - float(h))
    ht = f32(max(ra[3], rb[3])) * f32(0.3)
    return bool(dx < f32(0) and dy < ht)


def _partition(rects):
    """cvSeqPartition: classes numbered by first appearance."""
    n = len(rects)
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            a = parent[a]
        return a

    for i in range(n):
        for j in range(n):
            if i != j and _rects_close(rects[i], rects[j]):
                a, b = find(i), find(j)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    ids, cls = {}, []
    for i in range(n):
        r = find(i)
        if r not in ids:
            ids[r] = len(ids)
        cls.append(ids[r])
    return len(ids), cls


def _rx(b): return f32(0.5) * b.w
def _ry(b): return f32(0.5) * b.h


class CvBlobDetectorCC:
    def __init__(self, zero_border=True, latency=10):
        self.HMin, self.WMin, self.MinDistToBorder = f32(0.02), f32(0.01), f32(1.1)
        self.Clastering = 1
        self.SEQ_SIZE = latency
        self.zero_border = zero_border
        self.lists = [[] for _ in range(latency)]
        self.tracks = []      # each: {"size": int, "blobs": [Blob|None]*SEQ_SIZE}

    # steps 1-4
    def _frame_blobs(self, fg):
        H, W = fg.shape
        ib = cv2.threshold(fg, 128, 255, cv2.THRESH_BINARY)[1]
        if self.zero_border:      # OpenCV <= 3.1 cvFindContours clears the outer frame
            ib = ib.copy()
            ib[0, :] = 0; ib[-1, :] = 0; ib[:, 0] = 0; ib[:, -1] = 0
        contours, _ = cv2.findContours(ib, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        rects = [tuple(int(v) for v in cv2.boundingRect(c)) for c in contours]
        if self.Clastering:
            ncls, cls = _partition(rects)
            q = [None] * ncls
            for r, k in zip(rects, cls):
                if q[k] is None:
                    q[k] = r
                else:
                    R = q[k]
                    x0, y0 = min(R[0], r[0]), min(R[1], r[1])
                    x1, y1 = max(R[0] + R[2], r[0] + r[2]), max(R[1] + R[3], r[1] + r[3])
                    q[k] = (x0, y0, x1 - x0, y1 - y0)
        else:
            q = [r for r in rects if not (f32(r[3]) < f32(H) * self.HMin or f32(r[2]) < f32(W) * self.WMin)]
        blobs = []
        for (x, y, w, h) in q:
            if h < 1 or w < 1:
                X = Y = XX = YY = 0.0
            else:
                m = cv2.moments(fg[y:y + h, x:x + w], False)
                M00 = m["m00"]
                if M00 <= 0:
                    continue
                X = m["m10"] / M00; Y = m["m01"] / M00
                XX = m["m20"] / M00 - X * X; YY = m["m02"] / M00 - Y * Y
            sx = math.sqrt(XX) if XX >= 0 else float("nan")
            sy = math.sqrt(YY) if YY >= 0 else float("nan")
            blobs.append(Blob(f32(x) + f32(X), f32(y) + f32(Y), f32(4 * sx), f32(4 * sy)))
        return blobs

    def DetectNewBlob(self, fg, old_blobs=()):
        """-> (result, new_blob|None); self.lists[0] holds this frame's sorted top-10."""
        H, W = fg.shape
        S = self.SEQ_SIZE
        self.lists = [[]] + self.lists[:S - 1]
        blobs = self._frame_blobs(fg)
        # delete small and intersected
        for i in range(len(blobs), 0, -1):
            B = blobs[i - 1]
            if B.h < f32(H) * self.HMin or B.w < f32(W) * self.WMin:
                del blobs[i - 1]
                continue
            for O in reversed(list(old_blobs)):
                if abs(float(O.x - B.x)) < float(_rx(O) + _rx(B)) and abs(float(O.y - B.y)) < float(_ry(O) + _ry(B)):
                    del blobs[i - 1]
                    break
        # insertion sort by area, descending; later blob wins ties
        for i in range(1, len(blobs)):
            j = i
            while j > 0:
                if blobs[j].w * blobs[j].h < blobs[j - 1].w * blobs[j - 1].h:
                    break
                blobs[j], blobs[j - 1] = blobs[j - 1], blobs[j]
                j -= 1
        self.lists[0] = blobs[:10]
        # shift tracks
        for t in self.tracks:
            t["blobs"] = [None] + t["blobs"][:S - 1]
            if t["size"] == S:
                t["size"] -= 1
        ntracks = len(self.tracks)
        for B in reversed(self.lists[0]):
            assigned = 0
            for j in range(ntracks):
                t = self.tracks[j]
                last = t["blobs"][1] if t["size"] > 0 else None
                if last is None:
                    continue
                dx, dy = abs(float(last.x - B.x)), abs(float(last.y - B.y))
                if dx > float(f32(2) * last.w) or dy > float(f32(2) * last.h):
                    continue
                assigned += 1
                if t["blobs"][0] is None:
                    t["blobs"][0] = B
                    t["size"] += 1
                elif len(self.tracks) < SEQ_NUM:
                    d = {"size": t["size"], "blobs": list(t["blobs"])}
                    d["blobs"][0] = B
                    self.tracks.append(d)
            if assigned == 0 and len(self.tracks) < SEQ_NUM:
                self.tracks.append({"size": 1, "blobs": [B] + [None] * (S - 1)})
        best, best_err = -1, -1.0
        for i, t in enumerate(self.tracks):
            if t["size"] != S or t["blobs"][0] is None:
                continue
            B = t["blobs"][0]
            good = True
            for O in old_blobs:
                if abs(float(O.x - B.x)) < float(_rx(O) + _rx(B)) and abs(float(O.y - B.y)) < float(_ry(O) + _ry(B)):
                    good = False
            if good:
                dx = min(B.x, f32(W) - B.x) / _rx(B)
                dy = min(B.y, f32(H) - B.y) / _ry(B)
                if dx < self.MinDistToBorder or dy < self.MinDistToBorder:
                    good = False
            if good:
                N = t["size"]
                s0 = s1 = j0 = j1 = f32(0)
                for j in range(N):
                    x, y = t["blobs"][j].x, t["blobs"][j].y
                    s0 = f32(s0 + x); j0 = f32(j0 + f32(j) * x)
                    s1 = f32(s1 + y); j1 = f32(j1 + f32(j) * y)
                a0 = f32(f32(6) * f32(f32(f32(1 - N) * s0) + f32(f32(2) * j0))) / f32(N * (N * N - 1))
                b0 = f32(f32(-2) * f32(f32(f32(1 - 2 * N) * s0) + f32(f32(3) * j0))) / f32(N * (N + 1))
                a1 = f32(f32(6) * f32(f32(f32(1 - N) * s1) + f32(f32(2) * j1))) / f32(N * (N * N - 1))
                b1 = f32(f32(-2) * f32(f32(f32(1 - 2 * N) * s1) + f32(f32(3) * j1))) / f32(N * (N + 1))
                err = 0.0
                for j in range(N):
                    ex = f32(f32(f32(a0 * f32(j)) + b0) - t["blobs"][j].x)
                    ey = f32(f32(f32(a1 * f32(j)) + b1) - t["blobs"][j].y)
                    err += float(ex) ** 2 + float(ey) ** 2
                err = math.sqrt(err / N)
                if err > W * 0.01 or abs(float(a0)) > W * 0.1 or abs(float(a1)) > H * 0.1:
                    good = False
                if good and (best_err == -1 or best_err > err):
                    best, best_err = i, err
        result, new_blob = 0, None
        if best >= 0:
            t = self.tracks[best]
            new_blob = t["blobs"][0]
            t["blobs"][0] = None
            t["size"] -= 1
            result = 1
        i = len(self.tracks) - 1
        while i >= 0:
            if self.tracks[i]["blobs"][0] is None:
                self.tracks[i] = self.tracks[-1]
                self.tracks.pop()
            i -= 1
        return result, new_blob
