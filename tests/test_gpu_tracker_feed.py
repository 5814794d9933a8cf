"""SURVEY 8(f) N2 -- what CvBlobTracker takes from the mask per frame (ustc_src/trackingMain.cpp:70-78,166), served from
the GPU component table, against OpenCV itself: cv2.findContours(RETR_EXTERNAL) rectangles IN ITS ORDER and
cv2.sumElems of mask ROIs (the blob deleter's cvSum)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def masks_for_feed():
    rng = np.random.default_rng(21)
    h, w = 240, 352
    yy, xx = np.mgrid[0:h, 0:w]
    out = []
    for k in range(6):
        m = np.zeros((h, w), np.uint8)
        for _ in range(10):
            cy, cx, r = rng.integers(10, h - 10), rng.integers(10, w - 10), rng.integers(4, 40)
            d = (yy - cy) ** 2 + (xx - cx) ** 2
            if rng.random() < 0.5:
                m[(d < r * r) & (d >= (r * 0.6) ** 2)] = 255          # ring: whatever lies inside is NOT external
                m[d < (r * 0.2) ** 2] = 255
            else:
                m[d < r * r] = 255
        if k % 2:
            m[rng.random((h, w)) < 0.01] = 255                         # salt: many one-pixel contours
        if k == 4:
            m[0, :] = 255; m[:, 0] = 255                               # foreground on the frame
        out.append(m)
    out.append(np.zeros((h, w), np.uint8))
    return out


def test_tracker_feed_contour_rects_and_roi_sums_match_opencv():
    import cv2
    from tracking_b200 import blobs
    for m in masks_for_feed():
        h, w = m.shape
        cc = blobs.ConnectedComponents(w, h)
        n, _, comps = cc.label(m, zero_border=False, want_labels=False)      # cv2 >= 3.2 keeps the frame pixels
        feed = [(c["x"], c["y"], c["w"], c["h"]) for c in reversed(comps) if c["external"]]
        contours, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        assert feed == [tuple(cv2.boundingRect(c)) for c in contours]
        rois = feed[:40] + [(0, 0, w, h), (5, 7, 100, 50)]
        got = cc.rect_moments(rois)
        for (x, y, ww, hh), g in zip(rois, got):
            assert g[0] == int(cv2.sumElems(m[y:y + hh, x:x + ww])[0])
            mo = cv2.moments(m[y:y + hh, x:x + ww], False)
            assert (g[0], g[1], g[2], g[3], g[4], g[5]) == (int(mo["m00"]), int(mo["m10"]), int(mo["m01"]), int(mo["m20"]),
                                                           int(mo["m02"]), int(mo["m11"]))
        cc.close()
