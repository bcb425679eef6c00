"""oracle/fisheye.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Scalar f64 restatement of the two cv::fisheye functions the equidistant rig model goes through
  * cv::fisheye::initUndistortRectifyMap  (called at src/frontend/UndistorterRectifier.cpp:260-268)
  * cv::fisheye::undistortPoints          (called at src/frontend/UndistorterRectifier.cpp:49-56)
operation by operation, in the order the CUDA code of kimera_vio_b200/csrc/common.cuh (rect_map_at_fisheye,
undistort_point_fisheye) evaluates them.  OpenCV is a dependency of the reference, not part of /root/reference; the
algorithm restated here is the published one of OpenCV 4.13 (the wheel in this image), and it is pinned bit for bit
against cv2.fisheye itself by tests/test_oracle_fisheye.py on the reference's params/RealSenseIR rig.  The oracle proper
(oracle/rig.py, oracle/frontend.py) calls cv2.fisheye directly; this file exists so that the device formulas have a CPU
twin that runs without a GPU.
"""
from __future__ import annotations

import math

import numpy as np


def _rp(P, R):
    """P[:, :3] * R with the 3x3 product order (a0 b0 + a1 b1) + a2 b2."""
    PP = np.asarray(P, np.float64)[:3, :3]
    R = np.asarray(R, np.float64)
    out = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            out[i, j] = (PP[i, 0] * R[0, j] + PP[i, 1] * R[1, j]) + PP[i, 2] * R[2, j]
    return out


def _inv3(S):
    """3x3 inverse by cofactors (cv::Matx33d::inv takes this path whatever the decomposition flag)."""
    S = np.asarray(S, np.float64).reshape(9)
    d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6])
    d = 1.0 / d
    t = np.empty(9)
    t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d
    t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d
    return t


def init_undistort_rectify_map(K, D, R, P, size):
    """CV_32FC1 maps.  The source coordinates of a row are three running f64 sums (+= iR(.,0) per column): column j is
    reached by j sequential additions, vectorised here over the rows only."""
    W, H = size
    K, D = np.asarray(K, np.float64), np.asarray(D, np.float64).reshape(-1)
    iR = _inv3(_rp(P, R))
    f0, f1, c0, c1 = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    mx, my = np.empty((H, W), np.float32), np.empty((H, W), np.float32)
    i = np.arange(H, dtype=np.float64)
    _x, _y, _w = i * iR[1] + iR[2], i * iR[4] + iR[5], i * iR[7] + iR[8]
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(W):
            x, y = _x / _w, _y / _w
            r = np.sqrt(x * x + y * y)
            th = np.arctan(r)
            t2 = th * th
            t4 = t2 * t2
            t6 = t4 * t2
            t8 = t4 * t4
            thd = th * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8)
            sc = np.where(r == 0, 1.0, thd / np.where(r == 0, 1.0, r))
            u, v = f0 * x * sc + c0, f1 * y * sc + c1
            neg = _w <= 0
            u = np.where(neg, np.where(_x > 0, -np.inf, np.inf), u)
            v = np.where(neg, np.where(_y > 0, -np.inf, np.inf), v)
            mx[:, j], my[:, j] = u.astype(np.float32), v.astype(np.float32)
            _x, _y, _w = _x + iR[0], _y + iR[3], _w + iR[6]
    return mx, my


def undistort_points(pts, K, D, R=None, P=None):
    """Default criteria (COUNT + EPS, 10, 1e-8): Newton on theta; (-1e6, -1e6) when it does not converge or flips sign."""
    K, D = np.asarray(K, np.float64), np.asarray(D, np.float64).reshape(-1)
    f0, f1, c0, c1 = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    RR = np.eye(3) if R is None else np.asarray(R, np.float64)
    if P is not None:
        RR = _rp(P, RR)
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    out = np.empty_like(pts)
    for i, (u, v) in enumerate(pts.astype(np.float64)):
        pw0, pw1 = (u - c0) / f0, (v - c1) / f1
        thd = math.sqrt(pw0 * pw0 + pw1 * pw1)
        thd = min(max(-math.pi / 2.0, thd), math.pi / 2.0)
        conv, th, scale = False, thd, 0.0
        if abs(thd) > 1e-8:
            for _ in range(10):
                t2 = th * th
                t4 = t2 * t2
                t6 = t4 * t2
                t8 = t6 * t2
                a, b, c, d = D[0] * t2, D[1] * t4, D[2] * t6, D[3] * t8
                fix = (th * (1 + a + b + c + d) - thd) / (1 + 3 * a + 5 * b + 7 * c + 9 * d)
                th = th - fix
                if abs(fix) < 1e-8:
                    conv = True
                    break
            scale = math.tan(th) / thd
        else:
            conv = True
        flipped = (thd < 0 and th > 0) or (thd > 0 and th < 0)
        if conv and not flipped:
            x, y = pw0 * scale, pw1 * scale
            pr = [(RR[r, 0] * x + RR[r, 1] * y) + RR[r, 2] for r in range(3)]
            out[i] = (np.float32(pr[0] / pr[2]), np.float32(pr[1] / pr[2]))
        else:
            out[i] = (-1000000.0, -1000000.0)
    return out
