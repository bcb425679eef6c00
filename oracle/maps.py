"""oracle/maps.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU twin of kimera_vio_b200/csrc/common.cuh: rect_map_at (radial-tangential model) -- the CV_32FC1 value of
cv::initUndistortRectifyMap (called at src/frontend/UndistorterRectifier.cpp:246-258) at one pixel, with the operation
order and the three fused multiply-adds the CUDA code uses (exact FMA emulated with rationals).  Pinned against cv2 on
the reference's shipped rigs by tests/test_oracle_maps.py, including the places where the fusions decide the last bit
(zero crossings of the map of a zero-distortion camera, the f32 tie in row 15 of params/uHumans2).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from .fisheye import _inv3, _rp


def _fma(a, b, c) -> float:
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def radtan_map_at(K, D, R, P, u: int, v: int):
    K = np.asarray(K, np.float64)
    k1, k2, p1, p2 = (float(t) for t in np.asarray(D, np.float64).reshape(-1)[:4])
    iR = _inv3(_rp(P, R))
    ud, vd = float(u), float(v)
    X = _fma(ud, iR[0], _fma(vd, iR[1], iR[2]))
    Y = _fma(ud, iR[3], _fma(vd, iR[4], iR[5]))
    Wd = _fma(ud, iR[6], _fma(vd, iR[7], iR[8]))
    w = 1.0 / Wd
    x, y = X * w, Y * w
    x2, y2 = x * x, y * y
    r2, _2xy = x2 + y2, 2 * x * y
    kr = 1 + ((0.0 * r2 + k2) * r2 + k1) * r2
    xd = (x * kr + p1 * _2xy) + p2 * (r2 + 2 * x2)
    yd = (y * kr + p1 * (r2 + 2 * y2)) + p2 * _2xy
    return np.float32(_fma(K[0, 0], xd, K[0, 2])), np.float32(_fma(K[1, 1], yd, K[1, 2]))
