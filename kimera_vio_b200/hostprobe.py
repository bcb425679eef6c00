"""Host arithmetic probes a deployment needs to reproduce the reference bit for bit.

OpenCV's cv::Sobel row filter (inside cv::cornerMinEigenVal <- cv::goodFeaturesToTrack,
reference src/frontend/feature-detector/FeatureDetector.cpp:165-172) runs a fused-multiply-add SIMD
body and, depending on the build and the CPU it dispatches to, a NON-fused scalar tail over the last
columns of every row.  The response map therefore depends on the host the reference
runs on; `sobel_cpu_tail_start` measures where that tail starts with the OpenCV the caller links, and
the value goes into kvfe_config.sobel_cpu_tail_start (-1: no scalar tail).  The C++ shim does the same
probe once at start-up."""
from __future__ import annotations

import numpy as np


def sobel_cpu_tail_start(width: int) -> int:
    import cv2
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, width), dtype=np.uint8)
    dy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3, scale=1 / 3060.0)
    I = img.astype(np.float32)
    s = np.float32(1 / 3060.0)
    s2 = np.float32(2) * s
    xs = np.arange(width)
    xm, xp = np.abs(xs - 1), np.where(xs + 1 >= width, 2 * (width - 1) - (xs + 1), xs + 1)
    ys = np.arange(64)
    ym, yp = np.abs(ys - 1), np.where(ys + 1 >= 64, 2 * 63 - (ys + 1), ys + 1)
    t0 = s * I[:, xm]
    t1 = (I.astype(np.float64) * np.float64(s2) + t0.astype(np.float64)).astype(np.float32)
    t_fma = (I[:, xp].astype(np.float64) * np.float64(s) + t1.astype(np.float64)).astype(np.float32)
    t_nofma = (s * I[:, xm] + s2 * I) + s * I[:, xp]
    dy_fma = t_fma[yp] - t_fma[ym]
    dy_no = t_nofma[yp] - t_nofma[ym]
    col_fma_ok = np.all(dy_fma == dy, axis=0)
    col_no_ok = np.all(dy_no == dy, axis=0)
    # the scalar tail is the suffix of columns that only the non-FMA formula explains
    if col_fma_ok.all():
        return -1
    start = int(np.argmin(col_fma_ok))
    if not (col_no_ok[start:].all() and col_fma_ok[:start].all()):
        raise RuntimeError("unexpected cv2.Sobel arithmetic on this host")
    return start
