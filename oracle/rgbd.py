"""oracle/rgbd.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Row f2 (RGB-D, the stage-level part): CPU restatement of
  * DepthFrame::getDepthAtPoint     src/frontend/DepthFrame.cpp:39-73
  * DepthFrame::getDetectionMask    src/frontend/DepthFrame.cpp:75-98      (cv::inRange, made through cv2)
  * RgbdFrame::fillStereoFrame      src/frontend/RgbdFrame.cpp:52-115
  * RgbdCamera::distortKeypoints    src/frontend/RgbdCamera.cpp:81-85      (UndistorterRectifier::distortUnrectifyKeypoints)
with the reference's types: depth values, depth_to_meters, min/max depth and virtual_baseline are float32
(CameraParams.h:136-145), fx_b is double, the disparity and uR are float32 (RgbdFrame.cpp:66,96-97).
Pinned by the known answers of tests/testDepthFrame.cpp:129-173 (GetDepthAtPoint, float and uint16 images),
tests/testDepthFrame.cpp:58-96 (DetectionMask on tests/data/ForRgbd/depth_img_0.tiff) and tests/testRgbdFrame.cpp:84-172
(FillStereoFrame) in tests/test_oracle_rgbd.py.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import cv2
import numpy as np

from . import frontend as ofe

f32 = np.float32


def get_depth_at_point(depth_img: np.ndarray, depth: dict, point) -> float:
    """float32 result; NaN outside the image and below min_depth.  static_cast<int> truncates toward zero."""
    x, y = int(f32(point[0])), int(f32(point[1]))
    H, W = depth_img.shape[:2]
    if x < 0 or x >= W or y < 0 or y >= H:
        return f32(np.nan)
    if depth_img.dtype == np.float32:
        d = f32(depth_img[y, x])
    elif depth_img.dtype == np.uint16:
        d = f32(depth_img[y, x])
    else:
        raise TypeError("Invalid depth datatype")                    # LOG(FATAL), DepthFrame.cpp:60-62
    with np.errstate(invalid="ignore", over="ignore"):
        d = f32(d * f32(depth["depth_to_meters"]))
    if d < f32(depth["min_depth"]):
        return f32(np.nan)
    return d


def get_detection_mask(depth_img: np.ndarray, depth: dict) -> np.ndarray:
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        lo = f32(f32(depth["min_depth"]) * f32(1.0) / f32(depth["depth_to_meters"]))
        hi = f32(f32(depth["max_depth"]) * f32(1.0) / f32(depth["depth_to_meters"]))
    if depth_img.dtype == np.float32:
        return cv2.inRange(depth_img, float(lo), float(hi))
    if depth_img.dtype == np.uint16:
        # static_cast<uint16_t>(float): truncation (out-of-range values are undefined in C++; clamped here and in kvfe)
        def u16(v):
            return 0 if not v > 0 else (65535 if v >= 65535 else int(v))
        return cv2.inRange(depth_img, u16(lo), u16(hi))
    raise TypeError("Invalid depth datatype")


def fill_stereo_frame(depth_img: np.ndarray, cam, keypoints, left_rect, versors, map_x=None, map_y=None):
    """cam: CameraParams with the `depth` block.  keypoints: raw left keypoints; left_rect: [(status, (x, y))];
    versors: 3-vectors.  Returns (right_rect [(status, (x, y))], keypoints_depth, keypoints_3d, right_keypoints);
    right_keypoints need the camera's maps (map_x, map_y: CV_32FC1 of UndistorterRectifier(P = K, R = I))."""
    dp = cam.depth
    fx_b = float(cam.intrinsics[0]) * float(f32(dp["virtual_baseline"]))          # double * float -> double
    right, depths, p3d = [], [], []
    for i, (st, (lx, ly)) in enumerate(left_rect):
        if st != ofe.KP_VALID:
            right.append((st, (f32(0.0), f32(0.0))))
            depths.append(0.0)
            p3d.append(np.zeros(3))
            continue
        kd = get_depth_at_point(depth_img, dp, keypoints[i])
        if not math.isfinite(float(kd)):
            right.append((ofe.KP_NO_DEPTH, (f32(0.0), f32(0.0))))
            depths.append(0.0)
            p3d.append(np.zeros(3))
            continue
        with np.errstate(divide="ignore", over="ignore"):
            disparity = f32(np.float64(fx_b) / np.float64(kd))
            uR = f32(f32(lx) - disparity)
        if uR < f32(0.0):
            right.append((ofe.KP_NO_DEPTH, (f32(0.0), f32(0.0))))
            depths.append(0.0)
            p3d.append(np.zeros(3))
            continue
        v = np.asarray(versors[i], np.float64)
        right.append((ofe.KP_VALID, (uR, f32(ly))))
        depths.append(float(kd))
        p3d.append(v * float(kd) / v[2])
    right_kps: List[Tuple[float, float]] = []
    if map_x is not None:
        for st, (x, y) in right:                                    # distortUnrectifyKeypoints
            if st == ofe.KP_VALID:
                right_kps.append((map_x[ofe.c_round(y), ofe.c_round(x)], map_y[ofe.c_round(y), ofe.c_round(x)]))
            else:
                right_kps.append((f32(0.0), f32(0.0)))
    return right, depths, p3d, right_kps
