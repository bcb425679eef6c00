"""oracle/rig.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Row a1 of SURVEY.md section 8(a): the stereo rig, restating
  * StereoCamera::StereoCamera                      src/frontend/StereoCamera.cpp:34-94
  * StereoCamera::computeRectificationParameters    src/frontend/StereoCamera.cpp:292-379
  * UndistorterRectifier::initUndistortRectifyMaps  src/frontend/UndistorterRectifier.cpp:230-292
with the same OpenCV calls (cv2.stereoRectify, cv2.initUndistortRectifyMap; their cv2.fisheye counterparts for the
equidistant distortion model).
Pinned by tests/testStereoMatcher.cpp:148 (baseline 0.110078 on the Euroc rig).
"""
from __future__ import annotations

import cv2
import numpy as np

from kimera_vio_b200.params import CameraParams


class StereoRig:
    def __init__(self, left: CameraParams, right: CameraParams):
        self.left, self.right = left, right
        self.W, self.H = left.width, left.height
        size = (self.W, self.H)
        # camL_Pose_camR = body_Pose_camL.between(body_Pose_camR); OpenCV wants the inverse
        # (StereoCamera.cpp:313-322).
        camL_T_camR = np.linalg.inv(left.T_BS) @ right.T_BS
        inv = np.linalg.inv(camL_T_camR)
        R, T = inv[:3, :3].copy(), inv[:3, 3].copy()
        if left.distortion_model == "radtan":
            self.R1, self.R2, self.P1, self.P2, self.Q, self.roi1, self.roi2 = cv2.stereoRectify(
                left.K, left.D, right.K, right.D, size, R, T,
                flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)          # kAlpha = 0, StereoCamera.cpp:326
        elif left.distortion_model == "equidistant":
            # StereoCamera.cpp:350-373: cv::fisheye::stereoRectify(..., CALIB_ZERO_DISPARITY), no alpha, no ROIs
            self.R1, self.R2, self.P1, self.P2, self.Q = cv2.fisheye.stereoRectify(
                left.K, left.D, right.K, right.D, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY)
        else:
            raise NotImplementedError("Unknown DistortionModel")        # LOG(FATAL), StereoCamera.cpp:370-377
        # baseline = 1 / Q(3,2)  (StereoCamera.cpp:70-72)
        self.baseline = 1.0 / self.Q[3, 2]
        assert self.baseline > 0
        # Cal3_S2Stereo(fx, fy, skew, px, py, baseline) from P1 (StereoCamera.cpp:75-82)
        self.fx, self.fy = self.P1[0, 0], self.P1[1, 1]
        self.skew = self.P1[0, 1]
        self.cx, self.cy = self.P1[0, 2], self.P1[1, 2]
        # float32 maps, CV_32FC1 (UndistorterRectifier.cpp:238-258)
        # UndistorterRectifier.cpp:246-268: cv::initUndistortRectifyMap (RADTAN) / cv::fisheye:: (EQUIDISTANT)
        init = cv2.fisheye.initUndistortRectifyMap if left.distortion_model == "equidistant" else cv2.initUndistortRectifyMap
        self.map_lx, self.map_ly = init(left.K, left.D, self.R1, self.P1, size, cv2.CV_32FC1)
        self.map_rx, self.map_ry = init(right.K, right.D, self.R2, self.P2, size, cv2.CV_32FC1)

    # UndistorterRectifier::undistortRectifyImage  UndistorterRectifier.cpp:115-128
    def rectify_left(self, img: np.ndarray) -> np.ndarray:
        return cv2.remap(img, self.map_lx, self.map_ly, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)

    def rectify_right(self, img: np.ndarray) -> np.ndarray:
        return cv2.remap(img, self.map_rx, self.map_ry, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
