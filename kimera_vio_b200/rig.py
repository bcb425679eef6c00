"""Stereo rig set-up (row a1 of SURVEY.md section 8(a)): one-off host work, not on the hot path.

StereoCamera::computeRectificationParameters (reference src/frontend/StereoCamera.cpp:292-379) calls
cv::stereoRectify(CALIB_ZERO_DISPARITY, alpha = 0); in a drop-in deployment the reference's own
StereoCamera constructor has already produced R1/R2/P1/P2/Q and hands them to kvfe_create() through
kvfe_rig.  This helper produces the same struct for the Python harness (tests, bench) with the
very same OpenCV call; nothing per-frame goes through OpenCV.
"""
from __future__ import annotations

import numpy as np

from .params import CameraParams
from . import lib as _lib


class StereoRigSetup:
    def __init__(self, left: CameraParams, right: CameraParams):
        import cv2  # set-up only
        self.left, self.right = left, right
        self.W, self.H = left.width, left.height
        camL_T_camR = np.linalg.inv(left.T_BS) @ right.T_BS
        inv = np.linalg.inv(camL_T_camR)
        R, T = inv[:3, :3].copy(), inv[:3, 3].copy()
        if left.distortion_model == "radtan":
            self.R1, self.R2, self.P1, self.P2, self.Q, _, _ = cv2.stereoRectify(
                left.K, left.D, right.K, right.D, (self.W, self.H), R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
        elif left.distortion_model == "equidistant":      # StereoCamera.cpp:350-373
            self.R1, self.R2, self.P1, self.P2, self.Q = cv2.fisheye.stereoRectify(
                left.K, left.D, right.K, right.D, (self.W, self.H), R, T, flags=cv2.CALIB_ZERO_DISPARITY)
        else:
            raise NotImplementedError("distortion model %r: only radtan and equidistant pinhole cameras" % left.distortion_model)
        self.baseline = 1.0 / self.Q[3, 2]
        self.fx, self.fy, self.cx, self.cy = self.P1[0, 0], self.P1[1, 1], self.P1[0, 2], self.P1[1, 2]

    def to_c(self) -> "_lib.Rig":
        return _lib.make_rig(self.left, self.right, self.R1, self.R2, self.P1, self.P2, self.baseline)


class MonoRigSetup:
    """Camera (reference src/frontend/Camera.cpp:29-47) as a kvfe_rig: UndistorterRectifier(P = K, cam_params, R = I)."""

    def __init__(self, cam: CameraParams):
        if cam.distortion_model not in ("radtan", "equidistant"):
            raise NotImplementedError("distortion model %r: only radtan and equidistant pinhole cameras" % cam.distortion_model)
        self.left = self.right = cam
        self.W, self.H = cam.width, cam.height
        self.R1 = self.R2 = np.eye(3)
        self.P1 = self.P2 = np.hstack([cam.K, np.zeros((3, 1))])
        self.baseline = 0.0
        self.fx, self.fy, self.cx, self.cy = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]

    def to_c(self) -> "_lib.Rig":
        return _lib.make_rig(self.left, self.right, self.R1, self.R2, self.P1, self.P2, self.baseline)


class RgbdRigSetup(MonoRigSetup):
    """RgbdCamera (reference src/frontend/RgbdCamera.cpp:79-101): the mono camera (P = K, R = I) plus the fake stereo
    camera of getFakeStereoCamera -- Cal3_S2Stereo(K, virtual_baseline) at the identity pose -- which the stereo outlier
    rejection of the RGB-D front-end runs on (RgbdVisionImuFrontend.cpp:325-336)."""

    def __init__(self, cam: CameraParams):
        super().__init__(cam)
        if not cam.depth:
            raise ValueError("camera parameters without the RGB-D block (virtual_baseline, ...)")
        self.baseline = float(cam.depth["virtual_baseline"])
