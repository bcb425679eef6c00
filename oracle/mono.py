"""oracle/mono.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Row f2 of SURVEY.md section 8(f): MonoVisionImuFrontend, restated from
  * MonoVisionImuFrontend::processFirstFrame   src/frontend/MonoVisionImuFrontend.cpp:194-221
  * MonoVisionImuFrontend::processFrame        src/frontend/MonoVisionImuFrontend.cpp:223-338
  * MonoVisionImuFrontend::getSmartMonoMeasurements  :343-371
  * Camera::Camera / Camera::undistortKeypoints      src/frontend/Camera.cpp:29-47, :110-133
with the classes of oracle/frontend.py for everything the two front-ends share (FeatureDetector, Tracker,
shouldBeKeyframe, outlierRejectionMono).  Differences from the stereo front-end that matter for parity:
no rectification rotation in the bearing vectors (featureDetection / featureTracking are called without R),
the tracking statuses are reset on EVERY frame (mono INVALID, stereo DISABLED: :266-267), there is no
"all tracks lost" shortcut, and the keyframe's keypoints are undistorted with P = K, R = I and checked against
the maps of that same transformation (keypoints_undistorted_).
"""
from __future__ import annotations

from typing import List, Optional

import cv2
import numpy as np

from kimera_vio_b200.params import CameraParams, FrontendParams
from . import frontend as ofe
from . import ransac as rs


class MonoCamera:
    """Camera (Camera.cpp:29-47): UndistorterRectifier(P = K, cam_params, R = I)."""

    def __init__(self, cam: CameraParams):
        self.left = cam
        self.W, self.H = cam.width, cam.height
        self.R1 = np.eye(3)
        self.P1 = np.hstack([cam.K, np.zeros((3, 1))])
        init = cv2.fisheye.initUndistortRectifyMap if cam.distortion_model == "equidistant" else cv2.initUndistortRectifyMap
        self.map_lx, self.map_ly = init(cam.K, cam.D, np.eye(3, dtype=np.float32), cam.K, (self.W, self.H), cv2.CV_32FC1)

    def undistort_keypoints(self, kps, pixel_tol: float = 2.0):
        """Camera::undistortKeypoints: undistortPoints(K, D, R = I, P = K) + checkUndistortedRectifiedLeftKeypoints."""
        return ofe.undistort_rectify_left_keypoints(kps, self, pixel_tol)


class MonoFrontend:
    def __init__(self, p: FrontendParams, cam: CameraParams, rnd_libstdcxx: str = "lemire"):
        self.p, self.cam = p, MonoCamera(cam)
        self.detector = ofe.FeatureDetector(p)
        self.tracker = ofe.Tracker(p, self.cam, rnd_libstdcxx)
        self.frame_count = 0
        self.keyframe_count = 0
        self.km1: Optional[ofe.Frame] = None
        self.lkf: Optional[ofe.Frame] = None
        self.keyframe_R_ref = np.eye(3)
        self.mono_status, self.stereo_status = ofe.INVALID, ofe.DISABLED
        self.lkf_T_k_mono = np.hstack([np.eye(3), np.zeros((3, 1))])
        self._last_disparity = 0.0

    # VisionImuFrontend::shouldBeKeyframe is shared: reuse the stereo oracle's method on this object
    should_be_keyframe = ofe.StereoFrontend.should_be_keyframe

    def _undistort(self, f: ofe.Frame):
        f.keypoints_undistorted = self.cam.undistort_keypoints(f.keypoints)

    def spin(self, frame: ofe.Frame, keyframe_R_cur: np.ndarray):
        if self.frame_count == 0:
            k = frame.clone()
            k.is_keyframe = True
            self.detector.feature_detection(k, None)
            self._undistort(k)
            self.km1 = self.lkf = k
            self.frame_count += 1
            return k, True, []
        p = self.p
        R = np.asarray(keyframe_R_cur, np.float64)
        k = frame.clone()
        ref_R_cur = rs.matmul3(self.keyframe_R_ref.T.copy(), R)
        self._track(self.km1, k, ref_R_cur)
        self.mono_status, self.stereo_status = ofe.INVALID, ofe.DISABLED            # :266-267, every frame
        smart: List = []
        if self.should_be_keyframe(k, self.lkf):
            self.keyframe_count += 1
            if p.use_ransac:
                given_rot = not ofe.rot_equals_identity(R)
                st, pose, _ = self.tracker.outlier_rejection_2d2d(self.lkf, k, R if (p.ransac_use_2point_mono and given_rot) else None)
                self.mono_status = st
                if st == ofe.VALID:
                    self.lkf_T_k_mono = pose
            else:
                self.mono_status = ofe.DISABLED
            k.is_keyframe = True
            self.detector.feature_detection(k, None)
            self._undistort(k)
            self.lkf = k
            for lmk, (st_, (ux, uy)) in zip(k.landmarks, k.keypoints_undistorted):
                if lmk != -1:
                    smart.append((lmk, float(ux), float("nan"), float(uy)))
        else:
            k.is_keyframe = False
        self.keyframe_R_ref = np.eye(3) if k.is_keyframe else R
        self.km1 = k
        self.frame_count += 1
        return k, k.is_keyframe, smart

    def _track(self, ref: ofe.Frame, cur: ofe.Frame, ref_R_cur):
        """Tracker::featureTracking without R (bearing vectors in the camera frame, not the rectified one)."""
        rig_R1 = self.tracker.rig.R1
        self.tracker.rig.R1 = None
        try:
            self.tracker.feature_tracking(ref, cur, ref_R_cur)
        finally:
            self.tracker.rig.R1 = rig_R1
