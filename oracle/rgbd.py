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


# ----------------------------------------------------------------------------------------------------------------
# RgbdVisionImuFrontend (src/frontend/RgbdVisionImuFrontend.cpp:183-395) -- the caller of the two functions above, restated
# so that whole RGB-D sequences have an oracle (the frame-level RGB-D step of kvfe is not built yet; this is its checker).
# PnP (use_pnp_tracking, Tracker.cpp:1064-1288) is left out: it needs the back-end's landmark map.
# ----------------------------------------------------------------------------------------------------------------
class RgbdCamera:
    """RgbdCamera (RgbdCamera.cpp:79-101) = Camera (P = K, R = I, Camera.cpp:29-47) + the fake stereo calibration
    Cal3_S2Stereo(K, virtual_baseline) at the identity pose."""

    def __init__(self, cam):
        from .mono import MonoCamera
        m = MonoCamera(cam)
        self.left = self.right = cam
        self.W, self.H, self.R1, self.P1 = m.W, m.H, m.R1, m.P1
        self.map_lx, self.map_ly = m.map_lx, m.map_ly
        self.fx, self.fy, self.cx, self.cy = (float(v) for v in cam.intrinsics)
        self.baseline = float(f32(cam.depth["virtual_baseline"]))
        self._mono = m

    def undistort_keypoints(self, kps):
        return self._mono.undistort_keypoints(kps)


class _DepthMaskedDetector(ofe.FeatureDetector):
    """FeatureDetector::featureDetection with Frame::detection_mask_ set (FeatureDetector.cpp:185-203): the circles around
    the tracked keypoints are drawn into the caller's mask."""
    detection_mask = None

    def build_mask(self, frame):
        mask = self.detection_mask.copy()
        for kp, lmk in zip(frame.keypoints, frame.landmarks):
            if lmk != -1:
                c = (int(np.rint(f32(kp[0]))), int(np.rint(f32(kp[1]))))
                cv2.circle(mask, c, self.p.min_distance, 0, cv2.FILLED)
        return mask


class RgbdFrontend:
    def __init__(self, p, cam, rnd_libstdcxx: str = "lemire"):
        import numpy as _np
        from . import ransac as rs
        self._rs = rs
        self.p, self.cam_params, self.cam = p, cam, RgbdCamera(cam)
        self.detector = _DepthMaskedDetector(p)
        self.tracker = ofe.Tracker(p, self.cam, rnd_libstdcxx)
        self.frame_count = self.keyframe_count = 0
        self.km1 = self.lkf = None
        self.keyframe_R_ref = _np.eye(3)
        self.mono_status = self.stereo_status = ofe.INVALID
        self.lkf_T_k_mono = _np.hstack([_np.eye(3), _np.zeros((3, 1))])
        self.lkf_T_k_stereo = self.lkf_T_k_mono.copy()
        self._last_disparity = 0.0

    should_be_keyframe = ofe.StereoFrontend.should_be_keyframe

    def _stereo_frame(self, k, ts, img):
        return ofe.StereoFrame.make(k, ts, img, np.zeros((0, 0), np.uint8), self.cam)      # getStereoFrame: empty right image

    def _fill(self, sf, depth):
        """camera_->undistortKeypoints (when new keypoints exist) + RgbdFrame::fillStereoFrame."""
        lf = sf.left_frame
        if len(lf.keypoints) > len(sf.left_keypoints_rectified):
            sf.left_keypoints_rectified = self.cam.undistort_keypoints(lf.keypoints)
        r, d, p3, rk = fill_stereo_frame(depth, self.cam_params, lf.keypoints, sf.left_keypoints_rectified, lf.versors,
                                         self.cam.map_lx, self.cam.map_ly)
        sf.right_keypoints_rectified, sf.keypoints_depth, sf.keypoints_3d = r, d, p3
        sf.right_frame.keypoints = rk

    def _detect(self, sf, depth):
        self.detector.detection_mask = get_detection_mask(depth, self.cam_params.depth)
        self.detector.feature_detection(sf.left_frame, None)
        sf.left_keypoints_rectified = self.cam.undistort_keypoints(sf.left_frame.keypoints)

    def _track(self, ref, cur, ref_R_cur):
        R1 = self.tracker.rig.R1
        self.tracker.rig.R1 = None                    # featureTracking without R: versors in the camera frame
        try:
            self.tracker.feature_tracking(ref, cur, ref_R_cur)
        finally:
            self.tracker.rig.R1 = R1

    def smart_measurements(self, sf):
        """RgbdVisionImuFrontend::fillSmartStereoMeasurements (:368-395)."""
        out = []
        for i, l in enumerate(sf.left_frame.landmarks):
            if l == -1:
                continue
            uL, v = float(sf.left_keypoints_rectified[i][1][0]), float(sf.left_keypoints_rectified[i][1][1])
            st, (rx, _) = sf.right_keypoints_rectified[i]
            out.append((l, uL, float(rx) if st == ofe.KP_VALID else float("nan"), v))
        return out

    def spin(self, k: int, timestamp: int, img: np.ndarray, depth: np.ndarray, keyframe_R_cur: np.ndarray):
        """Returns (stereo frame, is_keyframe, smart measurements)."""
        rs, p = self._rs, self.p
        sf = self._stereo_frame(k, timestamp, img)
        if self.frame_count == 0:                                            # processFirstFrame :183-209
            sf.is_keyframe = sf.left_frame.is_keyframe = True
            self._detect(sf, depth)
            self._fill(sf, depth)
            self.km1 = self.lkf = sf
            self.frame_count += 1
            return sf, True, []
        R = np.asarray(keyframe_R_cur, np.float64)
        ref_R_cur = rs.matmul3(self.keyframe_R_ref.T.copy(), R)
        self._track(self.km1.left_frame, sf.left_frame, ref_R_cur)           # processFrame :236-292
        sf.left_keypoints_rectified = self.cam.undistort_keypoints(sf.left_frame.keypoints)
        self.mono_status = self.stereo_status = ofe.INVALID
        smart = []
        if self.should_be_keyframe(sf.left_frame, self.lkf.left_frame):
            if p.use_ransac:                                                 # handleKeyframe :313-366
                given_rot = not ofe.rot_equals_identity(R)
                st, pose, _ = self.tracker.outlier_rejection_2d2d(self.lkf.left_frame, sf.left_frame,
                                                                  R if (p.ransac_use_2point_mono and given_rot) else None)
                self.mono_status = st
                if st == ofe.VALID:
                    self.lkf_T_k_mono = pose
                self._fill(sf, depth)
                if p.use_stereo_tracking:
                    if p.ransac_use_1point_stereo and given_rot:
                        st, pose, _, _ = self.tracker.outlier_rejection_3d3d_given_rotation(self.lkf, sf, R)
                    else:
                        st, pose, _, _ = self.tracker.outlier_rejection_3d3d(self.lkf, sf)
                    self.stereo_status = st
                    if st == ofe.VALID:
                        self.lkf_T_k_stereo = pose
                else:
                    self.stereo_status = ofe.INVALID
            else:
                self.mono_status = self.stereo_status = ofe.DISABLED
            self._detect(sf, depth)
            self._fill(sf, depth)
            sf.is_keyframe = sf.left_frame.is_keyframe = True
            smart = self.smart_measurements(sf)
            self.lkf = sf
            self.keyframe_R_ref = np.eye(3)
            self.keyframe_count += 1
        else:
            sf.is_keyframe = False
            self.keyframe_R_ref = R
        self.km1 = sf
        self.frame_count += 1
        return sf, sf.is_keyframe, smart
