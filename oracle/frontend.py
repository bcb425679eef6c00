"""oracle/frontend.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Step-by-step CPU restatement of the reference stereo front-end hot path (SURVEY.md section 8(a),
rows a2-a15).  The reference's glue is re-written from the cited lines; each OpenCV call the
reference makes is made here through cv2 4.13 (the same C++ code).  Paths relative to
/root/reference.

  FeatureDetector            src/frontend/feature-detector/FeatureDetector.cpp:94-299
  NonMaximumSuppression      src/frontend/feature-detector/NonMaximumSuppression.cpp:33-169
  anms::TopN                 src/frontend/feature-detector/anms/anms.cpp:37-48
  UndistorterRectifier       src/frontend/UndistorterRectifier.cpp:33-228
  StereoCamera               src/frontend/StereoCamera.cpp:236-290
  OpticalFlowPredictor       src/frontend/optical-flow/OpticalFlowPredictor.cpp:70-126
  Tracker                    src/frontend/Tracker.cpp:92-378, 634-663, 744-769, 836-1018
  StereoMatcher              src/frontend/StereoMatcher.cpp:123-483
  VisionImuFrontend          src/frontend/VisionImuFrontend.cpp:90-232
  StereoVisionImuFrontend    src/frontend/StereoVisionImuFrontend.cpp:245-531

Pinned by the reference's own known answers in tests/test_oracle_pins.py (corner counts
393/400/300/20/200/140, baseline 0.110078, 849/900 shifted-image matches, 35 ground-truth corners,
rotational-flow numbers).  LK tracking and exact sub-pixel values are pinned by cv2 itself only.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import cv2
import numpy as np

from kimera_vio_b200.params import (ANMS_BINNING, ANMS_RANGETREE, ANMS_TOPN, CameraParams, DET_GFTT,
                                    FrontendParams)
from . import ransac as rs
from .ransac import DISABLED, FEW_MATCHES, INVALID, LOW_DISPARITY, VALID
from .rig import StereoRig

# KeypointStatus -- include/kimera-vio/common/vio_types.h:38-44
KP_VALID, KP_NO_LEFT_RECT, KP_NO_RIGHT_RECT, KP_NO_DEPTH, KP_FAILED_ARUN = range(5)

f32 = np.float32


def c_round(x: float) -> int:
    """C round(): half away from zero (StereoMatcher.cpp:298-299, UndistorterRectifier.cpp:157-222)."""
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


# ----------------------------------------------------------------------------------------------
# data carriers (row a15)
# ----------------------------------------------------------------------------------------------
@dataclass
class Frame:
    """include/kimera-vio/frontend/Frame.h:160-186"""
    id: int
    timestamp: int
    img: np.ndarray
    cam: CameraParams
    is_keyframe: bool = False
    keypoints: List[Tuple[np.float32, np.float32]] = field(default_factory=list)
    landmarks: List[int] = field(default_factory=list)
    landmarks_age: List[int] = field(default_factory=list)
    scores: List[float] = field(default_factory=list)
    versors: List[np.ndarray] = field(default_factory=list)

    def nr_valid_keypoints(self) -> int:   # Frame.h:97-105
        return sum(1 for l in self.landmarks if l != -1)

    def clone(self) -> "Frame":
        return Frame(self.id, self.timestamp, self.img, self.cam, self.is_keyframe,
                     list(self.keypoints), list(self.landmarks), list(self.landmarks_age),
                     list(self.scores), [v.copy() for v in self.versors])


@dataclass
class StereoFrame:
    """include/kimera-vio/frontend/StereoFrame.h:141-171"""
    id: int
    timestamp: int
    left_frame: Frame
    right_frame: Frame
    is_keyframe: bool = False
    is_rectified: bool = False
    left_img_rectified: Optional[np.ndarray] = None
    right_img_rectified: Optional[np.ndarray] = None
    left_keypoints_rectified: List[Tuple[int, Tuple[np.float32, np.float32]]] = field(default_factory=list)
    right_keypoints_rectified: List[Tuple[int, Tuple[np.float32, np.float32]]] = field(default_factory=list)
    keypoints_depth: List[float] = field(default_factory=list)
    keypoints_3d: List[np.ndarray] = field(default_factory=list)

    def clone(self) -> "StereoFrame":
        return StereoFrame(self.id, self.timestamp, self.left_frame.clone(), self.right_frame.clone(),
                           self.is_keyframe, self.is_rectified, self.left_img_rectified,
                           self.right_img_rectified, list(self.left_keypoints_rectified),
                           list(self.right_keypoints_rectified), list(self.keypoints_depth),
                           [p.copy() for p in self.keypoints_3d])

    @staticmethod
    def make(id_: int, ts: int, left: np.ndarray, right: np.ndarray, rig: StereoRig) -> "StereoFrame":
        return StereoFrame(id_, ts, Frame(id_, ts, left, rig.left), Frame(id_, ts, right, rig.right))


# ----------------------------------------------------------------------------------------------
# a7: sparse undistortion
# ----------------------------------------------------------------------------------------------
def undistort_rectify_keypoints(kps, cam: CameraParams, R=None, P=None) -> np.ndarray:
    """UndistorterRectifier::UndistortRectifyKeypoints (UndistorterRectifier.cpp:33-68): cv::undistortPoints for
    RADTAN (:41-48), cv::fisheye::undistortPoints for EQUIDISTANT (:49-56)."""
    if len(kps) == 0:
        return np.zeros((0, 2), f32)
    pts = np.asarray(kps, f32).reshape(-1, 1, 2)
    if cam.distortion_model == "equidistant":
        out = cv2.fisheye.undistortPoints(pts, cam.K, cam.D, R=R, P=P)
    elif cam.distortion_model == "radtan":
        out = cv2.undistortPoints(pts, cam.K, cam.D, R=R, P=P)
    else:
        raise NotImplementedError("Unknown distortion model.")      # LOG(FATAL), UndistorterRectifier.cpp:64-66
    return out.reshape(-1, 2)


def get_bearing_vectors(kps, cam: CameraParams, R) -> List[np.ndarray]:
    """UndistorterRectifier::GetBearingVector (UndistorterRectifier.cpp:73-113), vectorised."""
    und = undistort_rectify_keypoints(kps, cam, R, None)
    out = []
    for x, y in und:
        v = np.array([float(x), float(y), 1.0])
        n2 = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])
        out.append(v / math.sqrt(n2) if n2 > 0 else v)
    return out


def crop_to_size(x: np.float32, y: np.float32, W: int, H: int):
    """UtilsOpenCV::cropToSize (src/utils/UtilsOpenCV.cpp:215-235)."""
    cropped = False
    mw, mh = f32(W - 1), f32(H - 1)
    if x > mw:
        x, cropped = mw, True
    elif x < f32(0):
        x, cropped = f32(0), True
    if y > mh:
        y, cropped = mh, True
    elif y < f32(0):
        y, cropped = f32(0), True
    return x, y, cropped


def undistort_rectify_left_keypoints(kps, rig: StereoRig, pixel_tol: float = 2.0):
    """StereoCamera::undistortRectifyLeftKeypoints (StereoCamera.cpp:236-260) +
    checkUndistortedRectifiedLeftKeypoints (UndistorterRectifier.cpp:138-211)."""
    und = undistort_rectify_keypoints(kps, rig.left, rig.R1, rig.P1)
    out = []
    tol = f32(pixel_tol)
    for (dx, dy), (ux, uy) in zip(kps, und):
        ux, uy, cropped = crop_to_size(f32(ux), f32(uy), rig.W, rig.H)
        ex = rig.map_lx[c_round(uy), c_round(ux)]
        ey = rig.map_ly[c_round(uy), c_round(ux)]
        if cropped:
            out.append((KP_NO_LEFT_RECT, (ux, uy)))
        elif abs(f32(dx) - ex) > tol or abs(f32(dy) - ey) > tol:
            out.append((KP_NO_LEFT_RECT, (ux, uy)))
        else:
            out.append((KP_VALID, (ux, uy)))
    return out


def distort_unrectify_right_keypoints(right_rect, rig: StereoRig):
    """UndistorterRectifier::distortUnrectifyKeypoints (UndistorterRectifier.cpp:213-228), right cam."""
    out = []
    for st, (x, y) in right_rect:
        if st == KP_VALID:
            out.append((rig.map_rx[c_round(y), c_round(x)], rig.map_ry[c_round(y), c_round(x)]))
        else:
            out.append((f32(0), f32(0)))
    return out


# ----------------------------------------------------------------------------------------------
# a3-a6: detection
# ----------------------------------------------------------------------------------------------
class FeatureDetector:
    def __init__(self, p: FrontendParams):
        self.p = p
        assert p.feature_detector_type == DET_GFTT, "only GFTT is on the graded path (all 7 shipped rigs)"
        # FeatureDetector.cpp:71-82
        self.gftt = cv2.GFTTDetector_create(p.max_nr_keypoints_before_anms, p.quality_level,
                                            float(p.min_distance), p.block_size,
                                            p.use_harris_detector, p.k)
        self.lmk_id = 0       # FeatureDetector.cpp:141 (function-static: one per process)

    def build_mask(self, frame: Frame) -> np.ndarray:
        """FeatureDetector.cpp:185-203"""
        mask = np.full(frame.img.shape, 255, np.uint8)
        for kp, lmk in zip(frame.keypoints, frame.landmarks):
            if lmk != -1:
                # cv::Point(Point2f) rounds with cvRound (half to even)
                c = (int(np.rint(f32(kp[0]))), int(np.rint(f32(kp[1]))))
                cv2.circle(mask, c, self.p.min_distance, 0, cv2.FILLED)
        return mask

    def raw_feature_detection(self, img: np.ndarray, mask: np.ndarray):
        return list(self.gftt.detect(img, mask))

    def suppress_non_max(self, kps, num_ret: int, cols: int, rows: int):
        """AdaptiveNonMaximumSuppression::suppressNonMax (NonMaximumSuppression.cpp:33-122)."""
        p = self.p
        if len(kps) == 0:
            return []
        resp = np.array([int(k.response) for k in kps], np.int32)   # truncation to int (:50-53)
        idx = cv2.sortIdx(resp.reshape(1, -1), cv2.SORT_DESCENDING + cv2.SORT_EVERY_ROW).reshape(-1)
        sorted_kps = [kps[i] for i in idx]
        t = p.non_max_suppression_type
        if t == ANMS_TOPN:          # receives the UNSORTED list (:67)
            if num_ret > len(kps):
                return list(kps)
            return list(kps[:num_ret])
        if t == ANMS_BINNING:
            return self.binning(sorted_kps, num_ret, cols, rows)
        if t == ANMS_RANGETREE:
            return self.range_tree(sorted_kps, num_ret, 0.1, cols, rows)
        raise NotImplementedError("ANMS type %d is not on the graded path" % t)

    def binning(self, kps, num_ret: int, cols: int, rows: int):
        """AdaptiveNonMaximumSuppression::binning (NonMaximumSuppression.cpp:125-169)."""
        p = self.p
        if num_ret > len(kps):
            return list(kps)
        bin_row = f32(rows) / f32(p.nr_vertical_bins)
        bin_col = f32(cols) / f32(p.nr_horizontal_bins)
        n_active = f32(p.binning_mask.sum())
        per_bin = int(c_round(f32(num_ret) / n_active))
        out = []
        cnt = np.zeros((p.nr_vertical_bins, p.nr_horizontal_bins))
        for k in kps:
            r = int(f32(k.pt[1]) / bin_row)
            c = int(f32(k.pt[0]) / bin_col)
            if p.binning_mask[r, c] == 1 and cnt[r, c] < per_bin:
                out.append(k)
                cnt[r, c] += 1
        return out

    @staticmethod
    def range_tree(kps, num_ret: int, tolerance: float, cols: int, rows: int):
        """anms::RangeTree (anms/anms.cpp:254-340); the range tree is replaced by a brute-force
        inclusive box query on the u16-truncated coordinates (alternate ANMS, struct default only)."""
        n = len(kps)
        exp1 = rows + cols + 2 * num_ret
        exp2 = (4 * cols + 4 * num_ret + 4 * rows * num_ret + rows * rows + cols * cols
                - 2 * rows * cols + 4 * rows * cols * num_ret)
        exp3 = math.sqrt(exp2)
        exp4 = num_ret - 1
        sol1 = -c_round((exp1 + exp3) / exp4)
        sol2 = -c_round((exp1 - exp3) / exp4)
        high = int(sol1 if sol1 > sol2 else sol2)
        low = int(math.floor(math.sqrt(n / num_ret)))
        xs = np.array([int(k.pt[0]) for k in kps])
        ys = np.array([int(k.pt[1]) for k in kps])
        K = num_ret
        kmin, kmax = c_round(K - K * tolerance), c_round(K + K * tolerance)
        prevwidth, result, final = -1, [], []
        while True:
            included = np.ones(n, bool)
            width = low + int((high - low) / 2)
            if width == prevwidth or low > high:
                final = result
                break
            result = []
            for i in range(n):
                if included[i]:
                    included[i] = False
                    result.append(i)
                    minx, maxx = max(int(kps[i].pt[0] - width), 0), int(kps[i].pt[0] + width)
                    miny, maxy = max(int(kps[i].pt[1] - width), 0), int(kps[i].pt[1] + width)
                    hit = (xs >= minx) & (xs <= maxx) & (ys >= miny) & (ys <= maxy)
                    included[hit] = False
            if kmin <= len(result) <= kmax:
                final = result
                break
            elif len(result) < kmin:
                high = width - 1
            else:
                low = width + 1
            prevwidth = width
        return [kps[i] for i in final]

    def detect_corners(self, frame: Frame, need: int) -> np.ndarray:
        """FeatureDetector::featureDetection(const Frame&, int) (FeatureDetector.cpp:174-299)."""
        p = self.p
        mask = self.build_mask(frame)
        kps = self.raw_feature_detection(frame.img, mask)
        if p.enable_non_max_suppression:
            kps = self.suppress_non_max(kps, need, frame.img.shape[1], frame.img.shape[0])
        corners = np.array([k.pt for k in kps], f32).reshape(-1, 2)
        if len(corners) > 0 and p.enable_subpixel_corner_refinement:
            c = corners.reshape(-1, 1, 2).copy()
            crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, p.subpix_max_iters, p.subpix_epsilon)
            cv2.cornerSubPix(frame.img, c, (p.subpix_window_size, p.subpix_window_size),
                             (p.subpix_zero_zone, p.subpix_zero_zone), crit)
            corners = c.reshape(-1, 2)
        return corners

    def feature_detection(self, frame: Frame, R: Optional[np.ndarray]) -> None:
        """FeatureDetector::featureDetection(Frame*, R) (FeatureDetector.cpp:94-163)."""
        n_existing = 0
        for i in range(len(frame.landmarks)):
            if frame.landmarks[i] != -1:
                n_existing += 1
            frame.landmarks_age[i] += 1
        need = max(self.p.max_features_per_frame - n_existing, 0)
        corners = self.detect_corners(frame, need)
        if len(corners) > 0:
            versors = get_bearing_vectors(corners, frame.cam, R)
            for c, v in zip(corners, versors):
                frame.landmarks.append(self.lmk_id)
                frame.landmarks_age.append(1)
                frame.keypoints.append((f32(c[0]), f32(c[1])))
                frame.scores.append(0.0)
                frame.versors.append(v)
                self.lmk_id += 1


# ----------------------------------------------------------------------------------------------
# a8: rotational optical-flow prediction
# ----------------------------------------------------------------------------------------------
def matx33f_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """cv::Matx<float,3,3> product: s = 0; s += a(i,k)*b(k,j) in float."""
    c = np.zeros((3, 3), f32)
    for i in range(3):
        for j in range(3):
            s = f32(0)
            for k in range(3):
                s = f32(s + f32(a[i, k] * b[k, j]))
            c[i, j] = s
    return c


def matx33f_inv(a: np.ndarray) -> np.ndarray:
    """cv::Matx33f::inv() (Matx_FastInvOp<float,3,3>), float arithmetic."""
    a = a.astype(f32)
    d = f32(a[0, 0] * f32(a[1, 1] * a[2, 2] - a[2, 1] * a[1, 2])
            - a[0, 1] * f32(a[1, 0] * a[2, 2] - a[2, 0] * a[1, 2])
            + a[0, 2] * f32(a[1, 0] * a[2, 1] - a[2, 0] * a[1, 1]))
    d = f32(1) / d
    b = np.zeros((3, 3), f32)
    b[0, 0] = f32(a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) * d
    b[0, 1] = f32(a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]) * d
    b[0, 2] = f32(a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]) * d
    b[1, 0] = f32(a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) * d
    b[1, 1] = f32(a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]) * d
    b[1, 2] = f32(a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]) * d
    b[2, 0] = f32(a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]) * d
    b[2, 1] = f32(a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]) * d
    b[2, 2] = f32(a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]) * d
    return b


def quaternion_w(R: np.ndarray) -> float:
    """Eigen::Quaterniond(Matrix3d).w() as used by gtsam::Rot3::toQuaternion()."""
    t = R[0, 0] + (R[1, 1] + R[2, 2])
    if t > 0:
        return 0.5 * math.sqrt(t + 1.0)
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    return (R[k, j] - R[j, k]) * (0.5 / t)


def homography_from_rotation(K: np.ndarray, R: np.ndarray) -> np.ndarray:
    Kf = K.astype(f32)
    Rf = np.asarray(R, np.float64).astype(f32)
    return matx33f_mul(matx33f_mul(Kf, Rf.T.copy()), matx33f_inv(Kf))


def predict_sparse_flow(prev_kps, R: np.ndarray, K: np.ndarray, W: int, H: int, predictor_type: int):
    """RotationalOpticalFlowPredictor::predictSparseFlow (OpticalFlowPredictor.cpp:70-126)."""
    prev = [(f32(x), f32(y)) for x, y in prev_kps]
    if predictor_type == 0:
        return list(prev)
    if abs(1.0 - abs(quaternion_w(np.asarray(R, np.float64)))) < 1e-4:
        return list(prev)
    Hm = homography_from_rotation(K, R)
    out = []
    for x, y in prev:
        p = []
        for i in range(3):
            s = f32(0)
            s = f32(s + f32(Hm[i, 0] * x))
            s = f32(s + f32(Hm[i, 1] * y))
            s = f32(s + f32(Hm[i, 2] * f32(1)))
            p.append(s)
        if p[2] > f32(0):
            nx, ny = f32(p[0] / p[2]), f32(p[1] / p[2])
        else:
            nx, ny = x, y
        # cv::Rect2f(0,0,W,H).contains
        if nx >= f32(0) and nx < f32(W) and ny >= f32(0) and ny < f32(H):
            out.append((nx, ny))
        else:
            out.append((x, y))
    return out


# ----------------------------------------------------------------------------------------------
# a9, a12-a14: tracker
# ----------------------------------------------------------------------------------------------
def find_matching_keypoints(ref: Frame, cur: Frame) -> List[Tuple[int, int]]:
    """Tracker::findMatchingKeypoints (Tracker.cpp:919-946)."""
    m = {}
    for i, l in enumerate(ref.landmarks):
        if l != -1:
            m[l] = i
    out = []
    for i, l in enumerate(cur.landmarks):
        if l != -1 and l in m:
            out.append((m[l], i))
    return out


def find_matching_stereo_keypoints(ref: StereoFrame, cur: StereoFrame) -> List[Tuple[int, int]]:
    """Tracker::findMatchingStereoKeypoints (Tracker.cpp:948-989)."""
    out = []
    for (ir, ic) in find_matching_keypoints(ref.left_frame, cur.left_frame):
        if ref.right_keypoints_rectified[ir][0] == KP_VALID and cur.right_keypoints_rectified[ic][0] == KP_VALID:
            out.append((ir, ic))
    return out


def compute_median_disparity(ref_kps, cur_kps, matches) -> Tuple[bool, float]:
    """Tracker::computeMedianDisparity (Tracker.cpp:991-1018)."""
    d = []
    for (ir, ic) in matches:
        dx = f32(f32(cur_kps[ic][0]) - f32(ref_kps[ir][0]))
        dy = f32(f32(cur_kps[ic][1]) - f32(ref_kps[ir][1]))
        d.append(float(f32(f32(dx * dx) + f32(dy * dy))))
    if not d:
        return False, 0.0
    c = len(d) // 2
    return True, math.sqrt(float(np.partition(np.array(d), c)[c]))


class Tracker:
    def __init__(self, p: FrontendParams, rig: StereoRig, rnd_libstdcxx: str = "lemire"):
        self.p, self.rig = p, rig
        self.rnd_libstdcxx = rnd_libstdcxx
        self.debug = {}

    def _rnd(self):
        assert not self.p.ransac_randomize, "time-seeded RANSAC cannot be compared"
        return rs.rnd_table(16384, 12345, self.rnd_libstdcxx)

    # --- a9
    def feature_tracking(self, ref: Frame, cur: Frame, ref_R_cur: np.ndarray) -> None:
        """Tracker::featureTracking (Tracker.cpp:92-211)."""
        p = self.p
        idx = [i for i, l in enumerate(ref.landmarks) if l != -1]
        px_ref = [ref.keypoints[i] for i in idx]
        px_cur = predict_sparse_flow(px_ref, ref_R_cur, ref.cam.K, ref.img.shape[1], ref.img.shape[0],
                                     p.optical_flow_predictor_type)
        self.debug["px_ref"] = np.array(px_ref, f32).reshape(-1, 2)
        self.debug["px_pred"] = np.array(px_cur, f32).reshape(-1, 2)
        if len(px_ref) > 0:
            a = np.array(px_ref, f32).reshape(-1, 1, 2)
            b = np.array(px_cur, f32).reshape(-1, 1, 2)
            crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, p.klt_max_iter, p.klt_eps)
            nxt, status, _err = cv2.calcOpticalFlowPyrLK(ref.img, cur.img, a, b,
                                                         winSize=(p.klt_win_size, p.klt_win_size),
                                                         maxLevel=p.klt_max_level, criteria=crit,
                                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            nxt, status = nxt.reshape(-1, 2), status.reshape(-1)
        else:
            nxt, status = np.zeros((0, 2), f32), np.zeros((0,), np.uint8)
        self.debug["lk_next"], self.debug["lk_status"] = nxt.copy(), status.copy()
        assert len(cur.keypoints) == 0
        keep = []
        for j, i in enumerate(idx):
            if (not status[j]) or ref.landmarks_age[i] > p.max_feature_track_age:
                ref.landmarks[i] = -1
                continue
            keep.append(j)
            cur.landmarks.append(ref.landmarks[i])
            cur.landmarks_age.append(ref.landmarks_age[i])
            cur.scores.append(ref.scores[i])
            cur.keypoints.append((f32(nxt[j, 0]), f32(nxt[j, 1])))
        cur.versors.extend(get_bearing_vectors([cur.keypoints[k] for k in range(len(keep))], ref.cam, self.rig.R1))

    # --- a12
    def outlier_rejection_2d2d(self, ref: Frame, cur: Frame, R: Optional[np.ndarray]):
        """Tracker::geometricOutlierRejection2d2d (Tracker.cpp:213-378).  R=None -> the default argument
        gtsam::Pose3() (Tracker.h:97-100).  The solver is chosen from ransac_use_2point_mono ALONE
        (Tracker.cpp:248-276): the call outlierRejectionMono labels "5-point RANSAC" (VisionImuFrontend.cpp:108)
        is the 2-point problem with R12 = identity whenever that parameter is set."""
        p = self.p
        matches = find_matching_keypoints(ref, cur)
        ident = np.hstack([np.eye(3), np.zeros((3, 1))])
        self.debug["mono_matches"] = list(matches)
        if not matches:
            return INVALID, ident, []
        f_ref = [ref.versors[ir] for ir, _ in matches]
        f_cur = [cur.versors[ic] for _, ic in matches]
        if p.ransac_use_2point_mono:
            R12 = np.eye(3) if R is None else np.asarray(R, np.float64)
            prob = rs.Problem2d2dGivenRot(f_ref, f_cur, R12, self._rnd())
        else:
            prob = rs.Problem2d2dNister(f_ref, f_cur, self._rnd())
        ok, pose, inliers = rs.run_ransac(prob, p.ransac_threshold_mono, p.ransac_max_iterations,
                                          p.ransac_probability)
        if not ok:
            status, pose = INVALID, ident
        else:
            status = VALID
            if len(inliers) < p.min_nr_mono_inliers:
                status = FEW_MATCHES
        self.debug["mono_inliers"] = list(inliers)
        if status != FEW_MATCHES:
            # removeOutliersMono (Tracker.cpp:856-882)
            for o in rs.find_outliers(len(matches), inliers):
                ir, ic = matches[o]
                ref.landmarks[ir] = -1
                cur.landmarks[ic] = -1
            matches = [matches[i] for i in inliers]
        if status == VALID:
            ok2, disp = compute_median_disparity(ref.keypoints, cur.keypoints, matches)
            if ok2 and disp < p.disparity_threshold:
                status = LOW_DISPARITY
        return status, pose, inliers

    # --- a13
    def _remove_outliers_stereo(self, inliers, ref: StereoFrame, cur: StereoFrame, matches):
        """Tracker::removeOutliersStereo (Tracker.cpp:884-917)."""
        for o in rs.find_outliers(len(matches), inliers):
            ir, ic = matches[o]
            ref.right_keypoints_rectified[ir] = (KP_FAILED_ARUN, ref.right_keypoints_rectified[ir][1])
            ref.keypoints_depth[ir] = 0.0
            ref.keypoints_3d[ir] = np.zeros(3)
            cur.right_keypoints_rectified[ic] = (KP_FAILED_ARUN, cur.right_keypoints_rectified[ic][1])
            cur.keypoints_depth[ic] = 0.0
            cur.keypoints_3d[ic] = np.zeros(3)

    def outlier_rejection_3d3d_given_rotation(self, ref: StereoFrame, cur: StereoFrame, R: np.ndarray):
        """Tracker::geometricOutlierRejection3d3dGivenRotation (Tracker.cpp:634-663 -> :382-632)."""
        p, rig = self.p, self.rig
        matches = find_matching_stereo_keypoints(ref, cur)
        self.debug["stereo_matches"] = list(matches)
        calib = (rig.fx, rig.fy, rig.cx, rig.cy, rig.baseline)
        rl = [k[1] for k in ref.left_keypoints_rectified]
        rr = [k[1] for k in ref.right_keypoints_rectified]
        cl = [k[1] for k in cur.left_keypoints_rectified]
        cr = [k[1] for k in cur.right_keypoints_rectified]
        status, pose, inliers, info = rs.outlier_rejection_3d3d_given_rotation(
            rl, rr, cl, cr, ref.keypoints_3d, cur.keypoints_3d, calib, matches, R,
            p.ransac_threshold_stereo, p.min_nr_stereo_inliers)
        self.debug["stereo_inliers"] = list(inliers)
        self._remove_outliers_stereo(inliers, ref, cur, matches)
        return status, pose, inliers, info

    def outlier_rejection_3d3d(self, ref: StereoFrame, cur: StereoFrame):
        """Tracker::geometricOutlierRejection3d3d (Tracker.cpp:744-769 -> :667-742)."""
        p = self.p
        matches = find_matching_stereo_keypoints(ref, cur)
        self.debug["stereo_matches"] = list(matches)
        ident = np.hstack([np.eye(3), np.zeros((3, 1))])
        p_ref = [ref.keypoints_3d[ir] for ir, _ in matches]
        p_cur = [cur.keypoints_3d[ic] for _, ic in matches]
        prob = rs.Problem3d3d(p_ref, p_cur, self._rnd()) if matches else None
        if prob is None:
            ok, pose, inliers = False, ident, []
        else:
            ok, pose, inliers = rs.run_ransac(prob, p.ransac_threshold_stereo, p.ransac_max_iterations,
                                              p.ransac_probability)
        if not ok:
            status, pose = INVALID, ident
        else:
            status = VALID if len(inliers) >= p.min_nr_stereo_inliers else FEW_MATCHES
        self.debug["stereo_inliers"] = list(inliers)
        if status != INVALID:
            self._remove_outliers_stereo(inliers, ref, cur, matches)
        return status, pose, inliers, np.zeros((3, 3))


# ----------------------------------------------------------------------------------------------
# a10, a11: sparse stereo
# ----------------------------------------------------------------------------------------------
class StereoMatcher:
    def __init__(self, p: FrontendParams, rig: StereoRig):
        self.p, self.rig = p, rig

    def stripe_geometry(self, fx: float, baseline: float, cols: int) -> Tuple[int, int]:
        """StereoMatcher.cpp:214-231"""
        p = self.p
        stripe_rows = p.templ_rows + p.stripe_extra_rows
        stripe_cols = c_round(fx * baseline / p.min_point_dist) + p.templ_cols + 4
        if stripe_cols % 2 != 1:
            stripe_cols += 1
        if stripe_cols > cols:
            stripe_cols = cols
        return stripe_cols, stripe_rows

    def search_right_keypoint_epipolar(self, left_rect, kp, right_rect, stripe_cols, stripe_rows):
        """StereoMatcher::searchRightKeypointEpipolar (StereoMatcher.cpp:283-423)."""
        p = self.p
        rx, ry = c_round(kp[0]), c_round(kp[1])
        rows, cols = left_rect.shape
        tcy = ry - (p.templ_rows - 1) // 2
        if tcy < 0 or tcy + p.templ_rows > rows - 1:
            return (KP_NO_RIGHT_RECT, (f32(0), f32(0))), -1.0
        offset_temp = 0
        tcx = rx - (p.templ_cols - 1) // 2
        if tcx < 0:
            offset_temp = tcx
            tcx = 0
        if tcx + p.templ_cols > cols - 1:
            assert offset_temp == 0, "Offset_temp cannot exceed in both directions!"
            offset_temp = (tcx + p.templ_cols) - (cols - 1)
            tcx -= offset_temp
        templ = left_rect[tcy:tcy + p.templ_rows, tcx:tcx + p.templ_cols]
        scy = ry - (stripe_rows - 1) // 2
        if scy < 0 or scy + stripe_rows > right_rect.shape[0] - 1:
            return (KP_NO_RIGHT_RECT, (f32(0), f32(0))), -1.0
        scx = rx + (p.templ_cols - 1) // 2 - stripe_cols
        if scx + stripe_cols > right_rect.shape[1] - 1:
            scx -= (scx + stripe_cols) - (right_rect.shape[1] - 1)
        if scx < 0:
            scx = 0
        stripe = right_rect[scy:scy + stripe_rows, scx:scx + stripe_cols]
        result = cv2.matchTemplate(stripe, templ, cv2.TM_SQDIFF)
        result = cv2.normalize(result, None, 0, 1, cv2.NORM_MINMAX, -1)
        min_val, _max_val, min_loc, _max_loc = cv2.minMaxLoc(result)
        mx = min_loc[0] + scx + (p.templ_cols - 1) // 2 + offset_temp
        my = min_loc[1] + scy + (p.templ_rows - 1) // 2
        match = (f32(mx), f32(my))
        if p.subpixel_refinement_stereo:
            c = np.array([[match]], f32)
            cv2.cornerSubPix(right_rect, c, (10, 10), (-1, -1),
                             (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 40, 0.001))
            match = (f32(c[0, 0, 0]), f32(c[0, 0, 1]))
        if min_val < p.tolerance_template_matching:
            return (KP_VALID, match), min_val
        return (KP_NO_RIGHT_RECT, match), min_val

    def exact_sqdiff(self, left_rect, kp, right_rect, stripe_cols, stripe_rows):
        """The template / stripe of searchRightKeypointEpipolar for one keypoint and its TM_SQDIFF map in EXACT
        integers (cv::matchTemplate goes through a float DFT): returns (map int64 [rows x cols], scx, scy, offset_temp)
        or None when the keypoint is rejected before matching.  Used by the tests to classify near-tie divergences."""
        p = self.p
        rx, ry = c_round(kp[0]), c_round(kp[1])
        rows, cols = left_rect.shape
        tcy = ry - (p.templ_rows - 1) // 2
        if tcy < 0 or tcy + p.templ_rows > rows - 1:
            return None
        offset_temp = 0
        tcx = rx - (p.templ_cols - 1) // 2
        if tcx < 0:
            offset_temp, tcx = tcx, 0
        if tcx + p.templ_cols > cols - 1:
            offset_temp = (tcx + p.templ_cols) - (cols - 1)
            tcx -= offset_temp
        templ = left_rect[tcy:tcy + p.templ_rows, tcx:tcx + p.templ_cols].astype(np.int64)
        scy = ry - (stripe_rows - 1) // 2
        if scy < 0 or scy + stripe_rows > right_rect.shape[0] - 1:
            return None
        scx = rx + (p.templ_cols - 1) // 2 - stripe_cols
        if scx + stripe_cols > right_rect.shape[1] - 1:
            scx -= (scx + stripe_cols) - (right_rect.shape[1] - 1)
        if scx < 0:
            scx = 0
        stripe = right_rect[scy:scy + stripe_rows, scx:scx + stripe_cols].astype(np.int64)
        ny, nx = stripe.shape[0] - templ.shape[0] + 1, stripe.shape[1] - templ.shape[1] + 1
        out = np.zeros((ny, nx), np.int64)
        for dy in range(ny):
            win = np.lib.stride_tricks.sliding_window_view(stripe[dy:dy + templ.shape[0]], templ.shape[1], axis=1)  # rows x nx x tc
            out[dy] = ((win - templ[:, None, :]) ** 2).sum(axis=(0, 2))
        return out, scx, scy, offset_temp

    def get_right_keypoints_rectified(self, left_rect, right_rect, left_kps_rect, fx, baseline):
        """StereoMatcher::getRightKeypointsRectified (StereoMatcher.cpp:196-281)."""
        stripe_cols, stripe_rows = self.stripe_geometry(fx, baseline, right_rect.shape[1])
        out = []
        for st, kp in left_kps_rect:
            if st != KP_VALID:
                out.append((st, (f32(0), f32(0))))
                continue
            r, _ = self.search_right_keypoint_epipolar(left_rect, kp, right_rect, stripe_cols, stripe_rows)
            out.append(r)
        return out

    def get_depth_from_rectified_matches(self, left_kps, right_kps) -> List[float]:
        """StereoMatcher::getDepthFromRectifiedMatches (StereoMatcher.cpp:425-483); mutates right."""
        p = self.p
        fx_b = self.rig.fx * self.rig.baseline
        depths = []
        for i in range(len(left_kps)):
            ls, lp = left_kps[i]
            rs_, rp = right_kps[i]
            if ls == KP_VALID and rs_ == KP_VALID:
                disparity = float(f32(f32(lp[0]) - f32(rp[0])))
                if disparity >= 0.0:
                    depth = fx_b / disparity if disparity != 0.0 else math.inf
                    if depth < p.min_point_dist or depth > p.max_point_dist:
                        right_kps[i] = (KP_NO_DEPTH, rp)
                        depths.append(0.0)
                    else:
                        depths.append(depth)
                else:
                    right_kps[i] = (KP_NO_DEPTH, rp)
                    depths.append(0.0)
            else:
                if ls != KP_VALID and rs_ != ls:
                    right_kps[i] = (ls, rp)
                depths.append(0.0)
        return depths

    def sparse_stereo_reconstruction(self, sf: StereoFrame) -> None:
        """StereoMatcher::sparseStereoReconstruction(StereoFrame*) (StereoMatcher.cpp:123-175)."""
        rig = self.rig
        sf.left_img_rectified = rig.rectify_left(sf.left_frame.img)      # StereoCamera.cpp:269-290
        sf.right_img_rectified = rig.rectify_right(sf.right_frame.img)
        sf.is_rectified = True
        assert len(sf.left_frame.keypoints) > 0, "Call feature detection on left frame first..."
        sf.left_keypoints_rectified = undistort_rectify_left_keypoints(sf.left_frame.keypoints, rig)
        sf.right_keypoints_rectified = self.get_right_keypoints_rectified(
            sf.left_img_rectified, sf.right_img_rectified, sf.left_keypoints_rectified, rig.fx, rig.baseline)
        sf.keypoints_depth = self.get_depth_from_rectified_matches(sf.left_keypoints_rectified,
                                                                   sf.right_keypoints_rectified)
        sf.right_frame.keypoints = distort_unrectify_right_keypoints(sf.right_keypoints_rectified, rig)
        sf.keypoints_3d = []
        for i, (st, _) in enumerate(sf.right_keypoints_rectified):
            if st == KP_VALID:
                v = sf.left_frame.versors[i]
                assert v[2] >= 1e-3
                sf.keypoints_3d.append(v * sf.keypoints_depth[i] / v[2])
            else:
                sf.keypoints_3d.append(np.zeros(3))


# ----------------------------------------------------------------------------------------------
# front-end FSM (caller of the hot path; restated so whole sequences can be compared)
# ----------------------------------------------------------------------------------------------
def rot_equals_identity(R: np.ndarray, tol: float = 1e-9) -> bool:
    """gtsam::Rot3::equals(Rot3(), 1e-9)"""
    return bool(np.all(np.abs(np.asarray(R, np.float64) - np.eye(3)) <= tol))


@dataclass
class FrontendOutput:
    is_keyframe: bool
    frame: StereoFrame                    # deep copy of the stereo frame handed to the packet
    mono_status: int
    stereo_status: int
    lkf_T_k_mono: np.ndarray
    lkf_T_k_stereo: np.ndarray
    smart_measurements: List[Tuple[int, float, float, float]]
    debug: dict


class StereoFrontend:
    """StereoVisionImuFrontend restated without IMU/back-end plumbing: the caller provides the
    relative rotation camLrectLkf_R_camLrectK (StereoVisionImuFrontend.cpp:149-150) directly."""

    def __init__(self, p: FrontendParams, rig: StereoRig, rnd_libstdcxx: str = "lemire"):
        self.p, self.rig = p, rig
        self.detector = FeatureDetector(p)
        self.tracker = Tracker(p, rig, rnd_libstdcxx)
        self.matcher = StereoMatcher(p, rig)
        self.km1: Optional[StereoFrame] = None
        self.lkf: Optional[StereoFrame] = None
        self.keyframe_R_ref = np.eye(3)
        self.frame_count = 0
        self.keyframe_count = 0
        self.mono_status, self.stereo_status = INVALID, INVALID
        ident = np.hstack([np.eye(3), np.zeros((3, 1))])
        self.lkf_T_k_mono, self.lkf_T_k_stereo = ident.copy(), ident.copy()
        self.info_stereo = np.zeros((3, 3))
        self.force_keyframe = False

    # StereoVisionImuFrontend::processFirstStereoFrame (StereoVisionImuFrontend.cpp:245-276)
    def process_first(self, sf_in: StereoFrame) -> FrontendOutput:
        k = sf_in.clone()
        k.is_keyframe = True
        assert len(k.left_frame.keypoints) == 0
        self.detector.feature_detection(k.left_frame, self.rig.R1)
        self.matcher.sparse_stereo_reconstruction(k)
        self.km1 = k
        self.lkf = k
        self.frame_count += 1
        return self._output(k, True, [])

    def spin(self, sf_in: StereoFrame, keyframe_R_cur: np.ndarray) -> FrontendOutput:
        if self.frame_count == 0:
            return self.process_first(sf_in)
        return self.process(sf_in, np.asarray(keyframe_R_cur, np.float64))

    def should_be_keyframe(self, frame: Frame, frame_lkf: Frame) -> bool:
        """VisionImuFrontend::shouldBeKeyframe (VisionImuFrontend.cpp:175-232)."""
        p = self.p
        kf_diff = frame.timestamp - frame_lkf.timestamp
        nr_valid = frame.nr_valid_keypoints()
        min_time = kf_diff >= p.min_intra_keyframe_time_ns
        max_time = kf_diff >= p.max_intra_keyframe_time_ns
        nr_low = nr_valid <= p.min_number_features
        matches = find_matching_keypoints(frame_lkf, frame)
        _, disparity = compute_median_disparity(frame_lkf.keypoints, frame.keypoints, matches)
        self._last_disparity = disparity
        is_low = disparity < p.disparity_threshold
        low_first = is_low and not (self.mono_status == LOW_DISPARITY)
        enough = not is_low
        max_disp = disparity > p.max_disparity_since_lkf
        flipped = (enough or low_first) and min_time
        return bool(max_time or max_disp or flipped or nr_low or frame.is_keyframe)

    # StereoVisionImuFrontend::processStereoFrame (StereoVisionImuFrontend.cpp:283-481)
    def process(self, sf_in: StereoFrame, keyframe_R_cur: np.ndarray) -> FrontendOutput:
        p, rig, tr = self.p, self.rig, self.tracker
        k = sf_in.clone()
        left_k = k.left_frame
        ref_R_cur = rs.matmul3(self.keyframe_R_ref.T.copy(), keyframe_R_cur)
        tr.feature_tracking(self.km1.left_frame, left_k, ref_R_cur)
        if len(left_k.keypoints) == 0:
            self.detector.feature_detection(left_k, rig.R1)
            self.km1 = k
            self.frame_count += 1
            return self._output(k, False, [])
        smart = []
        new_kf = self.should_be_keyframe(left_k, self.lkf.left_frame)
        if new_kf:
            self.keyframe_count += 1
            self.mono_status, self.stereo_status = INVALID, INVALID
            if p.use_ransac:
                given_rot = not rot_equals_identity(keyframe_R_cur)
                imu_ok = given_rot      # time_aligned: frontend_state_ == Nominal
                # outlierRejectionMono (VisionImuFrontend.cpp:90-113)
                if p.ransac_use_2point_mono and imu_ok:
                    st, pose, _ = tr.outlier_rejection_2d2d(self.lkf.left_frame, left_k, keyframe_R_cur)
                else:
                    st, pose, _ = tr.outlier_rejection_2d2d(self.lkf.left_frame, left_k, None)
                self.mono_status = st
                if st == VALID:
                    self.lkf_T_k_mono = pose
                self.matcher.sparse_stereo_reconstruction(k)
                if p.use_stereo_tracking:
                    # outlierRejectionStereo (VisionImuFrontend.cpp:115-144)
                    if p.ransac_use_1point_stereo and imu_ok:
                        st, pose, _, info = tr.outlier_rejection_3d3d_given_rotation(self.lkf, k, keyframe_R_cur)
                    else:
                        st, pose, _, info = tr.outlier_rejection_3d3d(self.lkf, k)
                    self.info_stereo = info
                    self.stereo_status = st
                    if st == VALID:
                        self.lkf_T_k_stereo = pose
                else:
                    self.stereo_status = INVALID
            else:
                self.mono_status, self.stereo_status = DISABLED, DISABLED
            k.is_keyframe = True
            k.left_frame.is_keyframe = True
            self.detector.feature_detection(left_k, rig.R1)
            self.matcher.sparse_stereo_reconstruction(k)
            self.lkf = k
            smart = self.get_smart_stereo_measurements(k)
        else:
            k.is_keyframe = False
        self.keyframe_R_ref = np.eye(3) if k.is_keyframe else keyframe_R_cur.copy()
        self.km1 = k
        self.frame_count += 1
        return self._output(k, k.is_keyframe, smart)

    def get_smart_stereo_measurements(self, sf: StereoFrame):
        """StereoVisionImuFrontend::getSmartStereoMeasurements (StereoVisionImuFrontend.cpp:485-531)."""
        out = []
        for i, l in enumerate(sf.left_frame.landmarks):
            if l == -1:
                continue
            uL = float(sf.left_keypoints_rectified[i][1][0])
            v = float(sf.left_keypoints_rectified[i][1][1])
            uR = float("nan")
            if self.p.use_stereo_tracking and sf.right_keypoints_rectified[i][0] == KP_VALID:
                uR = float(sf.right_keypoints_rectified[i][1][0])
            out.append((l, uL, uR, v))
        return out

    def _output(self, k: StereoFrame, is_kf: bool, smart) -> FrontendOutput:
        return FrontendOutput(is_kf, k.clone(), self.mono_status, self.stereo_status,
                              self.lkf_T_k_mono.copy(), self.lkf_T_k_stereo.copy(), smart,
                              dict(self.tracker.debug))
