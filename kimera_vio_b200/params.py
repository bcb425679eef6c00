"""Front-end parameter structs and YAML parsing.

Mirrors the YAML keys the reference parses (file:line, relative to /root/reference):
  * TrackerParams::parseYAML            src/frontend/VisionImuTrackerParams.cpp:84-135
  * FeatureDetectorParams::parseYAML    src/frontend/feature-detector/FeatureDetectorParams.cpp:105-222
  * SubPixelCornerFinderParams          src/frontend/feature-detector/FeatureDetectorParams.cpp:43-75
  * StereoMatchingParams::parseYAML     src/frontend/StereoMatchingParams.cpp:80-90
  * FrontendParams::parseYAML           src/frontend/VisionImuFrontendParams.cpp:80-112
  * CameraParams::parseYAML             src/frontend/CameraParams.cpp:30-120

Struct defaults are the reference's struct defaults (FeatureDetectorParams.h:46-106,
VisionImuTrackerParams.h:42-70, StereoMatchingParams.h, VisionImuFrontendParams.h:40-60) so that
a partially specified YAML behaves identically.
"""
from __future__ import annotations

import dataclasses
import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import yaml


def _load_cv_yaml(path: str) -> dict:
    """cv::FileStorage YAML: '%YAML:1.0' header, otherwise plain YAML 1.0."""
    with open(path, "r") as f:
        txt = f.read()
    txt = re.sub(r"^%YAML:1\.0\s*\n", "", txt)
    # cv::FileStorage accepts tabs and '!!opencv-matrix' tags; strip the tags.
    txt = txt.replace("!!opencv-matrix", "")
    return yaml.safe_load(txt) or {}


# AnmsAlgorithmType -- include/kimera-vio/frontend/feature-detector/NonMaximumSuppression.h:52-60
ANMS_TOPN, ANMS_BROWN, ANMS_SDC, ANMS_KDTREE, ANMS_RANGETREE, ANMS_SSC, ANMS_BINNING = range(7)
# FeatureDetectorType -- FeatureDetector-definitions.h
DET_FAST, DET_ORB, DET_AGAST, DET_GFTT = range(4)


@dataclass
class FrontendParams:
    # --- tracker (VisionImuTrackerParams.h:42-70)
    klt_win_size: int = 24
    klt_max_iter: int = 30
    klt_max_level: int = 3
    klt_eps: float = 0.01
    max_feature_track_age: int = 25
    min_nr_mono_inliers: int = 10
    min_nr_stereo_inliers: int = 5
    ransac_threshold_mono: float = 1.0e-6
    ransac_threshold_stereo: float = 1.0
    ransac_max_iterations: int = 100
    ransac_probability: float = 0.995
    ransac_randomize: bool = True
    ransac_use_1point_stereo: bool = True
    ransac_use_2point_mono: bool = True
    pose_2d2d_algorithm: int = 1
    optical_flow_predictor_type: int = 1  # 0 static (NoPredictionOpticalFlowPredictor), 1 rotational
    disparity_threshold: float = 0.5
    # --- detector (FeatureDetectorParams.h:72-106)
    feature_detector_type: int = DET_GFTT
    max_features_per_frame: int = 400
    enable_subpixel_corner_refinement: bool = True
    subpix_max_iters: int = 10
    subpix_epsilon: float = 0.01
    subpix_window_size: int = 10
    subpix_zero_zone: int = -1
    enable_non_max_suppression: bool = True
    non_max_suppression_type: int = ANMS_RANGETREE
    min_distance: int = 10
    max_nr_keypoints_before_anms: int = 2000
    nr_horizontal_bins: int = 5
    nr_vertical_bins: int = 5
    binning_mask: Optional[np.ndarray] = None  # (nr_vertical_bins, nr_horizontal_bins) of {0,1}
    quality_level: float = 0.001
    block_size: int = 3
    use_harris_detector: bool = False
    k: float = 0.04
    # --- stereo matching (StereoMatchingParams.h)
    tolerance_template_matching: float = 0.15
    nominal_baseline: float = 0.11
    templ_cols: int = 101
    templ_rows: int = 11
    stripe_extra_rows: int = 0
    min_point_dist: float = 0.1
    max_point_dist: float = 15.0
    bidirectional_matching: bool = False
    subpixel_refinement_stereo: bool = False
    equalize_image: bool = False
    optimize_2d2d_pose_from_inliers: bool = False     # VisionImuTrackerParams.cpp:119-122 (rejected by kvfe_create when set)
    optimize_3d3d_pose_from_inliers: bool = False
    # --- front-end FSM (VisionImuFrontendParams.h)
    min_intra_keyframe_time_ns: int = int(0.2 * 10e6)   # sic: the reference's units slip (SURVEY B-11)
    max_intra_keyframe_time_ns: int = int(10.0 * 10e6)
    min_number_features: int = 0
    use_stereo_tracking: bool = True
    use_ransac: bool = True
    use_pnp_tracking: bool = False
    max_disparity_since_lkf: float = 200.0

    def __post_init__(self):
        if self.binning_mask is None:
            self.binning_mask = np.ones((self.nr_vertical_bins, self.nr_horizontal_bins), np.float64)

    @staticmethod
    def from_yaml(path: str) -> "FrontendParams":
        y = _load_cv_yaml(path)
        p = FrontendParams()

        def get(key, attr, cast):
            if key in y and y[key] is not None:
                setattr(p, attr, cast(y[key]))

        b = lambda v: bool(int(v))
        get("klt_win_size", "klt_win_size", int)
        get("klt_max_iter", "klt_max_iter", int)
        get("klt_max_level", "klt_max_level", int)
        get("klt_eps", "klt_eps", float)
        get("maxFeatureAge", "max_feature_track_age", int)
        get("minNrMonoInliers", "min_nr_mono_inliers", int)
        get("minNrStereoInliers", "min_nr_stereo_inliers", int)
        get("ransac_threshold_mono", "ransac_threshold_mono", float)
        get("ransac_threshold_stereo", "ransac_threshold_stereo", float)
        get("ransac_max_iterations", "ransac_max_iterations", int)
        get("ransac_probability", "ransac_probability", float)
        get("ransac_randomize", "ransac_randomize", b)
        get("ransac_use_1point_stereo", "ransac_use_1point_stereo", b)
        get("ransac_use_2point_mono", "ransac_use_2point_mono", b)
        get("2d2d_algorithm", "pose_2d2d_algorithm", int)
        get("optical_flow_predictor_type", "optical_flow_predictor_type", int)
        get("disparityThreshold", "disparity_threshold", float)
        get("feature_detector_type", "feature_detector_type", int)
        get("maxFeaturesPerFrame", "max_features_per_frame", int)
        get("enable_subpixel_corner_finder", "enable_subpixel_corner_refinement", b)
        if p.enable_subpixel_corner_refinement:
            get("max_iters", "subpix_max_iters", int)
            get("epsilon_error", "subpix_epsilon", float)
            get("window_size", "subpix_window_size", int)
            get("zero_zone", "subpix_zero_zone", int)
        get("enable_non_max_suppression", "enable_non_max_suppression", b)
        get("non_max_suppression_type", "non_max_suppression_type", int)
        get("min_distance", "min_distance", int)
        get("max_nr_keypoints_before_anms", "max_nr_keypoints_before_anms", int)
        get("nr_horizontal_bins", "nr_horizontal_bins", int)
        get("nr_vertical_bins", "nr_vertical_bins", int)
        p.binning_mask = np.ones((p.nr_vertical_bins, p.nr_horizontal_bins), np.float64)
        vm = y.get("binning_mask") or []
        if len(vm) > 0:
            if len(vm) != p.nr_vertical_bins * p.nr_horizontal_bins:
                raise ValueError("binning_mask size inconsistent with the number of bins")
            vm = np.asarray(vm, np.float64)
            if not np.all((vm == 0) | (vm == 1)):
                raise ValueError("binning_mask can only have binary entries")
            p.binning_mask = vm.reshape(p.nr_vertical_bins, p.nr_horizontal_bins)
        get("quality_level", "quality_level", float)
        get("block_size", "block_size", int)
        get("use_harris_detector", "use_harris_detector", b)
        get("k", "k", float)
        get("toleranceTemplateMatching", "tolerance_template_matching", float)
        get("nominalBaseline", "nominal_baseline", float)
        get("templ_cols", "templ_cols", int)
        get("templ_rows", "templ_rows", int)
        get("stripe_extra_rows", "stripe_extra_rows", int)
        get("minPointDist", "min_point_dist", float)
        get("maxPointDist", "max_point_dist", float)
        get("bidirectionalMatching", "bidirectional_matching", b)
        get("subpixelRefinementStereo", "subpixel_refinement_stereo", b)
        get("equalizeImage", "equalize_image", b)
        get("optimize_2d2d_pose_from_inliers", "optimize_2d2d_pose_from_inliers", b)
        get("optimize_3d3d_pose_from_inliers", "optimize_3d3d_pose_from_inliers", b)
        if "min_intra_keyframe_time" in y:
            p.min_intra_keyframe_time_ns = int(float(y["min_intra_keyframe_time"]) * 1e9)
        if "max_intra_keyframe_time" in y:
            p.max_intra_keyframe_time_ns = int(float(y["max_intra_keyframe_time"]) * 1e9)
        get("minNumberFeatures", "min_number_features", int)
        get("useStereoTracking", "use_stereo_tracking", b)
        get("useRANSAC", "use_ransac", b)
        get("use_pnp_tracking", "use_pnp_tracking", b)
        get("max_disparity_since_lkf", "max_disparity_since_lkf", float)
        return p

    @staticmethod
    def euroc() -> "FrontendParams":
        """params/Euroc/FrontendParams.yaml, inlined so the GPU box (no /root/reference) has it."""
        p = FrontendParams(
            klt_win_size=24, klt_max_iter=30, klt_max_level=4, klt_eps=0.1, max_feature_track_age=25,
            min_nr_mono_inliers=10, min_nr_stereo_inliers=5, ransac_threshold_mono=1e-6,
            ransac_threshold_stereo=1.0, ransac_max_iterations=100, ransac_probability=0.995,
            ransac_randomize=False, ransac_use_1point_stereo=True, ransac_use_2point_mono=True,
            pose_2d2d_algorithm=1, optical_flow_predictor_type=1, disparity_threshold=0.5,
            feature_detector_type=DET_GFTT, max_features_per_frame=300,
            enable_subpixel_corner_refinement=True, subpix_max_iters=40, subpix_epsilon=0.001,
            subpix_window_size=10, subpix_zero_zone=-1, enable_non_max_suppression=True,
            non_max_suppression_type=ANMS_BINNING, min_distance=20, max_nr_keypoints_before_anms=2000,
            nr_horizontal_bins=7, nr_vertical_bins=5, quality_level=0.001, block_size=3,
            use_harris_detector=False, k=0.04, tolerance_template_matching=0.15, nominal_baseline=0.11,
            templ_cols=101, templ_rows=11, stripe_extra_rows=0, min_point_dist=0.5, max_point_dist=10.0,
            min_intra_keyframe_time_ns=int(0.2 * 1e9), max_intra_keyframe_time_ns=int(5.0 * 1e9),
            min_number_features=0, use_stereo_tracking=True, use_ransac=True, use_pnp_tracking=False,
            max_disparity_since_lkf=1000.0)
        return p


@dataclass
class CameraParams:
    """Pinhole camera with radial-tangential or equidistant distortion (CameraParams.cpp:30-140). T_BS is body_Pose_cam.
    `depth`: the RGB-D block (CameraParams.cpp:342-349, parsed when the YAML carries `virtual_baseline`): a dict with
    virtual_baseline, depth_to_meters, min_depth, max_depth (float32 values, like the reference's struct) and is_registered."""
    width: int
    height: int
    intrinsics: List[float]            # fu, fv, cu, cv
    distortion: List[float]            # k1, k2, p1, p2 (radtan) / k1..k4 (equidistant)
    T_BS: np.ndarray = field(default_factory=lambda: np.eye(4))
    distortion_model: str = "radtan"
    depth: Optional[dict] = None

    @property
    def K(self) -> np.ndarray:
        fu, fv, cu, cv_ = self.intrinsics
        return np.array([[fu, 0.0, cu], [0.0, fv, cv_], [0.0, 0.0, 1.0]], np.float64)

    @property
    def D(self) -> np.ndarray:
        return np.asarray(self.distortion, np.float64).reshape(1, -1)

    @staticmethod
    def from_yaml(path: str) -> "CameraParams":
        y = _load_cv_yaml(path)
        res = y["resolution"]
        T = np.asarray(y["T_BS"]["data"], np.float64).reshape(4, 4)
        dm = str(y.get("distortion_model", "radtan")).lower()
        if dm in ("radtan", "radial-tangential", "plumb_bob"):
            dm = "radtan"
        elif dm in ("none",):
            dm = "none"
        depth = None
        if "virtual_baseline" in y:
            depth = {"virtual_baseline": float(np.float32(y["virtual_baseline"])),
                     "depth_to_meters": float(np.float32(y.get("depth_to_meters", 1.0))),
                     "min_depth": float(np.float32(y.get("min_depth", 0.0))),
                     "max_depth": float(np.float32(y.get("max_depth", 10.0))),
                     "is_registered": bool(int(y.get("is_registered", 1)))}
            if not depth["virtual_baseline"] > 0:
                raise ValueError("Baseline must be positive")            # CHECK_GT, CameraParams.cpp:344
        return CameraParams(int(res[0]), int(res[1]), [float(v) for v in y["intrinsics"]],
                            [float(v) for v in y["distortion_coefficients"]], T, dm, depth)

    @staticmethod
    def euroc_left() -> "CameraParams":
        T = np.array([0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975,
                      0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768,
                      -0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949,
                      0.0, 0.0, 0.0, 1.0]).reshape(4, 4)
        return CameraParams(752, 480, [458.654, 457.296, 367.215, 248.375],
                            [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], T)

    @staticmethod
    def euroc_right() -> "CameraParams":
        T = np.array([0.0125552670891, -0.999755099723, 0.0182237714554, -0.0198435579556,
                      0.999598781151, 0.0130119051815, 0.0251588363115, 0.0453689425024,
                      -0.0253898008918, 0.0179005838253, 0.999517347078, 0.00786212447038,
                      0.0, 0.0, 0.0, 1.0]).reshape(4, 4)
        return CameraParams(752, 480, [457.587, 456.134, 379.999, 255.238],
                            [-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05], T)

    def scaled(self, width: int, height: int) -> "CameraParams":
        """Same lens, different sensor resolution (used for the 720p/1080p/4K synthetic configs)."""
        sx, sy = width / self.width, height / self.height
        fu, fv, cu, cv_ = self.intrinsics
        return dataclasses.replace(self, width=width, height=height,
                                   intrinsics=[fu * sx, fv * sy, cu * sx, cv_ * sy])
