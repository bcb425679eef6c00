// kvfe_shim.hpp -- C++17 mirror of the reference's class boundary over the C-ABI (kvfe.h).
//
// One class per reference class on the hot path, one method per replaced member function, same
// names and argument meaning; images are plain (pointer, width, height, pitch) views and keypoints
// plain vectors so that the header compiles without OpenCV.  Inside Kimera-VIO the cv::Mat /
// KeypointsCV / gtsam glue of INTEGRATION.md sits on top of exactly these calls.  Error behaviour:
// the reference aborts through glog CHECK / LOG(FATAL); here every non-zero status becomes a
// kvfe::Error carrying kvfe_last_error().  There is no CPU fallback: constructing a Context
// without a CUDA device throws.
#ifndef KVFE_SHIM_HPP_
#define KVFE_SHIM_HPP_
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "kvfe.h"

namespace kvfe {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

struct ImageView {            // cv::Mat CV_8UC1
  const uint8_t* data; int width, height; size_t pitch;
};
struct MutableImage {
  uint8_t* data; int width, height; size_t pitch;
};
struct Keypoints { std::vector<float> x, y; size_t size() const { return x.size(); } };

// Owns the kvfe_ctx (one camera rig; stage-level calls are synchronous).
class Context {
 public:
  Context(const kvfe_config& cfg, const kvfe_rig& rig) {
    kvfe_ctx* c = nullptr;
    const int rc = kvfe_create(&cfg, &rig, &c);
    if (rc != KVFE_OK) throw Error(rc, std::string("kvfe_create: ") + kvfe_last_error(nullptr));
    ctx_.reset(c, kvfe_destroy);
  }
  kvfe_ctx* get() const { return ctx_.get(); }
  void check(int rc, const char* where) const {
    if (rc != KVFE_OK) throw Error(rc, std::string(where) + ": " + kvfe_last_error(ctx_.get()));
  }
 private:
  std::shared_ptr<kvfe_ctx> ctx_;
};

// UndistorterRectifier (src/frontend/UndistorterRectifier.cpp) + StereoCamera::undistortRectifyStereoFrame
class UndistorterRectifier {
 public:
  explicit UndistorterRectifier(Context ctx) : c_(std::move(ctx)) {}
  // StereoCamera.cpp:269-290: both images of a stereo frame
  void undistortRectifyStereoFrame(const ImageView& left, const ImageView& right, MutableImage* left_rect,
                                   MutableImage* right_rect) const {
    c_.check(kvfe_rectify_pair(c_.get(), left.data, right.data, left.pitch, left_rect->data, right_rect->data,
                               left_rect->pitch), "undistortRectifyStereoFrame");
  }
  // UndistorterRectifier.cpp:33-68 (cam 0 = left, 1 = right)
  Keypoints UndistortRectifyKeypoints(int cam, const Keypoints& kps, bool use_R = true, bool use_P = true) const {
    Keypoints out; out.x.resize(kps.size()); out.y.resize(kps.size());
    c_.check(kvfe_undistort_keypoints(c_.get(), cam, use_R, use_P, kps.x.data(), kps.y.data(), (int)kps.size(),
                                      out.x.data(), out.y.data()), "UndistortRectifyKeypoints");
    return out;
  }
  // StereoCamera::undistortRectifyLeftKeypoints (StereoCamera.cpp:236-260): statuses + rectified keypoints
  std::vector<int32_t> undistortRectifyLeftKeypoints(const Keypoints& kps, Keypoints* rectified) const {
    std::vector<int32_t> st(kps.size());
    rectified->x.resize(kps.size()); rectified->y.resize(kps.size());
    c_.check(kvfe_undistort_rectify_left_keypoints(c_.get(), kps.x.data(), kps.y.data(), (int)kps.size(), st.data(),
                                                   rectified->x.data(), rectified->y.data()), "undistortRectifyLeftKeypoints");
    return st;
  }
  // UndistorterRectifier.cpp:138-211
  std::vector<int32_t> checkUndistortedRectifiedLeftKeypoints(const Keypoints& distorted, const Keypoints& undistorted,
                                                              Keypoints* status_kps, float pixel_tol = 2.0f, int cam = 0) const {
    std::vector<int32_t> st(distorted.size());
    status_kps->x.resize(distorted.size()); status_kps->y.resize(distorted.size());
    c_.check(kvfe_check_rectified_keypoints(c_.get(), cam, distorted.x.data(), distorted.y.data(), undistorted.x.data(),
                                            undistorted.y.data(), (int)distorted.size(), pixel_tol, st.data(), status_kps->x.data(),
                                            status_kps->y.data()), "checkUndistortedRectifiedLeftKeypoints");
    return st;
  }
  // UndistorterRectifier.cpp:213-228 / StereoCamera::distortUnrectifyRightKeypoints (cam = 1)
  Keypoints distortUnrectifyKeypoints(int cam, const std::vector<int32_t>& status, const Keypoints& rectified) const {
    Keypoints out; out.x.resize(status.size()); out.y.resize(status.size());
    c_.check(kvfe_distort_unrectify_keypoints(c_.get(), cam, status.data(), rectified.x.data(), rectified.y.data(), (int)status.size(),
                                              out.x.data(), out.y.data()), "distortUnrectifyKeypoints");
    return out;
  }
  // UndistorterRectifier.cpp:73-113: unit bearing vectors (3 doubles per keypoint)
  std::vector<double> GetBearingVectors(const Keypoints& kps) const {
    std::vector<double> v(3 * kps.size());
    c_.check(kvfe_bearing_vectors(c_.get(), kps.x.data(), kps.y.data(), (int)kps.size(), v.data()), "GetBearingVector");
    return v;
  }
 private:
  Context c_;
};

// FeatureDetector (src/frontend/feature-detector/FeatureDetector.cpp)
class FeatureDetector {
 public:
  explicit FeatureDetector(Context ctx) : c_(std::move(ctx)) {}
  // featureDetection(const Frame&, need_n_corners) :174-299 -- the frame's current keypoints / landmarks
  // (landmark -1 = invalidated, not masked) in, the new sub-pixel corners out
  Keypoints featureDetection(const ImageView& img, const Keypoints& existing, const std::vector<int64_t>& landmarks,
                             int need_n_corners) const {
    const int cap = kvfe_max_keypoints(c_.get());
    Keypoints out; out.x.resize(cap); out.y.resize(cap);
    int n = 0;
    c_.check(kvfe_detect(c_.get(), img.data, img.pitch, existing.x.data(), existing.y.data(), landmarks.data(),
                         (int)existing.size(), need_n_corners, out.x.data(), out.y.data(), &n), "featureDetection");
    out.x.resize(n); out.y.resize(n);
    return out;
  }
  // the same with the frame's detection_mask_ (FeatureDetector.cpp:186-189)
  Keypoints featureDetection(const ImageView& img, const ImageView& detection_mask, const Keypoints& existing,
                             const std::vector<int64_t>& landmarks, int need_n_corners) const {
    const int cap = kvfe_max_keypoints(c_.get());
    Keypoints out; out.x.resize(cap); out.y.resize(cap);
    int n = 0;
    c_.check(kvfe_detect_masked(c_.get(), img.data, img.pitch, detection_mask.data, detection_mask.pitch, existing.x.data(),
                                existing.y.data(), landmarks.data(), (int)existing.size(), need_n_corners, out.x.data(), out.y.data(), &n),
             "featureDetection(mask)");
    out.x.resize(n); out.y.resize(n);
    return out;
  }
 private:
  Context c_;
};

struct TrackingResult { Keypoints predicted, tracked; std::vector<uint8_t> status; };
struct RansacResult { int status; std::vector<int32_t> inliers; double pose[12]; double info[9]; };

// Tracker (src/frontend/Tracker.cpp)
class Tracker {
 public:
  explicit Tracker(Context ctx) : c_(std::move(ctx)) {}
  // the optical-flow half of featureTracking :117-148 (prediction with the IMU rotation + pyramidal LK);
  // the landmark / age bookkeeping of :150-199 stays in the caller as in INTEGRATION.md
  TrackingResult featureTracking(const ImageView& ref, const ImageView& cur, const Keypoints& ref_kps,
                                 const double ref_R_cur[9]) const {
    const size_t n = ref_kps.size();
    TrackingResult r;
    r.predicted.x.resize(n); r.predicted.y.resize(n); r.tracked.x.resize(n); r.tracked.y.resize(n); r.status.resize(n);
    c_.check(kvfe_track(c_.get(), ref.data, cur.data, ref.pitch, ref_R_cur, ref_kps.x.data(), ref_kps.y.data(), (int)n,
                        r.predicted.x.data(), r.predicted.y.data(), r.tracked.x.data(), r.tracked.y.data(),
                        r.status.data()), "featureTracking");
    return r;
  }
  // geometricOutlierRejection2d2d :213-319 -- R12 given: 2-point; nullptr: 5-point Nister
  RansacResult geometricOutlierRejection2d2d(const std::vector<double>& f_ref, const std::vector<double>& f_cur,
                                             const double* R12) const {
    const int n = (int)(f_ref.size() / 3);
    RansacResult r{}; r.inliers.resize(n);
    int ni = 0;
    c_.check(kvfe_ransac_mono(c_.get(), f_ref.data(), f_cur.data(), n, R12, r.inliers.data(), &ni, r.pose, &r.status),
             "geometricOutlierRejection2d2d");
    r.inliers.resize(ni);
    return r;
  }
  // geometricOutlierRejection3d3d :667-742 (3-point Arun)
  RansacResult geometricOutlierRejection3d3d(const std::vector<double>& ref_3d, const std::vector<double>& cur_3d) const {
    const int n = (int)(ref_3d.size() / 3);
    RansacResult r{}; r.inliers.resize(n);
    int ni = 0;
    c_.check(kvfe_ransac_stereo_3pt(c_.get(), ref_3d.data(), cur_3d.data(), n, r.inliers.data(), &ni, r.pose, &r.status),
             "geometricOutlierRejection3d3d");
    r.inliers.resize(ni);
    return r;
  }
  // geometricOutlierRejection3d3dGivenRotation :382-632 (1-point voting); *_xy: 2 floats per match
  RansacResult geometricOutlierRejection3d3dGivenRotation(const std::vector<float>& ref_left_xy, const std::vector<float>& ref_right_xy,
                                                          const std::vector<float>& cur_left_xy, const std::vector<float>& cur_right_xy,
                                                          const std::vector<double>& ref_3d, const std::vector<double>& cur_3d,
                                                          const double R[9]) const {
    const int n = (int)(ref_3d.size() / 3);
    RansacResult r{}; r.inliers.resize(n);
    int ni = 0;
    c_.check(kvfe_ransac_stereo_1pt(c_.get(), ref_left_xy.data(), ref_right_xy.data(), cur_left_xy.data(), cur_right_xy.data(),
                                    ref_3d.data(), cur_3d.data(), n, R, r.inliers.data(), &ni, r.pose, r.info, &r.status),
             "geometricOutlierRejection3d3dGivenRotation");
    r.inliers.resize(ni);
    return r;
  }
  // computeMedianDisparity :991-1018 (matches: ref index, cur index pairs); false when there is no match
  bool computeMedianDisparity(const Keypoints& ref, const Keypoints& cur, const std::vector<int32_t>& match_ref,
                              const std::vector<int32_t>& match_cur, double* median) const {
    int ok = 0;
    c_.check(kvfe_compute_median_disparity(c_.get(), ref.x.data(), ref.y.data(), (int)ref.size(), cur.x.data(), cur.y.data(),
                                           (int)cur.size(), match_ref.data(), match_cur.data(), (int)match_ref.size(), median, &ok),
             "computeMedianDisparity");
    return ok != 0;
  }
  // getPoint3AndCovariance :772-818 for every rectified stereo point (stereo_point_covariance = identity)
  void getPoint3AndCovariance(const Keypoints& left_rect, const Keypoints& right_rect, const std::vector<double>& points_3d,
                              const double* Rmat, std::vector<double>* points, std::vector<double>* covariances) const {
    const size_t n = left_rect.size();
    points->resize(3 * n); covariances->resize(9 * n);
    c_.check(kvfe_point3_and_covariance(c_.get(), left_rect.x.data(), right_rect.x.data(), left_rect.y.data(), points_3d.data(), (int)n,
                                        Rmat, points->data(), covariances->data()), "getPoint3AndCovariance");
  }
  // findMatchingKeypoints :919-946 (landmark vectors of the two frames -> index pairs)
  static void findMatchingKeypoints(const std::vector<int64_t>& ref_landmarks, const std::vector<int64_t>& cur_landmarks,
                                    std::vector<int32_t>* match_ref, std::vector<int32_t>* match_cur) {
    match_ref->assign(cur_landmarks.size() + 1, 0); match_cur->assign(cur_landmarks.size() + 1, 0);
    int nm = 0;
    if (kvfe_find_matching_keypoints(ref_landmarks.data(), (int)ref_landmarks.size(), cur_landmarks.data(), (int)cur_landmarks.size(),
                                     match_ref->data(), match_cur->data(), &nm) != KVFE_OK) throw Error(KVFE_ERR_INVALID_ARG, "findMatchingKeypoints");
    match_ref->resize(nm); match_cur->resize(nm);
  }
  // findMatchingStereoKeypoints :948-989 (keeps the mono matches whose right keypoints are VALID in both frames)
  static void findMatchingStereoKeypoints(const std::vector<int32_t>& ref_right_status, const std::vector<int32_t>& cur_right_status,
                                          std::vector<int32_t>* match_ref, std::vector<int32_t>* match_cur) {
    int nm = 0;
    if (kvfe_find_matching_stereo_keypoints(ref_right_status.data(), (int)ref_right_status.size(), cur_right_status.data(),
                                            (int)cur_right_status.size(), match_ref->data(), match_cur->data(), (int)match_ref->size(),
                                            match_ref->data(), match_cur->data(), &nm) != KVFE_OK)
      throw Error(KVFE_ERR_INVALID_ARG, "findMatchingStereoKeypoints");
    match_ref->resize(nm); match_cur->resize(nm);
  }
  // findOutliers :836-853
  static std::vector<int32_t> findOutliers(int n_matches, const std::vector<int32_t>& inliers) {
    std::vector<int32_t> out(n_matches > 0 ? n_matches : 1);
    int no = 0;
    if (kvfe_find_outliers(n_matches, inliers.data(), (int)inliers.size(), out.data(), &no) != KVFE_OK) throw Error(KVFE_ERR_INVALID_ARG, "findOutliers");
    out.resize(no);
    return out;
  }
  // removeOutliersMono :856-882
  static void removeOutliersMono(const std::vector<int32_t>& inliers, std::vector<int64_t>* ref_landmarks, std::vector<int64_t>* cur_landmarks,
                                 std::vector<int32_t>* match_ref, std::vector<int32_t>* match_cur) {
    int nm = (int)match_ref->size();
    if (kvfe_remove_outliers_mono(inliers.data(), (int)inliers.size(), ref_landmarks->data(), (int)ref_landmarks->size(), cur_landmarks->data(),
                                  (int)cur_landmarks->size(), match_ref->data(), match_cur->data(), &nm) != KVFE_OK)
      throw Error(KVFE_ERR_INVALID_ARG, "removeOutliersMono");
    match_ref->resize(nm); match_cur->resize(nm);
  }
 private:
  Context c_;
};

struct SparseStereoResult {
  std::vector<int32_t> left_status, right_status;
  Keypoints left_rectified, right_rectified, right_keypoints;
  std::vector<double> depth, points_3d;
};

// StereoMatcher (src/frontend/StereoMatcher.cpp)
class StereoMatcher {
 public:
  explicit StereoMatcher(Context ctx) : c_(std::move(ctx)) {}
  // sparseStereoReconstruction(StereoFrame*) :123-175; left_rect / right_rect may be null
  SparseStereoResult sparseStereoReconstruction(const ImageView& left, const ImageView& right, const Keypoints& left_kps,
                                                const std::vector<double>& versors, MutableImage* left_rect = nullptr,
                                                MutableImage* right_rect = nullptr) const {
    const size_t n = left_kps.size();
    SparseStereoResult r;
    r.left_status.resize(n); r.right_status.resize(n); r.depth.resize(n); r.points_3d.resize(3 * n);
    for (Keypoints* k : {&r.left_rectified, &r.right_rectified, &r.right_keypoints}) { k->x.resize(n); k->y.resize(n); }
    kvfe_stereo_out o{r.left_status.data(), r.left_rectified.x.data(), r.left_rectified.y.data(), r.right_status.data(),
                      r.right_rectified.x.data(), r.right_rectified.y.data(), r.depth.data(), r.points_3d.data(),
                      r.right_keypoints.x.data(), r.right_keypoints.y.data()};
    c_.check(kvfe_sparse_stereo(c_.get(), left.data, right.data, left.pitch, left_kps.x.data(), left_kps.y.data(), versors.data(),
                                (int)n, &o, left_rect ? left_rect->data : nullptr, right_rect ? right_rect->data : nullptr,
                                left_rect ? left_rect->pitch : 0), "sparseStereoReconstruction");
    return r;
  }
  // getRightKeypointsRectified (StereoMatcher.cpp:196-281) on an already rectified pair
  std::vector<int32_t> getRightKeypointsRectified(const ImageView& left_rectified, const ImageView& right_rectified,
                                                  const std::vector<int32_t>& left_status, const Keypoints& left_rect,
                                                  Keypoints* right_rect) const {
    std::vector<int32_t> st(left_status.size());
    right_rect->x.resize(st.size()); right_rect->y.resize(st.size());
    c_.check(kvfe_right_keypoints_rectified(c_.get(), left_rectified.data, right_rectified.data, left_rectified.pitch, left_status.data(),
                                            left_rect.x.data(), left_rect.y.data(), (int)st.size(), st.data(), right_rect->x.data(),
                                            right_rect->y.data()), "getRightKeypointsRectified");
    return st;
  }
  // getDepthFromRectifiedMatches (StereoMatcher.cpp:425-483); right_status is updated like the reference's list
  std::vector<double> getDepthFromRectifiedMatches(const std::vector<int32_t>& left_status, const Keypoints& left_rect,
                                                   std::vector<int32_t>* right_status, const Keypoints& right_rect) const {
    std::vector<double> depth(left_status.size());
    c_.check(kvfe_depth_from_rectified_matches(c_.get(), left_status.data(), left_rect.x.data(), right_status->data(), right_rect.x.data(),
                                               (int)left_status.size(), depth.data()), "getDepthFromRectifiedMatches");
    return depth;
  }
 private:
  Context c_;
};

// The RGB-D additions (src/frontend/DepthFrame.cpp, src/frontend/RgbdFrame.cpp): a registered depth image of the context's
// size, CV_16UC1 or CV_32FC1 as kvfe_depth_params.depth_type says; pitch in bytes
struct DepthView { const void* data; size_t pitch_bytes; };
struct RgbdFillResult {
  std::vector<int32_t> right_status;
  Keypoints right_rectified, right_keypoints;
  std::vector<double> depth, points_3d;
};
class RgbdFrame {
 public:
  RgbdFrame(Context ctx, const kvfe_depth_params& dp) : c_(std::move(ctx)), dp_(dp) {}
  // DepthFrame::getDetectionMask (DepthFrame.cpp:75-98) -> Frame::detection_mask_ for FeatureDetector::featureDetection
  void getDetectionMask(const DepthView& depth, MutableImage* mask) const {
    c_.check(kvfe_depth_detection_mask(c_.get(), depth.data, depth.pitch_bytes, &dp_, mask->data, mask->pitch), "getDetectionMask");
  }
  // RgbdFrame::fillStereoFrame (RgbdFrame.cpp:52-115): the hallucinated right half of the StereoFrame
  RgbdFillResult fillStereoFrame(const DepthView& depth, const Keypoints& left_kps, const std::vector<int32_t>& left_status,
                                 const Keypoints& left_rect, const std::vector<double>& versors) const {
    const size_t n = left_kps.size();
    RgbdFillResult r;
    r.right_status.resize(n); r.depth.resize(n); r.points_3d.resize(3 * n);
    for (Keypoints* k : {&r.right_rectified, &r.right_keypoints}) { k->x.resize(n); k->y.resize(n); }
    c_.check(kvfe_rgbd_fill_stereo_frame(c_.get(), depth.data, depth.pitch_bytes, &dp_, left_kps.x.data(), left_kps.y.data(),
                                         left_status.data(), left_rect.x.data(), left_rect.y.data(), versors.data(), (int)n,
                                         r.right_status.data(), r.right_rectified.x.data(), r.right_rectified.y.data(), r.depth.data(),
                                         r.points_3d.data(), r.right_keypoints.x.data(), r.right_keypoints.y.data()), "fillStereoFrame");
    return r;
  }
 private:
  Context c_;
  kvfe_depth_params dp_;
};

}  // namespace kvfe
#endif  // KVFE_SHIM_HPP_
