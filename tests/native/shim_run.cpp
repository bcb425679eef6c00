// Test harness (GPU): drives EVERY method of include/kvfe_shim.hpp with real data and dumps the results, so that
// tests/test_gpu_shim.py can compare them with the same calls made through ctypes.  Built on the fly (g++), links
// libkvfe.so.  Input file: [kvfe_config][kvfe_rig][W*H left][W*H right][W*H left2]; output: named blobs.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "kvfe_shim.hpp"

static std::ofstream out;
template <typename T>
static void dump(const char* name, const std::vector<T>& v) {
  const unsigned nl = (unsigned)strlen(name), es = (unsigned)sizeof(T);
  const unsigned long long n = v.size();
  out.write((const char*)&nl, 4); out.write(name, nl); out.write((const char*)&es, 4); out.write((const char*)&n, 8);
  out.write((const char*)v.data(), (std::streamsize)(n * es));
}
static void dump(const char* name, const kvfe::Keypoints& k) {
  dump((std::string(name) + ".x").c_str(), k.x);
  dump((std::string(name) + ".y").c_str(), k.y);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  kvfe_config cfg; kvfe_rig rig;
  in.read((char*)&cfg, sizeof(cfg)); in.read((char*)&rig, sizeof(rig));
  const int W = cfg.width, H = cfg.height;
  std::vector<uint8_t> L((size_t)W * H), R(L.size()), L2(L.size());
  in.read((char*)L.data(), L.size()); in.read((char*)R.data(), R.size()); in.read((char*)L2.data(), L2.size());
  if (!in) return 4;
  out.open(argv[2], std::ios::binary);
  try {
    kvfe::Context c(cfg, rig);
    kvfe::UndistorterRectifier ur(c); kvfe::FeatureDetector fd(c); kvfe::Tracker tr(c); kvfe::StereoMatcher sm(c);
    const kvfe::ImageView vl{L.data(), W, H, (size_t)W}, vr{R.data(), W, H, (size_t)W}, vl2{L2.data(), W, H, (size_t)W};
    // rectification
    std::vector<uint8_t> Lr(L.size()), Rr(L.size());
    kvfe::MutableImage ml{Lr.data(), W, H, (size_t)W}, mr{Rr.data(), W, H, (size_t)W};
    ur.undistortRectifyStereoFrame(vl, vr, &ml, &mr);
    dump("rect_left", Lr); dump("rect_right", Rr);
    // detection (plain and masked)
    kvfe::Keypoints none;
    std::vector<int64_t> nolmk;
    kvfe::Keypoints kps = fd.featureDetection(vl, none, nolmk, 150);
    dump("detect", kps);
    std::vector<uint8_t> mask(L.size(), 255);
    for (int y = 0; y < H; ++y) for (int x = W / 3; x < W / 2; ++x) mask[(size_t)y * W + x] = 0;
    const kvfe::ImageView vm{mask.data(), W, H, (size_t)W};
    dump("detect_masked", fd.featureDetection(vl, vm, none, nolmk, 150));
    // keypoint geometry
    dump("undistort", ur.UndistortRectifyKeypoints(0, kps));
    std::vector<double> versors = ur.GetBearingVectors(kps);
    dump("versors", versors);
    kvfe::Keypoints lrect;
    std::vector<int32_t> lstat = ur.undistortRectifyLeftKeypoints(kps, &lrect);
    dump("left_status", lstat); dump("left_rect", lrect);
    kvfe::Keypoints checked;
    dump("check_status", ur.checkUndistortedRectifiedLeftKeypoints(kps, ur.UndistortRectifyKeypoints(0, kps), &checked, 1.0f));
    dump("check_kps", checked);
    // tracking
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    kvfe::TrackingResult t = tr.featureTracking(vl, vl2, kps, I3);
    dump("track", t.tracked); dump("track_status", t.status);
    // sparse stereo, whole and in pieces
    kvfe::SparseStereoResult s = sm.sparseStereoReconstruction(vl, vr, kps, versors);
    dump("ss_left_status", s.left_status); dump("ss_right_status", s.right_status); dump("ss_right_rect", s.right_rectified);
    dump("ss_depth", s.depth); dump("ss_points", s.points_3d); dump("ss_right_kps", s.right_keypoints);
    const kvfe::ImageView vlr{Lr.data(), W, H, (size_t)W}, vrr{Rr.data(), W, H, (size_t)W};
    kvfe::Keypoints rrect;
    std::vector<int32_t> rstat = sm.getRightKeypointsRectified(vlr, vrr, lstat, lrect, &rrect);
    dump("right_status", rstat); dump("right_rect", rrect);
    std::vector<int32_t> rstat2 = rstat;
    dump("depth", sm.getDepthFromRectifiedMatches(lstat, lrect, &rstat2, rrect));
    dump("right_status_after_depth", rstat2);
    dump("right_unrect", ur.distortUnrectifyKeypoints(1, rstat2, rrect));
    // tracker statics
    std::vector<int32_t> mref, mcur;
    for (size_t i = 0; i < kps.size(); ++i) if (t.status[i]) { mref.push_back((int32_t)i); mcur.push_back((int32_t)i); }
    double med = 0.0;
    const bool ok = tr.computeMedianDisparity(kps, t.tracked, mref, mcur, &med);
    dump("median", std::vector<double>{ok ? 1.0 : 0.0, med});
    std::vector<double> pts, covs;
    tr.getPoint3AndCovariance(s.left_rectified, s.right_rectified, s.points_3d, I3, &pts, &covs);
    dump("p3", pts); dump("cov", covs);
    // RANSAC on the tracked bearings / stereo points
    std::vector<double> fr, fc;
    std::vector<double> v2 = ur.GetBearingVectors(t.tracked);
    for (size_t i = 0; i < mref.size(); ++i) for (int k = 0; k < 3; ++k) { fr.push_back(versors[3 * mref[i] + k]); fc.push_back(v2[3 * mcur[i] + k]); }
    kvfe::RansacResult r2 = tr.geometricOutlierRejection2d2d(fr, fc, I3);
    dump("ransac2_status", std::vector<int32_t>{r2.status}); dump("ransac2_inliers", r2.inliers);
    kvfe::RansacResult r5 = tr.geometricOutlierRejection2d2d(fr, fc, nullptr);
    dump("ransac5_status", std::vector<int32_t>{r5.status}); dump("ransac5_inliers", r5.inliers);
    std::vector<double> p3v;
    std::vector<float> lxy, rxy;
    for (size_t i = 0; i < kps.size(); ++i) if (s.right_status[i] == 0) {
      for (int k = 0; k < 3; ++k) p3v.push_back(s.points_3d[3 * i + k]);
      lxy.push_back(s.left_rectified.x[i]); lxy.push_back(s.left_rectified.y[i]);
      rxy.push_back(s.right_rectified.x[i]); rxy.push_back(s.right_rectified.y[i]);
    }
    kvfe::RansacResult r3 = tr.geometricOutlierRejection3d3d(p3v, p3v);
    dump("ransac3_status", std::vector<int32_t>{r3.status}); dump("ransac3_inliers", r3.inliers);
    kvfe::RansacResult r1 = tr.geometricOutlierRejection3d3dGivenRotation(lxy, rxy, lxy, rxy, p3v, p3v, I3);
    dump("ransac1_status", std::vector<int32_t>{r1.status}); dump("ransac1_inliers", r1.inliers);
    // RGB-D additions: a synthetic CV_16UC1 depth image in millimetres, derived from the left image
    std::vector<uint16_t> depth16(L.size());
    for (size_t i = 0; i < L.size(); ++i) depth16[i] = (uint16_t)(500 + 20 * (int)L[i]);
    const kvfe_depth_params dpar{KVFE_DEPTH_U16, 0.1f, 0.001f, 2.0f, 4.0f};
    kvfe::RgbdFrame rg(c, dpar);
    const kvfe::DepthView dv{depth16.data(), (size_t)W * 2};
    std::vector<uint8_t> dmask(L.size());
    kvfe::MutableImage mm{dmask.data(), W, H, (size_t)W};
    rg.getDetectionMask(dv, &mm);
    dump("depth_mask", dmask);
    kvfe::RgbdFillResult fill = rg.fillStereoFrame(dv, kps, lstat, lrect, versors);
    dump("fill_right_status", fill.right_status); dump("fill_right_rect", fill.right_rectified); dump("fill_depth", fill.depth);
    dump("fill_points", fill.points_3d); dump("fill_right_kps", fill.right_keypoints);
    // host bookkeeping
    dump("outliers", kvfe::Tracker::findOutliers((int)mref.size(), r2.inliers));
    std::vector<int64_t> lr(kps.size()), lc(kps.size());
    for (size_t i = 0; i < kps.size(); ++i) { lr[i] = (int64_t)i; lc[i] = (int64_t)i; }
    std::vector<int32_t> fm_ref, fm_cur;
    std::vector<int64_t> lc2 = lc;
    for (size_t i = 0; i < lc2.size(); i += 3) lc2[i] = -1;             // every third keypoint lost its landmark
    kvfe::Tracker::findMatchingKeypoints(lr, lc2, &fm_ref, &fm_cur);
    dump("find_matches_ref", fm_ref);
    kvfe::Tracker::findMatchingStereoKeypoints(s.right_status, s.right_status, &fm_ref, &fm_cur);
    dump("find_stereo_matches_ref", fm_ref); dump("find_stereo_matches_cur", fm_cur);
    kvfe::Tracker::removeOutliersMono(r2.inliers, &lr, &lc, &mref, &mcur);
    dump("lmk_ref_after", lr); dump("matches_after", mref);
  } catch (const kvfe::Error& e) {
    std::printf("Error %d: %s\n", e.code, e.what());
    return 3;
  }
  return 0;
}
