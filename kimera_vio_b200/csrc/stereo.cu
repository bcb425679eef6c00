// stereo.cu -- rows a7, a10, a11: the sparse half of StereoMatcher::sparseStereoReconstruction
// (reference src/frontend/StereoMatcher.cpp:123-175).
//   left_rect_kernel   StereoCamera::undistortRectifyLeftKeypoints (StereoCamera.cpp:236-260):
//                      cv::undistortPoints(K, D, R1, P1) + checkUndistortedRectifiedLeftKeypoints
//                      (UndistorterRectifier.cpp:138-211); the map look-up is recomputed in registers.
//   match_kernel       getRightKeypointsRectified / searchRightKeypointEpipolar (StereoMatcher.cpp:
//                      196-423): one CTA per keypoint; template and stripe staged in shared memory;
//                      exact int32 TM_SQDIFF for every shift, first arg-min (cv::minMaxLoc rule).
//   depth_kernel       getDepthFromRectifiedMatches (:425-483), distortUnrectifyRightKeypoints
//                      (UndistorterRectifier.cpp:213-228, right maps recomputed), keypoints_3d (:157-174).
#include "common.cuh"
#include "subpix.cuh"

__global__ void __launch_bounds__(128) left_rect_kernel(DevCfg dc, DevBuf db, const CamModel* __restrict__ cams,
                                                        int mode_mask) {
  const int b = blockIdx.y;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= db.fr.n[fs]) return;
  const size_t k = (size_t)fs * dc.cap + i;
  const float dx = db.fr.kx[k], dy = db.fr.ky[k];
  float ux, uy;
  undistort_point(cams[0], dx, dy, 2, &ux, &uy);
  // UtilsOpenCV::cropToSize
  bool cropped = false;
  const float mw = (float)(dc.W - 1), mh = (float)(dc.H - 1);
  if (ux > mw) { ux = mw; cropped = true; } else if (ux < 0.0f) { ux = 0.0f; cropped = true; }
  if (uy > mh) { uy = mh; cropped = true; } else if (uy < 0.0f) { uy = 0.0f; cropped = true; }
  // NaN coordinates fail every comparison above; keep them in range for the map look-up
  int rx = clampi((int)roundf(ux), 0, dc.W - 1), ry = clampi((int)roundf(uy), 0, dc.H - 1);
  float ex, ey;
  rect_map_at(cams[0], rx, ry, &ex, &ey);
  int status = KVFE_KP_VALID;
  if (cropped) status = KVFE_KP_NO_LEFT_RECT;
  else if (fabsf(dx - ex) > 2.0f || fabsf(dy - ey) > 2.0f) status = KVFE_KP_NO_LEFT_RECT;
  db.fr.lstat[k] = status; db.fr.lrx[k] = ux; db.fr.lry[k] = uy;
}

#define MATCH_THREADS 128

// grid (cap, B); dynamic smem: templ (tc*tr) + stripe (sc*sr) bytes + scores (sc - tc + 1) ints
__global__ void __launch_bounds__(MATCH_THREADS) match_kernel(DevCfg dc, DevBuf db, int mode_mask, int reuse_tracked) {
  extern __shared__ unsigned char smraw[];
  const int b = blockIdx.y;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int i = blockIdx.x;
  if (i >= db.fr.n[fs]) return;
  const size_t k = (size_t)fs * dc.cap + i;
  const int lst = db.fr.lstat[k];
  // Second reconstruction of a keyframe (StereoVisionImuFrontend.cpp:426): the tracked keypoints
  // were already matched in the first one (:364) on the same rectified images with the same left
  // keypoints, so their match is restored instead of recomputed (identical by construction).
  if (reuse_tracked && s.mode == 2 && dc.use_ransac && i < s.nr_tracked) {
    if (threadIdx.x == 0) db.fr.rstat[k] = db.fr.mstat[k];
    return;
  }
  const int tc = dc.templ_cols, tr = dc.templ_rows, sc = dc.stripe_cols, sr = dc.stripe_rows;
  const int W = dc.W, H = dc.H;
  if (lst != KVFE_KP_VALID) {
    if (threadIdx.x == 0) { db.fr.rstat[k] = lst; db.fr.mstat[k] = lst; db.fr.rrx[k] = 0.f; db.fr.rry[k] = 0.f; }
    return;
  }
  const int rx = (int)roundf(db.fr.lrx[k]), ry = (int)roundf(db.fr.lry[k]);
  const int tcy = ry - (tr - 1) / 2;
  const int scy = ry - (sr - 1) / 2;
  if (tcy < 0 || tcy + tr > H - 1 || scy < 0 || scy + sr > H - 1) {
    if (threadIdx.x == 0) {
      db.fr.rstat[k] = KVFE_KP_NO_RIGHT_RECT; db.fr.mstat[k] = KVFE_KP_NO_RIGHT_RECT;
      db.fr.rrx[k] = 0.f; db.fr.rry[k] = 0.f;
    }
    return;
  }
  int offset_temp = 0;
  int tcx = rx - (tc - 1) / 2;
  if (tcx < 0) { offset_temp = tcx; tcx = 0; }
  if (tcx + tc > W - 1) { offset_temp = (tcx + tc) - (W - 1); tcx -= offset_temp; }
  int scx = rx + (tc - 1) / 2 - sc;
  if (scx + sc > W - 1) scx -= (scx + sc) - (W - 1);
  if (scx < 0) scx = 0;
  // cv::Rect of the template / stripe must lie inside the image (OpenCV would assert otherwise)
  tcx = clampi(tcx, 0, max(W - tc, 0));
  // shared memory: template rows padded to a multiple of 4 bytes (zero pad), stripe rows padded so
  // that every shifted 4-byte window can be assembled from two aligned words, then the scores.
  const int tstride = (tc + 3) & ~3;
  const int sstride = ((sc + 3) & ~3) + 8;
  unsigned char* templ = smraw;
  unsigned char* stripe = smraw + ((tstride * tr + 15) & ~15);
  int* score = reinterpret_cast<int*>(stripe + ((sstride * sr + 15) & ~15));
  const unsigned char* L = db.rectL + (size_t)b * dc.img_stride;
  const unsigned char* R = db.rectR + (size_t)b * dc.img_stride;
  for (int t = threadIdx.x; t < tstride * tr; t += blockDim.x) {
    int r = t / tstride, c = t - r * tstride;
    templ[t] = (c < tc) ? L[(size_t)(tcy + r) * dc.pitch + tcx + c] : 0;
  }
  for (int t = threadIdx.x; t < sstride * sr; t += blockDim.x) {
    int r = t / sstride, c = t - r * sstride;
    stripe[t] = (c < sc) ? R[(size_t)(scy + r) * dc.pitch + scx + c] : 0;
  }
  __syncthreads();
  const int npos_x = sc - tc + 1, npos_y = sr - tr + 1;
  const int npos = npos_x * npos_y;
  const int nwords = tstride >> 2;
  // byte mask of the last template word (columns >= tc must not contribute to sum S^2)
  const unsigned int lastmask = (tc & 3) ? (0xffffffffu >> (8 * (4 - (tc & 3)))) : 0xffffffffu;
  // exact integer TM_SQDIFF = sum S^2 - 2 sum S*T + sum T^2, four pixels per dp4a
  unsigned int tt = 0;
  for (int r = 0; r < tr; ++r) {
    const unsigned int* tw = reinterpret_cast<const unsigned int*>(templ + r * tstride);
    for (int w = 0; w < nwords; ++w) tt = __dp4a(tw[w], tw[w], tt);
  }
  for (int p = threadIdx.x; p < npos; p += blockDim.x) {
    const int py = p / npos_x, px = p - py * npos_x;
    const int sh = 8 * (px & 3);
    unsigned int st = 0, ss = 0;
    for (int r = 0; r < tr; ++r) {
      const unsigned int* sw = reinterpret_cast<const unsigned int*>(stripe + (py + r) * sstride) + (px >> 2);
      const unsigned int* tw = reinterpret_cast<const unsigned int*>(templ + r * tstride);
      unsigned int lo = sw[0];
      for (int w = 0; w < nwords; ++w) {
        const unsigned int hi = sw[w + 1];
        unsigned int v = __funnelshift_r(lo, hi, sh);
        lo = hi;
        st = __dp4a(v, tw[w], st);
        if (w == nwords - 1) v &= lastmask;
        ss = __dp4a(v, v, ss);
      }
    }
    score[p] = (int)(ss + tt - 2u * st);
  }
  __syncthreads();
  // first minimum in row-major order (cv::minMaxLoc)
  __shared__ unsigned long long best[MATCH_THREADS / 32];
  unsigned long long key = ~0ull;
  for (int p = threadIdx.x; p < npos; p += blockDim.x) {
    unsigned long long kk = ((unsigned long long)(unsigned int)score[p] << 32) | (unsigned int)p;
    key = kk < key ? kk : key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(KVFE_FULL_MASK, key, o);
    key = other < key ? other : key;
  }
  if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp 0 finishes; optional sub-pixel refinement needs a full warp
    unsigned long long kk = best[0];
    for (int w = 1; w < MATCH_THREADS / 32; ++w) kk = best[w] < kk ? best[w] : kk;
    int p = (int)(kk & 0xffffffffu);
    int py = p / npos_x, px = p - py * npos_x;
    float mx = (float)(px + scx + (tc - 1) / 2 + offset_temp);
    float my = (float)(py + scy + (tr - 1) / 2);
    if (dc.subpix_stereo) {
      // cv::cornerSubPix(right_rectified, win (10,10), zeroZone (-1,-1), criteria(EPS+ITER, 40, 0.001))
      float* buf = reinterpret_cast<float*>(smraw);     // template/stripe no longer needed (23*23*4 B)
      corner_subpix_warp(R, dc.pitch, W, H, 10, 40, 0.001 * 0.001, db.subpix_mask_stereo, buf, &mx, &my,
                         threadIdx.x);
    }
    if (threadIdx.x == 0) {
      // after NORM_MINMAX the minimum is ~0, so "min_val < tolerance" holds iff tolerance > 0
      const int ms = (dc.tol_templ > 0.f) ? KVFE_KP_VALID : KVFE_KP_NO_RIGHT_RECT;
      db.fr.rstat[k] = ms; db.fr.mstat[k] = ms;
      db.fr.rrx[k] = mx; db.fr.rry[k] = my;
    }
  }
}

__global__ void __launch_bounds__(128) depth_kernel(DevCfg dc, DevBuf db, const CamModel* __restrict__ cams,
                                                    int mode_mask) {
  const int b = blockIdx.y;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= db.fr.n[fs]) return;
  const size_t k = (size_t)fs * dc.cap + i;
  int ls = db.fr.lstat[k], rs = db.fr.rstat[k];
  double depth = 0.0;
  if (ls == KVFE_KP_VALID && rs == KVFE_KP_VALID) {
    double disparity = (double)(db.fr.lrx[k] - db.fr.rrx[k]);
    if (disparity >= 0.0) {
      double d = dc.fx_b / disparity;
      if (d < dc.min_depth || d > dc.max_depth) rs = KVFE_KP_NO_DEPTH; else depth = d;
    } else rs = KVFE_KP_NO_DEPTH;
  } else if (ls != KVFE_KP_VALID && rs != ls) rs = ls;
  db.fr.rstat[k] = rs;
  db.fr.depth[k] = depth;
  // right_frame_.keypoints_ : right maps at the rounded rectified position
  float rkx = 0.f, rky = 0.f;
  if (rs == KVFE_KP_VALID) {
    int xx = clampi((int)roundf(db.fr.rrx[k]), 0, dc.W - 1), yy = clampi((int)roundf(db.fr.rry[k]), 0, dc.H - 1);
    rect_map_at(cams[1], xx, yy, &rkx, &rky);
  }
  db.fr.rkx[k] = rkx; db.fr.rky[k] = rky;
  double p0 = 0, p1 = 0, p2 = 0;
  if (rs == KVFE_KP_VALID) {
    const double* v = db.fr.versor + 3 * k;
    p0 = v[0] * depth / v[2]; p1 = v[1] * depth / v[2]; p2 = v[2] * depth / v[2];
  }
  db.fr.p3d[3 * k] = p0; db.fr.p3d[3 * k + 1] = p1; db.fr.p3d[3 * k + 2] = p2;
}

__global__ void undistort_kernel(const CamModel* __restrict__ cams, int cam, int mode, const float* x,
                                 const float* y, int n, float* ox, float* oy) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  undistort_point(cams[cam], x[i], y[i], mode, &ox[i], &oy[i]);
}

__global__ void bearing_kernel(const CamModel* __restrict__ cams, const float* x, const float* y, int n,
                               double* versors) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ux, uy;
  undistort_point(cams[0], x[i], y[i], 1, &ux, &uy);
  double v0 = (double)ux, v1 = (double)uy, v2 = 1.0;
  double n2 = v0 * v0 + (v1 * v1 + v2 * v2);
  double nrm = sqrt(n2);
  if (n2 > 0) { v0 = v0 / nrm; v1 = v1 / nrm; v2 = v2 / nrm; }
  versors[3 * i] = v0; versors[3 * i + 1] = v1; versors[3 * i + 2] = v2;
}

int launch_sparse_stereo(const DevCfg& dc, const DevBuf& db, const CamModel* d_cam, int mode_mask,
                         int reuse_tracked, cudaStream_t s) {
  int n = 0;
  left_rect_kernel<<<dim3((dc.cap + 127) / 128, dc.B), 128, 0, s>>>(dc, db, d_cam, mode_mask); ++n;
  const int tstride = (dc.templ_cols + 3) & ~3, sstride = ((dc.stripe_cols + 3) & ~3) + 8;
  size_t sm = ((tstride * dc.templ_rows + 15) & ~15) + ((sstride * dc.stripe_rows + 15) & ~15) +
              sizeof(int) * (size_t)(dc.stripe_cols - dc.templ_cols + 1) * (dc.stripe_rows - dc.templ_rows + 1);
  if (sm < (size_t)SUBPIX_PATCH * 4) sm = (size_t)SUBPIX_PATCH * 4;
  static size_t attr = 0;
  if (sm > 48 * 1024 && sm > attr) {
    cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr = sm;
  }
  match_kernel<<<dim3(dc.cap, dc.B), MATCH_THREADS, sm, s>>>(dc, db, mode_mask, reuse_tracked); ++n;
  depth_kernel<<<dim3((dc.cap + 127) / 128, dc.B), 128, 0, s>>>(dc, db, d_cam, mode_mask); ++n;
  return n;
}

int launch_undistort(const DevCfg& dc, const CamModel* d_cam, int cam, int use_R, int use_P,
                     const float* x, const float* y, int n, float* ox, float* oy, cudaStream_t s) {
  int mode = use_R ? (use_P ? 2 : 1) : (use_P ? 3 : 0);
  undistort_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_cam, cam, mode, x, y, n, ox, oy);
  (void)dc;
  return 1;
}

int launch_bearing(const DevCfg& dc, const CamModel* d_cam, const float* x, const float* y, int n,
                   double* versors, cudaStream_t s) {
  bearing_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_cam, x, y, n, versors);
  (void)dc;
  return 1;
}
