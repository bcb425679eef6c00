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
#define MATCH_TP 8          // zero bytes in front of every template row (shifted reads start at -7)

// s32 += u8 x u8, 16x8x32 (legacy warp-level tensor-core path; exact integer arithmetic)
__device__ __forceinline__ void mma_u8(int (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                       unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// geometry of the shared-memory tiles, shared by the kernel and the launcher
struct MatchGeom { int nq, mt, tstride, sstride, npos_x, npos_y; size_t off_stripe, off_score, off_pref, bytes; };
__host__ __device__ inline MatchGeom match_geom(int tc, int tr, int sc, int sr) {
  MatchGeom g;
  g.npos_x = sc - tc + 1; g.npos_y = sr - tr + 1;
  g.nq = (tc + 7 + 31) / 32;                        // 32-byte K chunks per template row (c' = j + c < tc + 7)
  g.mt = ((g.npos_x + 7) / 8 + 15) / 16;            // 16-row M tiles; row i covers shifts 8i .. 8i+7
  g.tstride = MATCH_TP + 32 * g.nq + 4;
  g.sstride = 128 * g.mt + 32 * g.nq;
  g.off_stripe = ((size_t)g.tstride * tr + 15) & ~(size_t)15;
  g.off_score = g.off_stripe + (((size_t)g.sstride * sr + 15) & ~(size_t)15);
  g.off_pref = g.off_score + sizeof(int) * (size_t)g.npos_x * g.npos_y;
  g.bytes = g.off_pref + sizeof(int) * (size_t)(sc + 1);
  return g;
}

// One keypoint (CTA-wide).  TM_SQDIFF(p) = sum S^2 - 2 sum S*T + sum T^2 in exact integers:
//  * sum S*T for all shifts of one stripe row band is a GEMM: with p = 8 i + j,
//      corr[8i + j] = sum_r sum_c' S[r][8i + c'] * T[r][c' - j]        (c' = j + c)
//    A[i][(r, c')] = stripe bytes (16 x 32 fragments are plain aligned words of the staged stripe),
//    B[(r, c')][j] = the template shifted right by j (funnel-shifted words of the zero-padded
//    template), D = 16 x 8 int32 -> mma.sync m16n8k32 u8.  Warp w takes template rows r = w mod 4
//    for all M tiles; the partial sums meet in shared memory (integer adds, order-free).
//  * sum S^2 per shift = difference of a prefix sum over squared column sums.
__device__ void match_one(const DevCfg& dc, const DevBuf& db, unsigned char* smraw, const StreamState& s, const int b,
                          const int fs, const int i, const int reuse_tracked) {
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
  const MatchGeom g = match_geom(tc, tr, sc, sr);
  const int tstride = g.tstride, sstride = g.sstride;
  unsigned char* templ = smraw;
  unsigned char* stripe = smraw + g.off_stripe;
  int* score = reinterpret_cast<int*>(smraw + g.off_score);          // sum S*T, then the SQDIFF
  unsigned int* pref = reinterpret_cast<unsigned int*>(smraw + g.off_pref);
  __shared__ unsigned int red[MATCH_THREADS / 32];
  const unsigned char* L = db.rectL + (size_t)b * dc.img_stride;
  const unsigned char* R = db.rectR + (size_t)b * dc.img_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- stage the template (zero padded on both sides) and the stripe (aligned words, zero tail)
  for (int t = tid; t < tstride * tr; t += MATCH_THREADS) {
    const int r = t / tstride, c = t - r * tstride - MATCH_TP;
    templ[t] = (c >= 0 && c < tc) ? L[(size_t)(tcy + r) * dc.pitch + tcx + c] : 0;
  }
  {
    const int wpr = sstride >> 2;
    const float inv_wpr = 1.0f / (float)wpr;
    for (int t = tid; t < wpr * sr; t += MATCH_THREADS) {
      const int r = (int)(((float)t + 0.5f) * inv_wpr), w = t - r * wpr;
      unsigned int v = 0;
      const int c0 = 4 * w;
      if (c0 < sc) {
        const unsigned char* src = R + (size_t)(scy + r) * dc.pitch + scx + c0;
        const size_t a = reinterpret_cast<size_t>(src);
        const unsigned int* wp = reinterpret_cast<const unsigned int*>(a & ~(size_t)3);
        const int sh = 8 * (int)(a & 3);
        const unsigned int lo = wp[0], hi = sh ? wp[1] : 0u;
        v = __funnelshift_r(lo, hi, sh);
        if (c0 + 4 > sc) v &= 0xffffffffu >> (8 * (c0 + 4 - sc));
      }
      reinterpret_cast<unsigned int*>(stripe)[(size_t)r * wpr + w] = v;
    }
  }
  const int npos_x = g.npos_x, npos_y = g.npos_y;
  const int npos = npos_x * npos_y;
  for (int p = tid; p < npos; p += MATCH_THREADS) score[p] = 0;
  __syncthreads();
  // ---- sum T^2
  unsigned int tt = 0;
  for (int t = tid; t < (tstride >> 2) * tr; t += MATCH_THREADS) {
    const unsigned int v = reinterpret_cast<const unsigned int*>(templ)[t];
    tt = __dp4a(v, v, tt);
  }
  for (int o = 16; o > 0; o >>= 1) tt += __shfl_xor_sync(KVFE_FULL_MASK, tt, o);
  if (lane == 0) red[warp] = tt;
  // ---- sum S*T on the tensor cores
  const int gq = lane >> 2, tig = lane & 3;
  for (int py = 0; py < npos_y; ++py) {
    for (int mt0 = 0; mt0 < g.mt; mt0 += 4) {
      int acc[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0; }
      for (int r = warp; r < tr; r += MATCH_THREADS / 32) {
        const unsigned int* trow = reinterpret_cast<const unsigned int*>(templ + r * tstride);
        const unsigned char* srow = stripe + (size_t)(py + r) * sstride;
        for (int q = 0; q < g.nq; ++q) {
          // B fragment: template bytes x .. x+3 with x = 32 q + 4 tig - gq (+16), row-padded by MATCH_TP
          const int o0 = MATCH_TP + 32 * q + 4 * tig - gq;
          const int sh = 8 * (o0 & 3);
          const unsigned int b0 = __funnelshift_r(trow[o0 >> 2], trow[(o0 >> 2) + 1], sh);
          const unsigned int b1 = __funnelshift_r(trow[(o0 >> 2) + 4], trow[(o0 >> 2) + 5], sh);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (mt0 + m < g.mt) {
              const unsigned int* ap = reinterpret_cast<const unsigned int*>(srow + 128 * (mt0 + m) + 8 * gq + 32 * q) + tig;
              mma_u8(acc[m], ap[0], ap[16], ap[4], ap[20], b0, b1);
            }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (mt0 + m < g.mt) {
          const int p0 = 8 * (16 * (mt0 + m) + gq) + 2 * tig;
          int* sp = score + py * npos_x;
          if (p0 < npos_x) atomicAdd(sp + p0, acc[m][0]);
          if (p0 + 1 < npos_x) atomicAdd(sp + p0 + 1, acc[m][1]);
          if (p0 + 64 < npos_x) atomicAdd(sp + p0 + 64, acc[m][2]);
          if (p0 + 65 < npos_x) atomicAdd(sp + p0 + 65, acc[m][3]);
        }
      }
    }
  }
  __syncthreads();
  tt = 0;
  for (int w = 0; w < MATCH_THREADS / 32; ++w) tt += red[w];
  // ---- sum S^2 per shift: prefix over the squared column sums of the row band, then the SQDIFF
  for (int py = 0; py < npos_y; ++py) {
    const int ch = (sc + MATCH_THREADS - 1) / MATCH_THREADS;      // consecutive columns per thread
    const int c0 = tid * ch;
    unsigned int loc = 0;
    for (int c = c0; c < min(c0 + ch, sc); ++c) {
      unsigned int cs = 0;
      for (int r = 0; r < tr; ++r) { const unsigned int v = stripe[(size_t)(py + r) * sstride + c]; cs += v * v; }
      pref[c] = cs;                                               // own columns only: no barrier needed
      loc += cs;
    }
    // exclusive block scan of the per-thread sums
    unsigned int inc = loc;
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(KVFE_FULL_MASK, inc, o); if (lane >= o) inc += y; }
    __syncthreads();                                              // red free to be rewritten
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    unsigned int base = inc - loc;
    for (int w = 0; w < warp; ++w) base += red[w];
    for (int c = c0; c < min(c0 + ch, sc); ++c) { const unsigned int cs = pref[c]; pref[c] = base; base += cs; }
    if (c0 < sc && c0 + ch >= sc) pref[sc] = base;
    __syncthreads();
    for (int px = tid; px < npos_x; px += MATCH_THREADS) {
      const unsigned int ss = pref[px + tc] - pref[px];
      const unsigned int st = (unsigned int)score[py * npos_x + px];
      score[py * npos_x + px] = (int)(ss + tt - 2u * st);
    }
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

#define MATCH_KP_PER_CTA 4
// grid (ceil(cap / MATCH_KP_PER_CTA), B): a CTA walks keypoints blockIdx.x, blockIdx.x + gridDim.x, ...
// (fewer, longer-lived CTAs: most slots of the capacity-sized grid hold no live keypoint)
__global__ void __launch_bounds__(MATCH_THREADS) match_kernel(DevCfg dc, DevBuf db, int mode_mask, int reuse_tracked) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int b = blockIdx.y;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int n = db.fr.n[fs];
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    match_one(dc, db, smraw, s, b, fs, i, reuse_tracked);
    __syncthreads();                       // shared tiles are reused by the next keypoint
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
  size_t sm = match_geom(dc.templ_cols, dc.templ_rows, dc.stripe_cols, dc.stripe_rows).bytes;
  if (sm < (size_t)SUBPIX_PATCH * 4) sm = (size_t)SUBPIX_PATCH * 4;
  static size_t attr = 0;
  if (sm > 48 * 1024 && sm > attr) {
    cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr = sm;
  }
  match_kernel<<<dim3((dc.cap + MATCH_KP_PER_CTA - 1) / MATCH_KP_PER_CTA, dc.B), MATCH_THREADS, sm, s>>>(dc, db, mode_mask, reuse_tracked); ++n;
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


// ---- stage-level pieces of the sparse stereo path (include/kvfe.h, "boundary completion") ----------------
// UndistorterRectifier::checkUndistortedRectifiedLeftKeypoints (UndistorterRectifier.cpp:138-211) with the
// caller's undistorted keypoints and pixel tolerance; cam selects the maps (0: left, 1: right).
__global__ void __launch_bounds__(128) check_rect_raw_kernel(DevCfg dc, const CamModel* __restrict__ cams, int cam,
                                                             const float* __restrict__ dx, const float* __restrict__ dy,
                                                             const float* __restrict__ ux_in, const float* __restrict__ uy_in, int n,
                                                             float tol, int* __restrict__ status, float* __restrict__ ox,
                                                             float* __restrict__ oy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ux = ux_in[i], uy = uy_in[i];
  bool cropped = false;                              // UtilsOpenCV::cropToSize
  const float mw = (float)(dc.W - 1), mh = (float)(dc.H - 1);
  if (ux > mw) { ux = mw; cropped = true; } else if (ux < 0.0f) { ux = 0.0f; cropped = true; }
  if (uy > mh) { uy = mh; cropped = true; } else if (uy < 0.0f) { uy = 0.0f; cropped = true; }
  const int rx = clampi((int)roundf(ux), 0, dc.W - 1), ry = clampi((int)roundf(uy), 0, dc.H - 1);
  float ex, ey;
  rect_map_at(cams[cam], rx, ry, &ex, &ey);
  int st = KVFE_KP_VALID;
  if (cropped || fabsf(dx[i] - ex) > tol || fabsf(dy[i] - ey) > tol) st = KVFE_KP_NO_LEFT_RECT;
  status[i] = st; ox[i] = ux; oy[i] = uy;
}
// UndistorterRectifier::distortUnrectifyKeypoints (UndistorterRectifier.cpp:213-228)
__global__ void __launch_bounds__(128) distort_unrectify_raw_kernel(DevCfg dc, const CamModel* __restrict__ cams, int cam,
                                                                    const int* __restrict__ status, const float* __restrict__ x,
                                                                    const float* __restrict__ y, int n, float* __restrict__ ox,
                                                                    float* __restrict__ oy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float rx = 0.f, ry = 0.f;
  if (status[i] == KVFE_KP_VALID) {
    const int xx = clampi((int)roundf(x[i]), 0, dc.W - 1), yy = clampi((int)roundf(y[i]), 0, dc.H - 1);
    rect_map_at(cams[cam], xx, yy, &rx, &ry);
  }
  ox[i] = rx; oy[i] = ry;
}
int launch_check_rect_raw(const DevCfg& dc, const CamModel* d_cam, int cam, const float* dx, const float* dy, const float* ux,
                          const float* uy, int n, float tol, int* status, float* ox, float* oy, cudaStream_t s) {
  check_rect_raw_kernel<<<(n + 127) / 128, 128, 0, s>>>(dc, d_cam, cam, dx, dy, ux, uy, n, tol, status, ox, oy);
  return 1;
}
int launch_distort_unrectify_raw(const DevCfg& dc, const CamModel* d_cam, int cam, const int* status, const float* x,
                                 const float* y, int n, float* ox, float* oy, cudaStream_t s) {
  distort_unrectify_raw_kernel<<<(n + 127) / 128, 128, 0, s>>>(dc, d_cam, cam, status, x, y, n, ox, oy);
  return 1;
}
// the three kernels of launch_sparse_stereo one at a time (which: 0 left_rect, 1 match, 2 depth), stage state in stream 0
int launch_sparse_stereo_part(const DevCfg& dc, const DevBuf& db, const CamModel* d_cam, int mode_mask, int which, cudaStream_t s) {
  if (which == 0) { left_rect_kernel<<<dim3((dc.cap + 127) / 128, dc.B), 128, 0, s>>>(dc, db, d_cam, mode_mask); return 1; }
  if (which == 2) { depth_kernel<<<dim3((dc.cap + 127) / 128, dc.B), 128, 0, s>>>(dc, db, d_cam, mode_mask); return 1; }
  size_t sm = match_geom(dc.templ_cols, dc.templ_rows, dc.stripe_cols, dc.stripe_rows).bytes;
  if (sm < (size_t)SUBPIX_PATCH * 4) sm = (size_t)SUBPIX_PATCH * 4;
  if (sm > 48 * 1024) cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  match_kernel<<<dim3((dc.cap + MATCH_KP_PER_CTA - 1) / MATCH_KP_PER_CTA, dc.B), MATCH_THREADS, sm, s>>>(dc, db, mode_mask, 0);
  return 1;
}
