// lk.cu -- row a9: cv::calcOpticalFlowPyrLK as called by Tracker::featureTracking
// (reference src/frontend/Tracker.cpp:117-148; winSize 24, maxLevel 4, criteria(COUNT+EPS, 30, 0.1),
// OPTFLOW_USE_INITIAL_FLOW, minEigThreshold 1e-4).
//
// One warp per keypoint, levels maxLevel..0 inside the kernel.  Per level:
//   1. a (win+3)^2 patch of the PREVIOUS image level (BORDER_REFLECT_101 padding semantics) is
//      staged in shared memory; Scharr derivatives are computed from it on the fly (zero outside
//      the image, like the zero-padded derivative buffer of OpenCV) -- no derivative image is ever
//      written to HBM;
//   2. the fixed-point window I / Ix / Iy (14-bit bilinear weights) and the structure tensor are
//      formed; the float accumulations reproduce OpenCV's SIMD128 lane order exactly (bit-exact
//      with cv2 for windows that are a multiple of 8, e.g. the Euroc 24);
//   3. <= maxCount Gauss-Newton iterations, each staging a (win+1)^2 patch of the NEXT image level.
// All float scalar arithmetic follows LKTrackerInvoker's scalar code path operation by operation.
#include "common.cuh"

#define LK_MAX_WIN 32
#define LK_WARPS 4

// per-warp shared-memory carve-up (sizes depend on the runtime window):
//   P (win+3)^2 u8 | J (win+1)^2 u8 | Dx, Dy (win+1)^2 i16 | I, Ix, Iy win^2 i16
struct LkSmem {
  unsigned char* P; unsigned char* J;
  short *Dx, *Dy, *I, *Ix, *Iy;
};
__host__ __device__ inline size_t lk_warp_bytes(int win) {
  size_t a = ((size_t)(win + 3) * (win + 3) + 15) & ~(size_t)15;
  size_t b = ((size_t)(win + 1) * (win + 1) + 15) & ~(size_t)15;
  size_t d = (2 * (size_t)(win + 1) * (win + 1) + 15) & ~(size_t)15;
  size_t w = (2 * (size_t)win * win + 15) & ~(size_t)15;
  return a + b + 2 * d + 3 * w;
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// grid (ceil(cap / LK_WARPS), B)
__global__ void __launch_bounds__(LK_WARPS * 32) lk_kernel(DevCfg dc, DevBuf db, int prev_slot, int cur_slot) {
  extern __shared__ __align__(16) unsigned char lk_smem_raw[];
  const int b = blockIdx.y;
  const StreamState& st = db.st[b];
  if (st.mode == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = blockIdx.x * LK_WARPS + warp;
  if (pt >= st.n_ref) return;
  LkSmem s;
  {
    const int w_ = dc.win;
    unsigned char* base = lk_smem_raw + (size_t)warp * lk_warp_bytes(w_);
    size_t a = ((size_t)(w_ + 3) * (w_ + 3) + 15) & ~(size_t)15;
    size_t bq = ((size_t)(w_ + 1) * (w_ + 1) + 15) & ~(size_t)15;
    size_t d = (2 * (size_t)(w_ + 1) * (w_ + 1) + 15) & ~(size_t)15;
    size_t ww = (2 * (size_t)w_ * w_ + 15) & ~(size_t)15;
    s.P = base; s.J = base + a;
    s.Dx = reinterpret_cast<short*>(base + a + bq); s.Dy = reinterpret_cast<short*>(base + a + bq + d);
    s.I = reinterpret_cast<short*>(base + a + bq + 2 * d);
    s.Ix = reinterpret_cast<short*>(base + a + bq + 2 * d + ww);
    s.Iy = reinterpret_cast<short*>(base + a + bq + 2 * d + 2 * ww);
  }
  const size_t gi = (size_t)b * dc.cap + pt;
  const unsigned char* prevPyr = db.pyr[prev_slot] + (size_t)b * dc.pyr_stride;
  const unsigned char* nextPyr = db.pyr[cur_slot] + (size_t)b * dc.pyr_stride;
  const int win = dc.win;
  const float halfWin = (win - 1) * 0.5f;
  const float px0 = db.lk_px[gi], py0 = db.lk_py[gi];
  float nx = db.lk_qx[gi], ny = db.lk_qy[gi];     // running nextPts[ptidx]
  bool status = true;
  const int maxLevel = dc.n_levels - 1;
  const float FLT_SCALE = 1.f / (1 << 20);
  const int pw = win + 3, jw = win + 1, dw = win + 1;

  for (int level = maxLevel; level >= 0; --level) {
    const int cols = dc.lvl_w[level], rows = dc.lvl_h[level], pitch = dc.lvl_pitch[level];
    const unsigned char* I = prevPyr + dc.lvl_off[level];
    const unsigned char* Jimg = nextPyr + dc.lvl_off[level];
    const float scl = (float)(1. / (1 << level));
    float ppx = px0 * scl, ppy = py0 * scl;
    if (level == maxLevel) { nx = nx * scl; ny = ny * scl; }   // OPTFLOW_USE_INITIAL_FLOW
    else { nx = nx * 2.f; ny = ny * 2.f; }
    ppx -= halfWin; ppy -= halfWin;
    const int ipx = cv_floor(ppx), ipy = cv_floor(ppy);
    if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
      if (level == 0) status = false;
      continue;
    }
    // ---- stage prev patch rows ipy-1 .. ipy+win+1, cols ipx-1 .. ipx+win+1 (reflect101 padding)
    for (int i = lane; i < pw * pw; i += 32) {
      int r = i / pw, c = i - r * pw;
      int yy = reflect101(ipy - 1 + r, rows), xx = reflect101(ipx - 1 + c, cols);
      s.P[i] = I[(size_t)yy * pitch + xx];
    }
    __syncwarp();
    // ---- Scharr derivatives at the (win+1)^2 tap positions; zero outside the image
    for (int i = lane; i < dw * dw; i += 32) {
      int r = i / dw, c = i - r * dw;
      int X = ipx + c, Y = ipy + r;
      short dx = 0, dy = 0;
      if (X >= 0 && X < cols && Y >= 0 && Y < rows) {
        const unsigned char* p = s.P + (r + 1) * pw + (c + 1);
        int a00 = p[-pw - 1], a01 = p[-pw], a02 = p[-pw + 1];
        int a10 = p[-1], a12 = p[1];
        int a20 = p[pw - 1], a21 = p[pw], a22 = p[pw + 1];
        dx = (short)(3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20));
        dy = (short)(3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02));
      }
      s.Dx[i] = dx; s.Dy[i] = dy;
    }
    __syncwarp();
    // ---- window of I, Ix, Iy with 14-bit bilinear weights; structure tensor
    float a = ppx - ipx, bb = ppy - ipy;
    int iw00 = cv_round((1.f - a) * (1.f - bb) * (1 << 14));
    int iw01 = cv_round(a * (1.f - bb) * (1 << 14));
    int iw10 = cv_round((1.f - a) * bb * (1 << 14));
    int iw11 = (1 << 14) - iw00 - iw01 - iw10;
    for (int i = lane; i < win * win; i += 32) {
      int y = i / win, x = i - y * win;
      const unsigned char* p = s.P + (y + 1) * pw + (x + 1);
      int ival = descale(p[0] * iw00 + p[1] * iw01 + p[pw] * iw10 + p[pw + 1] * iw11, 14 - 5);
      const short* d = s.Dx + y * dw + x;
      int ixval = descale(d[0] * iw00 + d[1] * iw01 + d[dw] * iw10 + d[dw + 1] * iw11, 14);
      d = s.Dy + y * dw + x;
      int iyval = descale(d[0] * iw00 + d[1] * iw01 + d[dw] * iw10 + d[dw + 1] * iw11, 14);
      s.I[i] = (short)ival; s.Ix[i] = (short)ixval; s.Iy[i] = (short)iyval;
    }
    __syncwarp();
    // Structure tensor with OpenCV's SIMD128 accumulation order (pinned against cv2: scratch probe,
    // 2400/2400 points bit-exact): four float lanes, pixel i (row-major) -> lane i % 4, sequential
    // float accumulation per lane, final (L0 + L2) + (L1 + L3).  Lanes 0..11 = 3 sums x 4 chains.
    float A11, A12, A22;
    {
      float acc = 0.f;
      const int q = lane >> 2, c = lane & 3;
      if (lane < 12) {
        const short* u = (q == 2) ? s.Iy : s.Ix;
        const short* v = (q == 0) ? s.Ix : s.Iy;
        for (int i = c; i < win * win; i += 4) acc = acc + (float)((int)u[i] * (int)v[i]);
      }
      float l2 = __shfl_down_sync(KVFE_FULL_MASK, acc, 2);
      float l1 = __shfl_down_sync(KVFE_FULL_MASK, acc, 1);
      float l3 = __shfl_down_sync(KVFE_FULL_MASK, acc, 3);
      float tot = (acc + l2) + (l1 + l3);          // valid in lanes 0, 4, 8
      A11 = __shfl_sync(KVFE_FULL_MASK, tot, 0) * FLT_SCALE;
      A12 = __shfl_sync(KVFE_FULL_MASK, tot, 4) * FLT_SCALE;
      A22 = __shfl_sync(KVFE_FULL_MASK, tot, 8) * FLT_SCALE;
    }
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (minEig < dc.min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) status = false;
      continue;
    }
    D = 1.f / D;
    float qx = nx - halfWin, qy = ny - halfWin;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < dc.max_iter; ++j) {
      const int iqx = cv_floor(qx), iqy = cv_floor(qy);
      if (iqx < -win || iqx >= cols || iqy < -win || iqy >= rows) {
        if (level == 0) status = false;
        break;
      }
      a = qx - iqx; bb = qy - iqy;
      iw00 = cv_round((1.f - a) * (1.f - bb) * (1 << 14));
      iw01 = cv_round(a * (1.f - bb) * (1 << 14));
      iw10 = cv_round((1.f - a) * bb * (1 << 14));
      iw11 = (1 << 14) - iw00 - iw01 - iw10;
      for (int i = lane; i < jw * jw; i += 32) {
        int r = i / jw, c = i - r * jw;
        int yy = reflect101(iqy + r, rows), xx = reflect101(iqx + c, cols);
        s.J[i] = Jimg[(size_t)yy * pitch + xx];
      }
      __syncwarp();
      // residuals into shared memory (reuses the Dx tap buffer: (win+1)^2 >= win^2 shorts)
      short* diffv = s.Dx;
      for (int i = lane; i < win * win; i += 32) {
        int y = i / win, x = i - y * win;
        const unsigned char* p = s.J + y * jw + x;
        diffv[i] = (short)(descale(p[0] * iw00 + p[1] * iw01 + p[jw] * iw10 + p[jw + 1] * iw11, 14 - 5) - s.I[i]);
      }
      __syncwarp();
      // OpenCV's SIMD128 order for the mismatch vector (pinned against cv2): groups of 8 pixels
      // (row-major); chain m in 0..3 accumulates float(int32(d[k]*G[k] + d[k+4]*G[k+4])), k = 8g + m;
      // final (c0 + c2) + (c1 + c3).  Lanes 0..7 = 2 sums x 4 chains.
      float b1, b2;
      {
        float acc = 0.f;
        if (lane < 8) {
          const short* G = (lane < 4) ? s.Ix : s.Iy;
          const int m = lane & 3;
          for (int k = m; k + 4 < win * win; k += 8)
            acc = acc + (float)((int)diffv[k] * (int)G[k] + (int)diffv[k + 4] * (int)G[k + 4]);
        }
        float l2 = __shfl_down_sync(KVFE_FULL_MASK, acc, 2);
        float l1 = __shfl_down_sync(KVFE_FULL_MASK, acc, 1);
        float l3 = __shfl_down_sync(KVFE_FULL_MASK, acc, 3);
        float tot = (acc + l2) + (l1 + l3);        // valid in lanes 0 and 4
        b1 = __shfl_sync(KVFE_FULL_MASK, tot, 0) * FLT_SCALE;
        b2 = __shfl_sync(KVFE_FULL_MASK, tot, 4) * FLT_SCALE;
      }
      __syncwarp();
      float dxv = (float)((A12 * b2 - A22 * b1) * D);
      float dyv = (float)((A12 * b1 - A11 * b2) * D);
      qx += dxv; qy += dyv;
      nx = qx + halfWin; ny = qy + halfWin;
      if ((double)dxv * (double)dxv + (double)dyv * (double)dyv <= (double)dc.eps2) break;
      if (j > 0 && fabsf(dxv + pdx) < 0.01 && fabsf(dyv + pdy) < 0.01) {
        nx -= dxv * 0.5f; ny -= dyv * 0.5f;
        break;
      }
      pdx = dxv; pdy = dyv;
    }
  }
  if (status) {
    // the error pass of calcOpticalFlowPyrLK (level 0, `err` requested as the reference does,
    // Tracker.cpp:137-146): a FINAL position whose window origin left the image clears the status
    // (the iteration loop only tests the position it starts an iteration from)
    const int fx = cv_floor(nx - halfWin), fy = cv_floor(ny - halfWin);
    if (fx < -win || fx >= dc.lvl_w[0] || fy < -win || fy >= dc.lvl_h[0]) status = false;
  }
  if (lane == 0) {
    db.lk_qx[gi] = nx; db.lk_qy[gi] = ny;
    db.lk_status[gi] = status ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Column-per-lane specialisation for windows that are a multiple of 8 (Euroc: 24): lane x owns
// column x of the window.  Patch columns are loaded one row at a time (a coalesced <=27-byte row per
// instruction), right-hand neighbours come from warp shuffles, the fixed-point window I / Ix / Iy is
// kept in shared memory (conflict-free column access) so that the row loops stay rolled: small code
// (no instruction-cache thrash), few registers, high occupancy.  Arithmetic identical to lk_kernel.
// ------------------------------------------------------------------------------------------------
#define LKC_WARPS 4
#define LKC_CHUNK 6      // window rows per software-pipelined chunk of J loads

template <int WIN>
__global__ void __launch_bounds__(LKC_WARPS * 32) lk_kernel_col(DevCfg dc, DevBuf db, int prev_slot, int cur_slot) {
  constexpr int PW = WIN + 3, TW = WIN + 1;
  static_assert(WIN % LKC_CHUNK == 0 || WIN % 8 == 0, "window must be a multiple of 8");
  __shared__ short sI[LKC_WARPS][WIN * WIN], sIx[LKC_WARPS][WIN * WIN], sIy[LKC_WARPS][WIN * WIN];
  __shared__ float sP[LKC_WARPS][3][WIN * WIN];     // structure-tensor products (xx, xy, yy)
  const int b = blockIdx.y;
  const StreamState& st = db.st[b];
  if (st.mode == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = blockIdx.x * LKC_WARPS + warp;
  if (pt >= st.n_ref) return;
  short* wI = sI[warp]; short* wIx = sIx[warp]; short* wIy = sIy[warp];
  float (*wP)[WIN * WIN] = sP[warp];
  const int xl = min(lane, WIN - 1);               // window column of this lane (lanes >= WIN idle copies)
  const size_t gi = (size_t)b * dc.cap + pt;
  const unsigned char* prevPyr = db.pyr[prev_slot] + (size_t)b * dc.pyr_stride;
  const unsigned char* nextPyr = db.pyr[cur_slot] + (size_t)b * dc.pyr_stride;
  const float halfWin = (WIN - 1) * 0.5f;
  const float px0 = db.lk_px[gi], py0 = db.lk_py[gi];
  float nx = db.lk_qx[gi], ny = db.lk_qy[gi];
  bool status = true;
  const int maxLevel = dc.n_levels - 1;
  const float FLT_SCALE = 1.f / (1 << 20);
  const int m4 = lane & 3;

  for (int level = maxLevel; level >= 0; --level) {
    const int cols = dc.lvl_w[level], rows = dc.lvl_h[level], pitch = dc.lvl_pitch[level];
    const unsigned char* I = prevPyr + dc.lvl_off[level];
    const unsigned char* Jimg = nextPyr + dc.lvl_off[level];
    const float scl = (float)(1. / (1 << level));
    float ppx = px0 * scl, ppy = py0 * scl;
    if (level == maxLevel) { nx = nx * scl; ny = ny * scl; }
    else { nx = nx * 2.f; ny = ny * 2.f; }
    ppx -= halfWin; ppy -= halfWin;
    const int ipx = cv_floor(ppx), ipy = cv_floor(ppy);
    if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
      if (level == 0) status = false;
      continue;
    }
    float a = ppx - ipx, bb = ppy - ipy;
    int iw00 = cv_round((1.f - a) * (1.f - bb) * (1 << 14));
    int iw01 = cv_round(a * (1.f - bb) * (1 << 14));
    int iw10 = cv_round((1.f - a) * bb * (1 << 14));
    int iw11 = (1 << 14) - iw00 - iw01 - iw10;
    __syncwarp();
    // ---- window of I, Ix, Iy streamed over the (WIN+3) patch rows; structure tensor products on the fly
    {
      const int colx = reflect101(ipx - 1 + min(lane, PW - 1), cols);
      const bool xin = (ipx + lane) >= 0 && (ipx + lane) < cols;
      int C0 = 0, R0 = 0, C1 = 0, R1 = 0;
      int h0 = 0, h1 = 0, s0 = 0, s1 = 0;          // per row: horizontal difference R - L and smoothing 3L + 10C + 3R
      int dxp = 0, dyp = 0, dxpn = 0, dypn = 0;
      // one patch row: L2 = pixel of image row (ipy - 1 + r) at this lane's column
      auto patch_row = [&](const int r, const int L2) {
        const int C2 = __shfl_down_sync(KVFE_FULL_MASK, L2, 1);
        const int R2 = __shfl_down_sync(KVFE_FULL_MASK, L2, 2);
        const int h2 = R2 - L2, s2 = 3 * (L2 + R2) + 10 * C2;
        if (r >= 2) {
          const int d = r - 2;
          // Scharr: dx = [3 10 3]^T (rows) x [-1 0 1] (cols), dy = [-1 0 1]^T x [3 10 3]
          int dx = 3 * (h0 + h2) + 10 * h1;
          int dy = s2 - s0;
          const bool yin = (ipy + d) >= 0 && (ipy + d) < rows;
          if (!(xin && yin)) { dx = 0; dy = 0; }
          const int dxn = __shfl_down_sync(KVFE_FULL_MASK, dx, 1);
          const int dyn = __shfl_down_sync(KVFE_FULL_MASK, dy, 1);
          if (d >= 1) {
            const int y = d - 1;
            const int iv = descale(C0 * iw00 + R0 * iw01 + C1 * iw10 + R1 * iw11, 14 - 5);
            const int ix = descale(dxp * iw00 + dxpn * iw01 + dx * iw10 + dxn * iw11, 14);
            const int iy = descale(dyp * iw00 + dypn * iw01 + dy * iw10 + dyn * iw11, 14);
            if (lane < WIN) {
              const int o = y * WIN + lane;
              wI[o] = (short)iv; wIx[o] = (short)ix; wIy[o] = (short)iy;
              wP[0][o] = (float)(ix * ix); wP[1][o] = (float)(ix * iy); wP[2][o] = (float)(iy * iy);
            }
          }
          dxp = dx; dyp = dy; dxpn = dxn; dypn = dyn;
        }
        h0 = h1; h1 = h2; s0 = s1; s1 = s2;
        C0 = C1; R0 = R1; C1 = C2; R1 = R2;
      };
      if (ipy - 1 >= 0 && ipy - 1 + PW <= rows) {
        // all patch rows inside the image: running 32-bit offset, no reflection
        unsigned int o = (unsigned int)(ipy - 1) * (unsigned int)pitch + (unsigned int)colx;
        int Lnext = I[o];
#pragma unroll 3
        for (int r = 0; r < PW; ++r) {
          const int L2 = Lnext;
          o += (unsigned int)pitch;
          if (r + 1 < PW) Lnext = I[o];                                                 // prefetch next row
          patch_row(r, L2);
        }
      } else {
        int Lnext = I[(size_t)reflect101(ipy - 1, rows) * pitch + colx];
#pragma unroll 3
        for (int r = 0; r < PW; ++r) {
          const int L2 = Lnext;
          if (r + 1 < PW) Lnext = I[(size_t)reflect101(ipy + r, rows) * pitch + colx];   // prefetch next row
          patch_row(r, L2);
        }
      }
    }
    __syncwarp();
    float A11, A12, A22;
    {
      // OpenCV lane c = i % 4, sequential float accumulation; 12 chains (3 sums x 4 lanes) on lanes 0..11
      float acc = 0.f;
      if (lane < 12) {
        const float* P = wP[lane >> 2];
#pragma unroll 8
        for (int i = m4; i < WIN * WIN; i += 4) acc = acc + P[i];
      }
      // (L0 + L2) + (L1 + L3)
      const float tot = (acc + __shfl_down_sync(KVFE_FULL_MASK, acc, 2)) +
                        (__shfl_down_sync(KVFE_FULL_MASK, acc, 1) + __shfl_down_sync(KVFE_FULL_MASK, acc, 3));
      A11 = __shfl_sync(KVFE_FULL_MASK, tot, 0) * FLT_SCALE;
      A12 = __shfl_sync(KVFE_FULL_MASK, tot, 4) * FLT_SCALE;
      A22 = __shfl_sync(KVFE_FULL_MASK, tot, 8) * FLT_SCALE;
    }
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
    if (minEig < dc.min_eig_thr || D < 1.1920929e-07f) {
      if (level == 0) status = false;
      continue;
    }
    D = 1.f / D;
    float qx = nx - halfWin, qy = ny - halfWin;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < dc.max_iter; ++j) {
      const int iqx = cv_floor(qx), iqy = cv_floor(qy);
      if (iqx < -WIN || iqx >= cols || iqy < -WIN || iqy >= rows) {
        if (level == 0) status = false;
        break;
      }
      a = qx - iqx; bb = qy - iqy;
      iw00 = cv_round((1.f - a) * (1.f - bb) * (1 << 14));
      iw01 = cv_round(a * (1.f - bb) * (1 << 14));
      iw10 = cv_round((1.f - a) * bb * (1 << 14));
      iw11 = (1 << 14) - iw00 - iw01 - iw10;
      const int colx = reflect101(iqx + min(lane, TW - 1), cols);
      const bool inner = iqy >= 0 && iqy + TW <= rows;
      float bacc = 0.f;                            // lanes 0..3: b1 chains, lanes 4..7: b2 chains
      const int m8 = lane & 7;
      int jprev, jprevR;
      // LKC_CHUNK window rows whose J pixels (rows y0+1 .. y0+LKC_CHUNK of the patch) were loaded together
      auto consume = [&](const int (&Jr)[LKC_CHUNK], const int y0) {
#pragma unroll
        for (int u = 0; u < LKC_CHUNK; ++u) {
          const int y = y0 + u;
          const int jn = Jr[u];
          const int jnR = __shfl_down_sync(KVFE_FULL_MASK, jn, 1);
          const int diff = descale(jprev * iw00 + jprevR * iw01 + jn * iw10 + jnR * iw11, 14 - 5) - wI[y * WIN + xl];
          jprev = jn; jprevR = jnR;
          const int t1 = diff * wIx[y * WIN + xl], t2 = diff * wIy[y * WIN + xl];
          // pixel pairs (x, x + 4): the Ix pair sum lands on lanes with (x & 4) == 0, the Iy pair sum
          // on lanes with (x & 4) != 0, so that ONE gather shuffle per group feeds both chain sets
          const int s1 = t1 + __shfl_down_sync(KVFE_FULL_MASK, t1, 4);
          const int s2 = t2 + __shfl_up_sync(KVFE_FULL_MASK, t2, 4);
          const float g = (float)((lane & 4) ? s2 : s1);
#pragma unroll
          for (int q = 0; q < WIN / 8; ++q) bacc = bacc + __shfl_sync(KVFE_FULL_MASK, g, m8 + 8 * q);
        }
      };
      if (inner) {
        // patch rows inside the image: running 32-bit offset from the (warp-uniform) level base
        unsigned int o = (unsigned int)iqy * (unsigned int)pitch + (unsigned int)colx;
        jprev = Jimg[o];
        jprevR = __shfl_down_sync(KVFE_FULL_MASK, jprev, 1);
#pragma unroll 1
        for (int y0 = 0; y0 < WIN; y0 += LKC_CHUNK) {
          int Jr[LKC_CHUNK];
#pragma unroll
          for (int u = 0; u < LKC_CHUNK; ++u) { o += (unsigned int)pitch; Jr[u] = Jimg[o]; }
          consume(Jr, y0);
        }
      } else {
        jprev = Jimg[(size_t)reflect101(iqy, rows) * pitch + colx];
        jprevR = __shfl_down_sync(KVFE_FULL_MASK, jprev, 1);
#pragma unroll 1
        for (int y0 = 0; y0 < WIN; y0 += LKC_CHUNK) {
          int Jr[LKC_CHUNK];
#pragma unroll
          for (int u = 0; u < LKC_CHUNK; ++u)
            Jr[u] = Jimg[(size_t)reflect101(iqy + y0 + u + 1, rows) * pitch + colx];
          consume(Jr, y0);
        }
      }
      float b1, b2;
      {
        // per sum: (c0 + c2) + (c1 + c3)
        const float tot = (bacc + __shfl_down_sync(KVFE_FULL_MASK, bacc, 2)) +
                          (__shfl_down_sync(KVFE_FULL_MASK, bacc, 1) + __shfl_down_sync(KVFE_FULL_MASK, bacc, 3));
        b1 = __shfl_sync(KVFE_FULL_MASK, tot, 0) * FLT_SCALE;
        b2 = __shfl_sync(KVFE_FULL_MASK, tot, 4) * FLT_SCALE;
      }
      float dxv = (float)((A12 * b2 - A22 * b1) * D);
      float dyv = (float)((A12 * b1 - A11 * b2) * D);
      qx += dxv; qy += dyv;
      nx = qx + halfWin; ny = qy + halfWin;
      if ((double)dxv * (double)dxv + (double)dyv * (double)dyv <= (double)dc.eps2) break;
      if (j > 0 && fabsf(dxv + pdx) < 0.01 && fabsf(dyv + pdy) < 0.01) {
        nx -= dxv * 0.5f; ny -= dyv * 0.5f;
        break;
      }
      pdx = dxv; pdy = dyv;
    }
  }
  if (status) {
    // the error pass of calcOpticalFlowPyrLK (level 0, `err` requested as the reference does,
    // Tracker.cpp:137-146): a FINAL position whose window origin left the image clears the status
    // (the iteration loop only tests the position it starts an iteration from)
    const int fx = cv_floor(nx - halfWin), fy = cv_floor(ny - halfWin);
    if (fx < -WIN || fx >= dc.lvl_w[0] || fy < -WIN || fy >= dc.lvl_h[0]) status = false;
  }
  if (lane == 0) {
    db.lk_qx[gi] = nx; db.lk_qy[gi] = ny;
    db.lk_status[gi] = status ? 1 : 0;
  }
}

int launch_lk(const DevCfg& dc, const DevBuf& db, int prev_slot, int cur_slot, cudaStream_t s) {
  dim3 gridc((dc.cap + LKC_WARPS - 1) / LKC_WARPS, dc.B);
  if (dc.win == 24) { lk_kernel_col<24><<<gridc, LKC_WARPS * 32, 0, s>>>(dc, db, prev_slot, cur_slot); return 1; }
  dim3 grid((dc.cap + LK_WARPS - 1) / LK_WARPS, dc.B);
  size_t sm = LK_WARPS * lk_warp_bytes(dc.win);
  static size_t attr = 0;
  if (sm > 48 * 1024 && sm > attr) {
    cudaFuncSetAttribute(lk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr = sm;
  }
  lk_kernel<<<dim3((dc.cap + LK_WARPS - 1) / LK_WARPS, dc.B), LK_WARPS * 32, sm, s>>>(dc, db, prev_slot, cur_slot);
  return 1;
}
