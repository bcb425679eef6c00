// gftt.cu -- row a3: cv::goodFeaturesToTrack as called by FeatureDetector::rawFeatureDetection
// (reference src/frontend/feature-detector/FeatureDetector.cpp:165-203, GFTTDetector :71-82).
//
//   mask     : 255 image minus filled circles of radius min_distance at every keypoint with a
//              valid landmark (FeatureDetector.cpp:185-203; cv::circle raster, App. A.2)
//   response : cv::cornerMinEigenVal(blockSize 3, ksize 3, BORDER_REFLECT_101), bit-exact:
//              Sobel with OpenCV's FMA op-order, f32 products, 3-tap f64 row sums and the
//              HISTORY-DEPENDENT f64 running column sum of cv::boxFilter (one lane per column,
//              marching down the rows), min-eigenvalue formula in f32.
//   select   : max over mask, threshold (quality * max), 3x3 non-max ("== dilate"), sort by
//              (value desc, address desc), greedy min-distance on a cell grid (parallel fixed
//              point that reproduces the sequential result), stop at maxCorners.
//
// HBM traffic per keyframe image: image 1 B/px (L2 hits for the 3-row window), mask 1 B/px write
// + read, response 4 B/px write + 9 reads served by L1/L2.  Non-compulsory.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// mask
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_fill_kernel(DevCfg dc, unsigned char* __restrict__ mask,
                                                        const StreamState* __restrict__ st, int mode_mask) {
  const int b = blockIdx.y;
  if (st && !mode_on(st[b].mode, mode_mask)) return;
  uint4* p = reinterpret_cast<uint4*>(mask + (size_t)b * dc.img_stride);
  size_t n16 = dc.img_stride / 16;
  const uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// one CTA per stream; one warp per keypoint; hw[dy + r] = half width of the raster row
__global__ void __launch_bounds__(256) mask_circles_kernel(DevCfg dc, DevBuf db, const int* __restrict__ hw, int r,
                                                           int mode_mask) {
  const int b = blockIdx.x;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int n = db.fr.n[fs];
  unsigned char* m = db.mask + (size_t)b * dc.img_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < n; i += nw) {
    size_t k = (size_t)fs * dc.cap + i;
    if (db.fr.lmk[k] == -1) continue;
    // cv::Point(Point2f): saturate_cast<int>(float) == cvRound
    int cx = cv_round(db.fr.kx[k]), cy = cv_round(db.fr.ky[k]);
    for (int dy = -r; dy <= r; ++dy) {
      int y = cy + dy;
      if (y < 0 || y >= dc.H) continue;
      int h = hw[dy + r];
      int xa = max(cx - h, 0), xb = min(cx + h, dc.W - 1);
      for (int x = xa + lane; x <= xb; x += 32) m[(size_t)y * dc.pitch + x] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// response map
// ------------------------------------------------------------------------------------------------
struct RT { float r, t; };

__device__ __forceinline__ RT row_rt(const unsigned char* __restrict__ img, int pitch, int W, int y, int c,
                                     float s, float s2, bool tail) {
  const unsigned char* row = img + (size_t)y * pitch;
  float im = (float)row[reflect101(c - 1, W)], i0 = (float)row[c], ip = (float)row[reflect101(c + 1, W)];
  RT o;
  o.r = ip - im;                                   // exact
  if (!tail) o.t = fmaf(ip, s, fmaf(i0, s2, s * im));
  else o.t = (s * im + s2 * i0) + s * ip;          // host scalar loop tail: no FMA
  return o;
}

// grid (ceil(ngroups/4), B); block 128 = 4 warps, each warp owns 30 output columns (+2 halo lanes)
__global__ void __launch_bounds__(128) mineig_kernel(DevCfg dc, DevBuf db, const unsigned char* __restrict__ imgs,
                                                     size_t img_stride, int mode_mask, int use_mask) {
  const int b = blockIdx.y;
  if (!mode_on(db.st[b].mode, mode_mask)) return;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int x0 = group * 30;
  const int W = dc.W, H = dc.H;
  if (x0 >= W) return;
  const unsigned char* img = imgs + (size_t)b * img_stride;
  const unsigned char* msk = db.mask + (size_t)b * dc.img_stride;
  float* eig = db.eig + (size_t)b * W * H;
  const int cx = x0 - 1 + lane;                 // lanes 0 and 31 are halo
  int c = reflect101(cx, W);
  c = clampi(c, 0, W - 1);                      // lanes beyond the reflected border are inactive
  const bool writer = lane >= 1 && lane <= 30 && cx < W;
  const float s = (float)(1.0 / 3060.0), s2 = 2.0f * s;
  const bool tail = dc.sobel_tail_start >= 0 && c >= dc.sobel_tail_start;

  RT a = row_rt(img, dc.pitch, W, reflect101(-1, H), c, s, s2, tail);   // row p-1
  RT m = row_rt(img, dc.pitch, W, 0, c, s, s2, tail);                   // row p
  double sum0 = 0, sum1 = 0, sum2 = 0;             // running column sums (xx, xy, yy)
  double rm2_0 = 0, rm2_1 = 0, rm2_2 = 0;          // R(p-2)
  double rm1_0 = 0, rm1_1 = 0, rm1_2 = 0;          // R(p-1)
  float vmax = -INFINITY;

  for (int p = 0; p <= H; ++p) {
    double r0, r1, r2;
    if (p < H) {
      RT n = row_rt(img, dc.pitch, W, reflect101(p + 1, H), c, s, s2, tail);  // row p+1
      // cv::Sobel column pass: Dx = fma(r(y-1) + r(y+1), s, (2s) * r(y)); Dy = t(y+1) - t(y-1)
      float dx = fmaf(a.r + n.r, s, s2 * m.r);
      float dy = n.t - a.t;
      float pxx = dx * dx, pxy = dx * dy, pyy = dy * dy;
      // RowSum<float,double>, ksize 3: (S[x-1] + S[x]) + S[x+1]
      float l0 = __shfl_up_sync(KVFE_FULL_MASK, pxx, 1), g0 = __shfl_down_sync(KVFE_FULL_MASK, pxx, 1);
      float l1 = __shfl_up_sync(KVFE_FULL_MASK, pxy, 1), g1 = __shfl_down_sync(KVFE_FULL_MASK, pxy, 1);
      float l2 = __shfl_up_sync(KVFE_FULL_MASK, pyy, 1), g2 = __shfl_down_sync(KVFE_FULL_MASK, pyy, 1);
      r0 = ((double)l0 + (double)pxx) + (double)g0;
      r1 = ((double)l1 + (double)pxy) + (double)g1;
      r2 = ((double)l2 + (double)pyy) + (double)g2;
      a = m;
      m = n;
    } else {
      r0 = rm2_0; r1 = rm2_1; r2 = rm2_2;            // R(H) = R(H-2) (reflect)
    }
    if (p == 1) {                                    // ColumnSum init: SUM = (0 + R(-1)) + R(0), R(-1) = R(1)
      sum0 = (0.0 + r0) + rm1_0;
      sum1 = (0.0 + r1) + rm1_1;
      sum2 = (0.0 + r2) + rm1_2;
    }
    if (p >= 1) {
      const int y = p - 1;
      // s0 = SUM + R(y+1); out = (float)s0; SUM = s0 - R(y-1)   (R(-1) = R(1))
      double s0 = sum0 + r0, s1 = sum1 + r1, s2d = sum2 + r2;
      double o0 = (y == 0) ? r0 : rm2_0, o1 = (y == 0) ? r1 : rm2_1, o2 = (y == 0) ? r2 : rm2_2;
      sum0 = s0 - o0; sum1 = s1 - o1; sum2 = s2d - o2;
      if (writer) {
        float A = (float)s0 * 0.5f, Bv = (float)s1, C = (float)s2d * 0.5f;
        float e = (A + C) - sqrtf((A - C) * (A - C) + Bv * Bv);
        eig[(size_t)y * W + cx] = e;
        if (!use_mask || msk[(size_t)y * dc.pitch + cx]) vmax = fmaxf(vmax, e);
      }
    }
    rm2_0 = rm1_0; rm2_1 = rm1_1; rm2_2 = rm1_2;
    rm1_0 = r0; rm1_1 = r1; rm1_2 = r2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(KVFE_FULL_MASK, vmax, o));
  if (lane == 0 && vmax > -INFINITY) atomicMax(&db.eig_max[b], f2ord(vmax));
}

__global__ void gftt_init_kernel(DevCfg dc, DevBuf db, int mode_mask) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= dc.B) return;
  if (!mode_on(db.st[b].mode, mode_mask)) return;
  db.eig_max[b] = f2ord(-INFINITY);
  db.cand_n[b] = 0;
  db.corner_n[b] = 0;
}

// ------------------------------------------------------------------------------------------------
// candidates: interior pixels with eig > thr, eig == 3x3 max, mask != 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cand_kernel(DevCfg dc, DevBuf db, int mode_mask) {
  const int b = blockIdx.z;
  if (!mode_on(db.st[b].mode, mode_mask)) return;
  const int W = dc.W, H = dc.H;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < 1 || y < 1 || x > W - 2 || y > H - 2) return;
  const float* e = db.eig + (size_t)b * W * H;
  float maxv = ord2f(db.eig_max[b]);
  if (!(maxv > -INFINITY)) maxv = 0.f;              // empty mask: minMaxLoc leaves maxVal = 0
  const float thr = (float)((double)maxv * (double)dc.quality);
  const float v = e[(size_t)y * W + x];
  if (!(v > thr)) return;
  if (!db.mask[(size_t)b * dc.img_stride + (size_t)y * dc.pitch + x]) return;
  const float* r0 = e + (size_t)(y - 1) * W + x;
  const float* r1 = r0 + W;
  const float* r2 = r1 + W;
  float nb = fmaxf(fmaxf(fmaxf(r0[-1], r0[0]), fmaxf(r0[1], r1[-1])), fmaxf(fmaxf(r1[1], r2[-1]), fmaxf(r2[0], r2[1])));
  if (v < nb) return;
  int slot = atomicAdd(&db.cand_n[b], 1);
  if (slot < dc.cand_cap)
    db.cand[(size_t)b * dc.cand_cap + slot] =
        ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(y * W + x);
}

// ------------------------------------------------------------------------------------------------
// sort (descending 64-bit keys) + greedy min-distance, one CTA per stream
// ------------------------------------------------------------------------------------------------
__device__ void bitonic_desc(unsigned long long* k, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = k[lo], c = k[hi];
        if ((a < c) == desc) { k[lo] = c; k[hi] = a; }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024) sort_greedy_kernel(DevCfg dc, DevBuf db, int mode_mask, int smem_keys) {
  extern __shared__ unsigned long long skeys[];
  const int b = blockIdx.x;
  if (!mode_on(db.st[b].mode, mode_mask)) return;
  const int W = dc.W, H = dc.H;
  int n = min(db.cand_n[b], dc.cand_cap);
  unsigned long long* gk = db.cand + (size_t)b * dc.cand_cap;
  int P = 1;
  while (P < n) P <<= 1;
  if (P < 2) P = 2;
  unsigned long long* k = (P <= smem_keys) ? skeys : gk;     // gk has cand_cap >= P (cand_cap is pow2)
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long v = (i < n) ? gk[i] : 0ull;
    k[i] = v;
  }
  __syncthreads();
  bitonic_desc(k, P);

  int* sc = db.scratch_i + (size_t)b * db.scratch_stride;
  int* state = sc;                          // [cand_cap] 0 undecided, 1 accepted, 2 rejected
  int* cellof = sc + dc.cand_cap;           // [cand_cap]
  int* items = sc + 2 * dc.cand_cap;        // [cand_cap] candidate ranks grouped by cell
  int* cstart = sc + 3 * dc.cand_cap;       // [ncells + 1]
  int* corner = db.corner_idx + (size_t)b * dc.max_before_anms;
  __shared__ int s_total, s_flag;

  const int md = dc.min_distance;
  if (md < 1) {                             // no min-distance: top maxCorners
    int m = min(n, dc.max_before_anms);
    for (int i = threadIdx.x; i < m; i += blockDim.x) corner[i] = (int)(k[i] & 0xffffffffu);
    if (threadIdx.x == 0) db.corner_n[b] = m;
    return;
  }
  const int cell = md;                      // cvRound(minDistance), minDistance is an int parameter
  const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
  const int ncells = gw * gh;
  for (int i = threadIdx.x; i <= ncells; i += blockDim.x) cstart[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int idx = (int)(k[i] & 0xffffffffu);
    int y = idx / W, x = idx - y * W;
    int c = (y / cell) * gw + (x / cell);
    cellof[i] = c;
    state[i] = 0;
    atomicAdd(&cstart[c + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                   // exclusive scan over <= a few thousand cells
    int acc = 0;
    for (int c = 0; c <= ncells; ++c) { int v = cstart[c]; cstart[c] = acc; acc += v; }
    // cstart[c] now = start of cell c-1 ... shift: cstart[c+1] held count(c); after the scan
    // cstart[c+1] = sum_{j<=c-1}... see fill below (uses cstart[c+1] as the running cursor of cell c)
  }
  __syncthreads();
  // after the scan: cstart[c+1] == number of items in cells < c  == start offset of cell c
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int pos = atomicAdd(&cstart[cellof[i] + 1], 1);
    items[pos] = i;
  }
  __syncthreads();
  // now cstart[c+1] == end of cell c, and start of cell c == (c == 0 ? 0 : cstart[c]) == cstart[c]
  // because cstart[0] == 0 and cstart[c] (c >= 1) was advanced to the end of cell c-1.
  const int md2 = md * md;
  volatile int* vstate = state;
  for (int round = 0; round < 4096; ++round) {
    int pending_any = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (vstate[i] != 0) continue;
      int idx = (int)(k[i] & 0xffffffffu);
      int y = idx / W, x = idx - y * W;
      int cxl = x / cell, cyl = y / cell;
      int x1 = max(cxl - 1, 0), x2 = min(cxl + 1, gw - 1), y1 = max(cyl - 1, 0), y2 = min(cyl + 1, gh - 1);
      int verdict = 1;                       // accept unless a better conflicting one is accepted / pending
      for (int yy = y1; yy <= y2 && verdict != 2; ++yy)
        for (int xx = x1; xx <= x2 && verdict != 2; ++xx) {
          int c = yy * gw + xx;
          for (int q = cstart[c]; q < cstart[c + 1]; ++q) {
            int j = items[q];
            if (j >= i) continue;
            int sj = vstate[j];
            if (sj == 2) continue;
            int jdx = (int)(k[j] & 0xffffffffu);
            int jy = jdx / W, jx = jdx - jy * W;
            int ddx = x - jx, ddy = y - jy;
            if (ddx * ddx + ddy * ddy < md2) {
              if (sj == 1) { verdict = 2; break; }
              verdict = 0;                   // undecided better neighbour: wait
            }
          }
        }
      if (verdict == 0) pending_any = 1;
      else vstate[i] = verdict;
    }
    int any = __syncthreads_or(pending_any);
    if (!any) break;
  }
  __syncthreads();
  // compact accepted candidates in rank order, stop at maxCorners
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int acc = (i < n && state[i] == 1) ? 1 : 0;
    // block-wide exclusive scan via warp ballots
    unsigned bal = __ballot_sync(KVFE_FULL_MASK, acc);
    __shared__ int wsum[32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(KVFE_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      wsum[lane] = incl - v;
      if (lane == 31) s_flag = incl;
    }
    __syncthreads();
    int pos = s_total + wsum[warp] + __popc(bal & ((1u << lane) - 1));
    if (acc && pos < dc.max_before_anms) corner[pos] = (int)(k[i] & 0xffffffffu);
    __syncthreads();
    if (threadIdx.x == 0) s_total += s_flag;
    __syncthreads();
  }
  if (threadIdx.x == 0) db.corner_n[b] = min(s_total, dc.max_before_anms);
  (void)H;
}

int launch_min_eig(const DevCfg& dc, const DevBuf& db, const unsigned char* img, size_t img_stride,
                   int mode_mask, cudaStream_t s) {
  int n = 0;
  gftt_init_kernel<<<(dc.B + 63) / 64, 64, 0, s>>>(dc, db, mode_mask); ++n;
  int groups = (dc.W + 29) / 30;
  dim3 grid((groups + 3) / 4, dc.B);
  mineig_kernel<<<grid, 128, 0, s>>>(dc, db, img, img_stride, mode_mask, 0); ++n;
  return n;
}

int launch_gftt(const DevCfg& dc, const DevBuf& db, const unsigned char* img, size_t img_stride,
                const int* circle_hw, int circle_r, int mode_mask, cudaStream_t s) {
  int n = 0;
  gftt_init_kernel<<<(dc.B + 63) / 64, 64, 0, s>>>(dc, db, mode_mask); ++n;
  mask_fill_kernel<<<dim3(64, dc.B), 256, 0, s>>>(dc, db.mask, db.st, mode_mask); ++n;
  mask_circles_kernel<<<dc.B, 256, 0, s>>>(dc, db, circle_hw, circle_r, mode_mask); ++n;
  int groups = (dc.W + 29) / 30;
  mineig_kernel<<<dim3((groups + 3) / 4, dc.B), 128, 0, s>>>(dc, db, img, img_stride, mode_mask, 1); ++n;
  cand_kernel<<<dim3((dc.W + 31) / 32, (dc.H + 7) / 8, dc.B), 256, 0, s>>>(dc, db, mode_mask); ++n;
  int smem_keys = 16384;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(sort_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_keys * 8);
    attr_set = true;
  }
  sort_greedy_kernel<<<dc.B, 1024, smem_keys * 8, s>>>(dc, db, mode_mask, smem_keys); ++n;
  return n;
}
