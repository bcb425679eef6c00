// select.cu -- rows a4, a5, a6.
//   anms_kernel    AdaptiveNonMaximumSuppression::suppressNonMax + binning
//                  (reference src/frontend/feature-detector/NonMaximumSuppression.cpp:33-169,
//                  anms::TopN anms/anms.cpp:37-48).  The "sort by int(response)" is cv::sortIdx on
//                  all-equal keys (GFTT responses < 1) == libstdc++ introsort scramble, which only
//                  depends on N: the permutation table is produced on the host by the same
//                  std::sort and looked up here.
//   subpix_kernel  cv::cornerSubPix (FeatureDetector.cpp:283-296): one warp per corner, 23x23
//                  cv::getRectSubPix patch in shared memory, Gaussian-weighted 2x2 normal equations
//                  accumulated in f64.
//   append_kernel  FeatureDetector::featureDetection(Frame*, R) bookkeeping (FeatureDetector.cpp:
//                  129-152): landmark ids, age 1, bearing vectors.
#include "common.cuh"
#include "subpix.cuh"

__global__ void __launch_bounds__(1024) anms_kernel(DevCfg dc, DevBuf db, int mode_mask) {
  const int b = blockIdx.x;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  __shared__ unsigned char bin_of[4096];
  __shared__ int s_cnt;
  const int N = db.corner_n[b];
  const int need = s.need;
  const int W = dc.W, H = dc.H;
  const int* corner = db.corner_idx + (size_t)b * dc.max_before_anms;
  float* ox = db.new_x + (size_t)b * dc.cap;
  float* oy = db.new_y + (size_t)b * dc.cap;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (N == 0) { if (threadIdx.x == 0) db.new_n[b] = 0; return; }
  const unsigned short* perm = db.sort_perm + (size_t)N * (N - 1) / 2;
  int out_n = 0;
  if (!dc.nms_enabled) {
    out_n = min(N, dc.cap);
    for (int i = threadIdx.x; i < out_n; i += blockDim.x) {
      int idx = corner[i];
      ox[i] = (float)(idx % W); oy[i] = (float)(idx / W);
    }
  } else if (dc.nms_type == 0) {               // TopN receives the UNSORTED list
    out_n = (need > N) ? N : need;
    out_n = min(out_n, dc.cap);
    for (int i = threadIdx.x; i < out_n; i += blockDim.x) {
      int idx = corner[i];
      ox[i] = (float)(idx % W); oy[i] = (float)(idx / W);
    }
  } else {                                     // Binning on the scramble-sorted list
    if (need > N) {
      out_n = min(N, dc.cap);
      for (int i = threadIdx.x; i < out_n; i += blockDim.x) {
        int idx = corner[perm[i]];
        ox[i] = (float)(idx % W); oy[i] = (float)(idx / W);
      }
    } else {
      const float binRow = (float)H / (float)dc.vbins, binCol = (float)W / (float)dc.hbins;
      const int per_bin = (int)roundf((float)need / (float)dc.n_active_bins);
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        int idx = corner[perm[i]];
        float x = (float)(idx % W), y = (float)(idx / W);
        int r = (int)(y / binRow), c = (int)(x / binCol);
        int bi = r * dc.hbins + c;
        bin_of[i] = (r < dc.vbins && c < dc.hbins && dc.bin_mask[bi]) ? (unsigned char)bi : 255;
      }
      __syncthreads();
      // keep element i iff its bin is active and fewer than per_bin earlier elements share the bin;
      // output order = list order of the kept ones.
      for (int base = 0; base < N; base += blockDim.x) {
        int i = base + threadIdx.x;
        int keep = 0;
        if (i < N && bin_of[i] != 255) {
          int ord = 0;
          unsigned char mine = bin_of[i];
          for (int j = 0; j < i; ++j) ord += (bin_of[j] == mine);
          keep = ord < per_bin;
        }
        // ordered compaction: ballot + warp prefix
        unsigned bal = __ballot_sync(KVFE_FULL_MASK, keep);
        __shared__ int wsum[32];
        __shared__ int s_chunk;
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
          int v = wsum[lane], incl = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(KVFE_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
          }
          wsum[lane] = incl - v;
          if (lane == 31) s_chunk = incl;
        }
        __syncthreads();
        int pos = s_cnt + wsum[warp] + __popc(bal & ((1u << lane) - 1));
        if (keep && pos < dc.cap) {
          int idx = corner[perm[i]];
          ox[pos] = (float)(idx % W); oy[pos] = (float)(idx / W);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_cnt += s_chunk;
        __syncthreads();
      }
      out_n = min(s_cnt, dc.cap);
    }
  }
  if (threadIdx.x == 0) db.new_n[b] = out_n;
}

// ------------------------------------------------------------------------------------------------
// cv::cornerSubPix
// ------------------------------------------------------------------------------------------------
// grid (ceil(cap/4), B), block 128: one warp per new corner
__global__ void __launch_bounds__(128) subpix_kernel(DevCfg dc, DevBuf db, const unsigned char* __restrict__ imgs,
                                                     size_t img_stride, const float* __restrict__ gmask,
                                                     int mode_mask) {
  __shared__ float bufs[4][SUBPIX_PATCH];
  const int b = blockIdx.y;
  if (!mode_on(db.st[b].mode, mode_mask)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= db.new_n[b]) return;
  const unsigned char* img = imgs + (size_t)b * img_stride;
  float x = db.new_x[(size_t)b * dc.cap + i], y = db.new_y[(size_t)b * dc.cap + i];
  corner_subpix_warp(img, dc.pitch, dc.W, dc.H, dc.subpix_win, dc.subpix_iters, dc.subpix_eps2, gmask,
                     bufs[warp], &x, &y, lane);
  if (lane == 0) {
    db.new_x[(size_t)b * dc.cap + i] = x;
    db.new_y[(size_t)b * dc.cap + i] = y;
  }
}

// ------------------------------------------------------------------------------------------------
// append new corners to frame k (FeatureDetector.cpp:129-152)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) append_kernel(DevCfg dc, DevBuf db, const CamModel* __restrict__ cams,
                                                     int mode_mask) {
  const int b = blockIdx.x;
  StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask)) return;
  const int fs = b * 3 + s.slot_k;
  const int n0 = db.fr.n[fs];
  int nn = db.new_n[b];
  if (n0 + nn > dc.cap) nn = dc.cap - n0;
  const long long id0 = s.lmk_next;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) {
    size_t k = (size_t)fs * dc.cap + n0 + i;
    float x = db.new_x[(size_t)b * dc.cap + i], y = db.new_y[(size_t)b * dc.cap + i];
    db.fr.kx[k] = x; db.fr.ky[k] = y;
    db.fr.lmk[k] = id0 + i;
    db.fr.age[k] = 1;
    float ux, uy;
    undistort_point(cams[0], x, y, 1, &ux, &uy);
    double v0 = (double)ux, v1 = (double)uy, v2 = 1.0;
    double n2 = v0 * v0 + (v1 * v1 + v2 * v2);
    double nrm = sqrt(n2);
    if (n2 > 0) { v0 = v0 / nrm; v1 = v1 / nrm; v2 = v2 / nrm; }
    db.fr.versor[3 * k + 0] = v0; db.fr.versor[3 * k + 1] = v1; db.fr.versor[3 * k + 2] = v2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    db.fr.n[fs] = n0 + nn;
    s.lmk_next = id0 + nn;
    s.n_new = nn;
  }
}

int launch_select(const DevCfg& dc, const DevBuf& db, const unsigned char* img, size_t img_stride,
                  const CamModel* d_cam, int mode_mask, int append, cudaStream_t s) {
  int n = 0;
  anms_kernel<<<dc.B, 1024, 0, s>>>(dc, db, mode_mask); ++n;
  if (dc.subpix_enabled) {
    subpix_kernel<<<dim3((dc.cap + 3) / 4, dc.B), 128, 0, s>>>(dc, db, img, img_stride,
                                                              db.subpix_mask, mode_mask);
    ++n;
  }
  if (append) { append_kernel<<<dc.B, 256, 0, s>>>(dc, db, d_cam, mode_mask); ++n; }
  return n;
}
