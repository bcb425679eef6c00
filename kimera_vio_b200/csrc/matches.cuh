// matches.cuh -- CTA-wide match bookkeeping shared by the RANSAC and FSM kernels:
// Tracker::findMatchingKeypoints / findMatchingStereoKeypoints (reference src/frontend/Tracker.cpp:919-989)
// and Tracker::computeMedianDisparity (:991-1018).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// match bookkeeping (CTA-wide helpers)
// ------------------------------------------------------------------------------------------------
// Tracker::findMatchingKeypoints: pairs (ref idx, cur idx) in cur order; optional stereo filter.
static __device__ int block_find_matches(const DevCfg& dc, const DevBuf& db, int fs_ref, int fs_cur, bool stereo,
                                  int* m_ref, int* m_cur) {
  __shared__ int s_n;
  __shared__ int wsum[32];
  __shared__ int s_chunk;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int nr = db.fr.n[fs_ref], nc = db.fr.n[fs_cur];
  const long long* lr = db.fr.lmk + (size_t)fs_ref * dc.cap;
  const long long* lc = db.fr.lmk + (size_t)fs_cur * dc.cap;
  __shared__ long long s_lr[1024];                 // reference landmark ids (searched once per current keypoint)
  if (nr <= 1024) {
    for (int j = threadIdx.x; j < nr; j += blockDim.x) s_lr[j] = lr[j];
    __syncthreads();
    lr = s_lr;
  }
  for (int base = 0; base < nc; base += blockDim.x) {
    int i = base + threadIdx.x;
    int found = -1;
    if (i < nc) {
      long long id = lc[i];
      if (id != -1)
        for (int j = nr - 1; j >= 0; --j) if (lr[j] == id) { found = j; break; }   // std::map: last wins
      if (found >= 0 && stereo) {
        if (db.fr.rstat[(size_t)fs_ref * dc.cap + found] != KVFE_KP_VALID ||
            db.fr.rstat[(size_t)fs_cur * dc.cap + i] != KVFE_KP_VALID) found = -1;
      }
    }
    int keep = found >= 0;
    unsigned bal = __ballot_sync(KVFE_FULL_MASK, keep);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0, incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(KVFE_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      wsum[lane] = incl - v;
      if (lane == 31) s_chunk = incl;
    }
    __syncthreads();
    int pos = s_n + wsum[warp] + __popc(bal & ((1u << lane) - 1));
    if (keep) { m_ref[pos] = found; m_cur[pos] = i; }
    __syncthreads();
    if (threadIdx.x == 0) s_n += s_chunk;
    __syncthreads();
  }
  return s_n;
}

// Tracker::computeMedianDisparity over matches flagged by `use` (nullptr = all): returns the
// sqrt of the element of rank size/2 of the squared displacements, or -1 when there is none.
static __device__ double block_median_disparity(const DevCfg& dc, const DevBuf& db, int fs_ref, int fs_cur,
                                         const int* m_ref, const int* m_cur, const int* use, int n, double* tmp) {
  __shared__ int s_m;
  __shared__ double s_med;
  // compact squared distances (order irrelevant for a rank statistic)
  if (threadIdx.x == 0) { s_m = 0; s_med = -1.0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (use && !use[i]) continue;
    size_t kr = (size_t)fs_ref * dc.cap + m_ref[i], kc = (size_t)fs_cur * dc.cap + m_cur[i];
    float dx = db.fr.kx[kc] - db.fr.kx[kr], dy = db.fr.ky[kc] - db.fr.ky[kr];
    float d = dx * dx + dy * dy;
    int pos = atomicAdd(&s_m, 1);
    tmp[pos] = (double)d;
  }
  __syncthreads();
  const int m = s_m;
  if (m == 0) return -1.0;
  const int center = m / 2;
  __shared__ double s_tmp[1024];
  if (m <= 1024) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_tmp[i] = tmp[i];
    __syncthreads();
    tmp = s_tmp;
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double v = tmp[i];
    int less = 0, eq_before = 0;
    for (int j = 0; j < m; ++j) {
      double w = tmp[j];
      less += (w < v);
      eq_before += (w == v && j < i);
    }
    if (less + eq_before == center) s_med = sqrt(v);
  }
  __syncthreads();
  return s_med;
}

