// matches.cuh -- CTA-wide match bookkeeping shared by the RANSAC and FSM kernels:
// Tracker::findMatchingKeypoints / findMatchingStereoKeypoints (reference src/frontend/Tracker.cpp:919-989)
// and Tracker::computeMedianDisparity (:991-1018).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// match bookkeeping (CTA-wide helpers)
// ------------------------------------------------------------------------------------------------
// Tracker::findMatchingKeypoints: pairs (ref idx, cur idx) in cur order; optional stereo filter.
//
// The reference builds a std::map landmark id -> index over the reference frame (the last index wins) and looks
// every current landmark up.  Here: the valid (id != -1) reference entries are compacted in index order; landmark
// ids grow along a frame's keypoint vector (tracked keypoints keep their order, new ones are appended with larger
// ids: FeatureDetector.cpp:141-157, Tracker.cpp:160-200), so that list is sorted and every current keypoint is
// found by binary search -- O(n log n) instead of the O(n^2) scan that cost 0.57 ms per 4K frame.  Sortedness is
// CHECKED every call; an unsorted list (never observed) takes the exhaustive scan.  `ws` is int scratch of at
// least 3 * cap + 8 entries (compacted ids as int64 pairs + indices).
static __device__ int block_find_matches(const DevCfg& dc, const DevBuf& db, int fs_ref, int fs_cur, bool stereo,
                                  int* m_ref, int* m_cur, int* ws) {
  __shared__ int s_n, s_nv, s_unsorted;
  __shared__ int wsum[32];
  __shared__ int s_chunk;
  if (threadIdx.x == 0) { s_n = 0; s_nv = 0; s_unsorted = 0; }
  __syncthreads();
  const int nr = db.fr.n[fs_ref], nc = db.fr.n[fs_cur];
  const long long* lr = db.fr.lmk + (size_t)fs_ref * dc.cap;
  const long long* lc = db.fr.lmk + (size_t)fs_cur * dc.cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  long long* vid = reinterpret_cast<long long*>(ws + ((reinterpret_cast<size_t>(ws) & 4) ? 1 : 0));   // 8-byte aligned
  int* vix = reinterpret_cast<int*>(vid + dc.cap);
  // ---- compact the valid reference entries in index order
  for (int base = 0; base < nr; base += blockDim.x) {
    const int j = base + threadIdx.x;
    const long long id = j < nr ? lr[j] : -1;
    const int keep = id != -1;
    const unsigned bal = __ballot_sync(KVFE_FULL_MASK, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = (lane < nwarp) ? wsum[lane] : 0, incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(KVFE_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      wsum[lane] = incl - v;
      if (lane == 31) s_chunk = incl;
    }
    __syncthreads();
    const int pos = s_nv + wsum[warp] + __popc(bal & ((1u << lane) - 1));
    if (keep) { vid[pos] = id; vix[pos] = j; }
    __syncthreads();
    if (threadIdx.x == 0) s_nv += s_chunk;
    __syncthreads();
  }
  const int nv = s_nv;
  for (int k = threadIdx.x; k + 1 < nv; k += blockDim.x) if (!(vid[k] < vid[k + 1])) s_unsorted = 1;
  __syncthreads();
  const bool sorted = !s_unsorted;
  // ---- look every current keypoint up
  for (int base = 0; base < nc; base += blockDim.x) {
    int i = base + threadIdx.x;
    int found = -1;
    if (i < nc) {
      long long id = lc[i];
      if (id != -1) {
        if (sorted) {
          int lo = 0, hi = nv - 1;
          while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const long long v = vid[mid];
            if (v == id) { found = vix[mid]; break; }
            if (v < id) lo = mid + 1; else hi = mid - 1;
          }
        } else {
          for (int j = nr - 1; j >= 0; --j) if (lr[j] == id) { found = j; break; }   // std::map: last wins
        }
      }
      if (found >= 0 && stereo) {
        if (db.fr.rstat[(size_t)fs_ref * dc.cap + found] != KVFE_KP_VALID ||
            db.fr.rstat[(size_t)fs_cur * dc.cap + i] != KVFE_KP_VALID) found = -1;
      }
    }
    int keep = found >= 0;
    unsigned bal = __ballot_sync(KVFE_FULL_MASK, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = (lane < nwarp) ? wsum[lane] : 0, incl = v;
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
  } else {
    // rank selection by radix on the float bits (the values are squared distances computed in f32: non-negative
    // floats order like their bit patterns; the element of rank `center` is a value, ties are indistinguishable)
    __shared__ int s_hist[2048];
    __shared__ unsigned int s_prefix;
    __shared__ int s_rank;
    if (threadIdx.x == 0) { s_prefix = 0u; s_rank = center; }
    const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
    unsigned int mask_hi = 0u;
    for (int pass = 0; pass < 3; ++pass) {
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_hist[i] = 0;
      __syncthreads();
      const unsigned int pre = s_prefix;
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned int u = __float_as_uint((float)tmp[i]);
        if ((u & mask_hi) == pre) atomicAdd(&s_hist[(u >> shifts[pass]) & ((1u << widths[pass]) - 1u)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int r = s_rank, bin = 0;
        const int nb = 1 << widths[pass];
        for (; bin < nb; ++bin) { if (r < s_hist[bin]) break; r -= s_hist[bin]; }
        s_rank = r;
        s_prefix = pre | ((unsigned int)bin << shifts[pass]);
      }
      mask_hi |= ((1u << widths[pass]) - 1u) << shifts[pass];
      __syncthreads();
    }
    if (threadIdx.x == 0) s_med = sqrt((double)__uint_as_float(s_prefix));
  }
  __syncthreads();
  return s_med;
}

