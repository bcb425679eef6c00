// ransac.cu -- rows a12, a13: geometric verification.
//   * OpenGV sample-consensus loop (opengv::sac::Ransac::computeModel as driven by Tracker::runRansac,
//     reference include/kimera-vio/frontend/Tracker.h:247-296): the sample sequence depends only on
//     (N, seed), so all hypotheses are drawn up front from the pre-computed std::mt19937(12345) /
//     uniform_int_distribution stream, evaluated in parallel (one warp per hypothesis), and the
//     sequential "best-so-far / adaptive k" logic is then replayed over the per-hypothesis inlier
//     counts -> the same final model and inlier mask as the serial loop.
//   * 2-point mono (TranslationOnlySacProblem), 3-point Arun (PointCloudSacProblem)
//     -- Tracker.cpp:213-378, :667-769.
//   * 1-point stereo voting -- Tracker.cpp:382-663, fully specified in the reference (f32 voting,
//     f64 estimate).
//   * match bookkeeping: findMatchingKeypoints / findMatchingStereoKeypoints / removeOutliers* /
//     computeMedianDisparity -- Tracker.cpp:836-1018.
#include "common.cuh"
#include "matches.cuh"
#include "fivept.cuh"

#define RS_THREADS 512

// ------------------------------------------------------------------------------------------------
// relative-pose scoring (opengv triangulation::triangulate2 + reprojection error)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm3(const double* a) { return sqrt(dot3(a, a)); }

__device__ double relpose_score(const double* R, const double* t, const double* f1, const double* f2) {
  double f2u[3];
  matvec3(R, f2, f2u);
  double b0 = dot3(t, f1), b1 = dot3(t, f2u);
  double a00 = dot3(f1, f1), a10 = dot3(f1, f2u), a01 = -a10, a11 = -dot3(f2u, f2u);
  double det = a00 * a11 - a10 * a01;
  double invdet = 1.0 / det;
  double i00 = a11 * invdet, i10 = -a10 * invdet, i01 = -a01 * invdet, i11 = a00 * invdet;
  double l0 = i00 * b0 + i01 * b1, l1 = i10 * b0 + i11 * b1;
  double X[3];
  for (int k = 0; k < 3; ++k) X[k] = (l0 * f1[k] + (t[k] + l1 * f2u[k])) / 2.0;
  // inverse transformation: R^T, -R^T t
  double tinv[3], r2[3];
  mattvec3(R, t, tinv);
  mattvec3(R, X, r2);
  // inverseSolution (3x4) * p_hom: Eigen size-4 redux = (a0 + a1) + (a2 + a3)
  for (int k = 0; k < 3; ++k)
    r2[k] = (R[k] * X[0] + R[3 + k] * X[1]) + (R[6 + k] * X[2] + (-tinv[k]) * 1.0);
  double n1 = norm3(X), n2 = norm3(r2);
  double r1[3] = {X[0] / n1, X[1] / n1, X[2] / n1};
  double rr2[3] = {r2[0] / n2, r2[1] / n2, r2[2] / n2};
  return (1.0 - dot3(f1, r1)) + (1.0 - dot3(f2, rr2));
}

// opengv relative_pose::twopt(adapter, unrotate = true, i0, i1); model = [R12 | t]
__device__ void twopt_model(const double* R12, const double* fa, const double* fb, int i0, int i1, double* model) {
  double f1p[3], f2p[3], n1[3], n2[3], t[3];
  const double* f1 = fa + 3 * i0;
  const double* f2 = fa + 3 * i1;
  matvec3(R12, fb + 3 * i0, f1p);
  matvec3(R12, fb + 3 * i1, f2p);
  cross3(f1, f1p, n1);
  cross3(f2, f2p, n2);
  cross3(n1, n2, t);
  double nt = norm3(t);
  t[0] /= nt; t[1] /= nt; t[2] /= nt;
  double flow[3] = {f1[0] - f1p[0], f1[1] - f1p[1], f1[2] - f1p[2]};
  if (dot3(flow, t) < 0) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
  for (int r = 0; r < 3; ++r) {
    model[4 * r] = R12[3 * r]; model[4 * r + 1] = R12[3 * r + 1]; model[4 * r + 2] = R12[3 * r + 2];
    model[4 * r + 3] = t[r];
  }
}

// ------------------------------------------------------------------------------------------------
// Arun: optimal rotation R (p1 ~ R p2 + t) from H = sum (p2 - c2)(p1 - c1)^T.  R = V U^T with the
// det fix is the maximiser of trace(R H); computed here with Horn's quaternion method (largest
// eigenvector of the 4x4 symmetric N matrix, cyclic Jacobi).
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eig4(double A[4][4], double V[4][4]) {
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < 4; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

__device__ void arun_model(const double* pa, const double* pb, const int* idx, int n, double* model) {
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { c1[k] += pa[3 * idx[i] + k]; c2[k] += pb[3 * idx[i] + k]; }
  for (int k = 0; k < 3; ++k) { c1[k] /= n; c2[k] /= n; }
  // S = sum (p2c)(p1c)^T : S[a][b] = sum p2c[a] * p1c[b]
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; ++i) {
    double f[3], fp[3];
    for (int k = 0; k < 3; ++k) { f[k] = pa[3 * idx[i] + k] - c1[k]; fp[k] = pb[3 * idx[i] + k] - c2[k]; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] += fp[a] * f[b];
  }
  // Horn: rotation taking p2 -> p1 maximises sum p1^T R p2; with M = sum p2 p1^T = S
  double Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2], Syx = S[1][0], Syy = S[1][1], Syz = S[1][2],
         Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
  double N[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double V[4][4];
  jacobi_eig4(N, V);
  int best = 0;
  for (int i = 1; i < 4; ++i) if (N[i][i] > N[best][best]) best = i;
  double q0 = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
  double nq = sqrt(q0 * q0 + qx * qx + qy * qy + qz * qz);
  q0 /= nq; qx /= nq; qy /= nq; qz /= nq;
  double R[9] = {q0 * q0 + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - q0 * qz), 2 * (qx * qz + q0 * qy),
                 2 * (qy * qx + q0 * qz), q0 * q0 - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - q0 * qx),
                 2 * (qz * qx - q0 * qy), 2 * (qz * qy + q0 * qx), q0 * q0 - qx * qx - qy * qy + qz * qz};
  double Rc2[3];
  matvec3(R, c2, Rc2);
  for (int r = 0; r < 3; ++r) {
    model[4 * r] = R[3 * r]; model[4 * r + 1] = R[3 * r + 1]; model[4 * r + 2] = R[3 * r + 2];
    model[4 * r + 3] = c1[r] - Rc2[r];
  }
}

__device__ __forceinline__ double cloud_score(const double* model, const double* p1, const double* p2) {
  double e[3];
  for (int r = 0; r < 3; ++r) {
    double m[3] = {model[4 * r], model[4 * r + 1], model[4 * r + 2]};
    e[r] = p1[r] - (dot3(m, p2) + model[4 * r + 3]);
  }
  return norm3(e);
}

// ------------------------------------------------------------------------------------------------
// generic sample-consensus CTA routine. problem: 0 = 2-pt (needs R12), 1 = 3-pt Arun, 2 = 5-pt Nister.
// a, b: n x 3 doubles.  Workspace (ints): shuffled[n] + samples[NS*ssz] + counts[NS];
// models: NS x 12 doubles.  Returns success; writes best model, inlier flags (0/1) and *n_inl.
// ------------------------------------------------------------------------------------------------
struct SacResult { int success; int n_inl; int iterations; };

template <int problem>
__device__ SacResult sac_run(const double* a, const double* b, int n, const double* R12,
                             double threshold, int max_it, double prob, const int* __restrict__ rnd, int rnd_n,
                             int* wi, double* models, double* best_model, int* inl_flag) {
  __shared__ SacResult res;
  __shared__ int s_best, s_done, s_iter, s_nbest, s_j0;
  __shared__ double s_k;
  __shared__ int s_counts[RS_THREADS / 32];
  __shared__ int s_valid[RS_THREADS / 32];
  __shared__ int s_skipped;
  const int ssz = problem == 0 ? 2 : (problem == 1 ? 3 : 8);
  // the 5-point solver can fail (no real root): such draws are skipped without counting as an
  // iteration (opengv Ransac::computeModel), so a few more samples than iterations are drawn
  const int NS = (problem == 2) ? 2 * (max_it + 1) : max_it + 1;
  __shared__ int s_shuffled[2048];
  __shared__ int s_samples[3 * 256];
  int* shuffled = (n <= 2048) ? s_shuffled : wi;
  int* samples = (NS * ssz <= 3 * 256) ? s_samples : wi + n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const bool enough = n >= ssz;
  if (tid == 0) { s_best = -1; s_done = enough ? 0 : 1; s_iter = 0; s_nbest = -2147483647; s_k = 1.0; s_j0 = 0; s_skipped = 0; }
  if (enough) {
    for (int i = tid; i < n; i += blockDim.x) shuffled[i] = i;
    __syncthreads();
    if (tid == 0) {                      // SampleConsensusProblem::drawIndexSample, persistent shuffle
      int pos = 0;
      for (int j = 0; j < NS; ++j) {
        for (int i = 0; i < ssz; ++i) {
          int r = rnd[pos < rnd_n ? pos : rnd_n - 1]; ++pos;
          int o = i + (r % (n - i));
          int tmp = shuffled[i]; shuffled[i] = shuffled[o]; shuffled[o] = tmp;
        }
        for (int i = 0; i < ssz; ++i) samples[j * ssz + i] = shuffled[i];
      }
    }
  }
  __syncthreads();
  // hypotheses are evaluated one chunk (one per warp) at a time and the sequential
  // "best-so-far / adaptive k" loop is replayed over each chunk; evaluation stops as soon as the
  // serial loop would have stopped (typically after the first chunk).
  while (!s_done) {
    const int j0 = s_j0;
    const int j = j0 + warp;
    double* m = models + 12 * (size_t)warp;     // chunk-local model slots
    int cnt = 0;
    if (j < NS) {
      int ok = 1;
      if (lane == 0) {
        if (problem == 0) twopt_model(R12, a, b, samples[j * 2], samples[j * 2 + 1], m);
        else if (problem == 1) arun_model(a, b, samples + j * 3, 3, m);
        else ok = fivept::solve(a, b, samples + j * 8, m) ? 1 : 0;
        s_valid[warp] = ok;
      }
      __syncwarp();
      ok = s_valid[warp];
      for (int i = lane; ok && i < n; i += 32) {
        double sc;
        if (problem != 1) {
          double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
          double t[3] = {m[3], m[7], m[11]};
          sc = relpose_score(R, t, a + 3 * i, b + 3 * i);
        } else sc = cloud_score(m, a + 3 * i, b + 3 * i);
        cnt += (sc < threshold) ? 1 : 0;
      }
      cnt = warp_sum_i(cnt);
      if (lane == 0) s_counts[warp] = cnt;
    }
    __syncthreads();
    if (tid == 0) {                        // replay of the sequential loop over this chunk
      int iterations = s_iter, n_best = s_nbest, best = s_best;
      double k = s_k;
      int done = 0;
      int skipped = s_skipped;
      for (int w = 0; w < nw; ++w) {
        if (!((double)iterations < k) || j0 + w >= NS || skipped >= 10 * max_it) { done = 1; break; }
        if (!s_valid[w]) { ++skipped; continue; }
        int c = s_counts[w];
        if (c > n_best) {
          n_best = c; best = j0 + w;
          for (int q = 0; q < 12; ++q) best_model[q] = models[12 * (size_t)w + q];
          double wr = (double)n_best / (double)n;
          double p_no = 1.0 - pow(wr, (double)ssz);
          p_no = fmax(2.220446049250313e-16, p_no);
          p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
          k = log(1.0 - prob) / log(p_no);
        }
        ++iterations;
        if (iterations > max_it) { done = 1; break; }
      }
      if (!done && (!((double)iterations < k) || j0 + nw >= NS)) done = 1;
      s_skipped = skipped;
      s_iter = iterations; s_nbest = n_best; s_best = best; s_k = k; s_done = done; s_j0 = j0 + nw;
    }
    __syncthreads();
  }
  if (tid == 0) { res.iterations = s_iter; res.success = s_best >= 0; }
  __syncthreads();
  int cnt_local = 0;
  if (s_best >= 0) {
    const double* m = best_model;
    for (int i = tid; i < n; i += blockDim.x) {
      double sc;
      if (problem != 1) {
        double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
        double t[3] = {m[3], m[7], m[11]};
        sc = relpose_score(R, t, a + 3 * i, b + 3 * i);
      } else sc = cloud_score(m, a + 3 * i, b + 3 * i);
      int f = sc < threshold ? 1 : 0;
      inl_flag[i] = f;
      cnt_local += f;
    }
  } else {
    for (int i = tid; i < n; i += blockDim.x) inl_flag[i] = 0;
  }
  __shared__ int s_cnt;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (cnt_local) atomicAdd(&s_cnt, cnt_local);
  __syncthreads();
  if (tid == 0) {
    res.n_inl = s_cnt;
    // Tracker::runRansac: success && iterations >= max_iterations && inliers.empty() -> failure
    if (res.success && res.iterations >= max_it && s_cnt == 0) res.success = 0;
    if (!res.success) res.n_inl = 0;
  }
  __syncthreads();
  SacResult out = res;
  if (!out.success) {
    for (int i = tid; i < n; i += blockDim.x) inl_flag[i] = 0;
    for (int i = tid; i < 12; i += blockDim.x) best_model[i] = (i == 0 || i == 5 || i == 10) ? 1.0 : 0.0;
  }
  __syncthreads();
  return out;
}

// layout of the per-stream double workspace rs_d: [a: cap*3][b: cap*3][models: (max_it+1)*12][tmp: cap*16]
__device__ __forceinline__ double* ws_a(const DevCfg& dc, const DevBuf& db, int b) { return db.rs_d + (size_t)b * db.rs_stride; }
__device__ __forceinline__ double* ws_b(const DevCfg& dc, const DevBuf& db, int b) { return ws_a(dc, db, b) + 3 * dc.cap; }
__device__ __forceinline__ double* ws_models(const DevCfg& dc, const DevBuf& db, int b) { return ws_b(dc, db, b) + 3 * dc.cap; }
__device__ __forceinline__ double* ws_tmp(const DevCfg& dc, const DevBuf& db, int b) {
  return ws_models(dc, db, b) + 12 * (size_t)(dc.ransac_iters + 1 > 32 ? dc.ransac_iters + 1 : 32);
}

// outlierRejectionMono (VisionImuFrontend.cpp:90-113) -> geometricOutlierRejection2d2d(Frame*, Frame*, Pose3)
template <int PROBLEM>
__device__ void mono_ransac_body(const DevCfg& dc, const DevBuf& db, const int b) {
  StreamState& s = db.st[b];
  const int fs_ref = b * 3 + s.slot_lkf, fs_cur = b * 3 + s.slot_k;
  int* m_ref = db.m_ref + (size_t)b * dc.cap;
  int* m_cur = db.m_cur + (size_t)b * dc.cap;
  int* inl = db.inl + (size_t)b * dc.cap;
  __shared__ double R12[9];
  __shared__ double model[12];
  const int n = block_find_matches(dc, db, fs_ref, fs_cur, false, m_ref, m_cur, db.scratch_i + (size_t)b * db.scratch_stride);
  if (threadIdx.x == 0) { s.nr_mono_put = n; s.nr_mono_inl = 0; }
  if (n == 0) {                       // Tracker.cpp:336-340
    if (threadIdx.x == 0) s.mono_status = KVFE_TRK_INVALID;
    return;
  }
  double* a = ws_a(dc, db, b);
  double* bb = ws_b(dc, db, b);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int k = 0; k < 3; ++k) {
      a[3 * i + k] = db.fr.versor[3 * ((size_t)fs_ref * dc.cap + m_ref[i]) + k];
      bb[3 * i + k] = db.fr.versor[3 * ((size_t)fs_cur * dc.cap + m_cur[i]) + k];
    }
  // used by the 2-point problem only; without a usable IMU rotation the reference passes Pose3() (exact identity)
  if (threadIdx.x < 9) R12[threadIdx.x] = s.given_rot ? s.kf_R_cur[threadIdx.x] : ((threadIdx.x & 3) == 0 ? 1.0 : 0.0);
  __syncthreads();
  int* wi = db.scratch_i + (size_t)b * db.scratch_stride;
  SacResult r = sac_run<PROBLEM>(a, bb, n, R12, dc.thr_mono, dc.ransac_iters, dc.ransac_prob,
                                 db.rnd_table, db.rnd_n, wi, ws_models(dc, db, b), model, inl);
  int status;
  if (!r.success) status = KVFE_TRK_INVALID;
  else status = (r.n_inl < dc.min_mono_inl) ? KVFE_TRK_FEW_MATCHES : KVFE_TRK_VALID;
  if (status != KVFE_TRK_FEW_MATCHES) {           // removeOutliersMono
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (!inl[i]) {
        db.fr.lmk[(size_t)fs_ref * dc.cap + m_ref[i]] = -1;
        db.fr.lmk[(size_t)fs_cur * dc.cap + m_cur[i]] = -1;
      }
  }
  __syncthreads();
  if (status == KVFE_TRK_VALID) {
    double med = block_median_disparity(dc, db, fs_ref, fs_cur, m_ref, m_cur, inl, n, ws_tmp(dc, db, b));
    if (med >= 0.0 && med < dc.disparity_thr) status = KVFE_TRK_LOW_DISPARITY;
  }
  if (threadIdx.x == 0) {
    s.mono_status = status;
    s.nr_mono_inl = r.n_inl;
    // tracker_status_summary_.lkf_T_k_mono_ only updated when VALID (StereoVisionImuFrontend.cpp:358-360)
    if (status == KVFE_TRK_VALID) for (int i = 0; i < 12; ++i) s.pose_mono[i] = model[i];
  }
}

// VisionImuFrontend::outlierRejectionMono (src/frontend/VisionImuFrontend.cpp:90-113) calls
// geometricOutlierRejection2d2d with the IMU rotation when it is usable (keyframe_R_cur != identity) and with
// the DEFAULT pose otherwise (Tracker.h:97-100: cam_lkf_Pose_cam_kf = gtsam::Pose3()).  The solver itself is
// chosen inside that function from the PARAMETER alone (Tracker.cpp:248-276): ransac_use_2point_mono -> the
// 2-point problem with R12 = that pose's rotation -- the identity in the second case, whatever the "5-point
// RANSAC" comment at VisionImuFrontend.cpp:108 says -- else 5-point Nister.  (The stereo side is different:
// outlierRejectionStereo really switches between two functions, see stereo_ransac_kernel.)
__global__ void __launch_bounds__(RS_THREADS) mono_ransac_kernel(DevCfg dc, DevBuf db, int mode_mask) {
  const int b = blockIdx.x;
  const StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask) || !dc.use_ransac) return;
  if (dc.use_2pt) mono_ransac_body<0>(dc, db, b);
  else mono_ransac_body<2>(dc, db, b);
}

// Eigen::Matrix3d::inverse() (cofactor formula, compute_inverse_size3_helper)
__device__ void inv3_eigen(const double* m, double* r) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[7] * m[2] - m[8] * m[1];     // cofactor<1,0>: rows (2,0), cols (1,2)
  double c20 = m[1] * m[5] - m[2] * m[4];
  double det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
  double id = 1.0 / det;
  r[0] = c00 * id; r[1] = c10 * id; r[2] = c20 * id;
  r[3] = (m[5] * m[6] - m[3] * m[8]) * id;   // cofactor<0,1>
  r[4] = (m[8] * m[0] - m[6] * m[2]) * id;   // cofactor<1,1>
  r[5] = (m[2] * m[3] - m[0] * m[5]) * id;   // cofactor<2,1>
  r[6] = (m[3] * m[7] - m[4] * m[6]) * id;   // cofactor<0,2>
  r[7] = (m[6] * m[1] - m[7] * m[0]) * id;   // cofactor<1,2>
  r[8] = (m[0] * m[4] - m[1] * m[3]) * id;   // cofactor<2,2>
}

// gtsam::StereoCamera(Pose3(), K).backproject2 Jacobian wrt (uL, uR, v) and J Sigma J^T (Sigma = I)
__device__ void point_cov(const DevCfg& dc, float uLf, float uRf, float vf, const double* Rm, double* cov) {
  double uL = (double)uLf, uR = (double)uRf, v = (double)vf;
  double d = uL - uR;
  double z = dc.baseline * dc.fx / d;
  double x = z * (uL - dc.cxr) / dc.fx, y = z * (v - dc.cyr) / dc.fy;
  double zp = z / d, xp = x / d, yp = y / d;
  double J[9] = {-xp + z / dc.fx, xp, 0.0, -yp, yp, z / dc.fy, -zp, zp, 0.0};
  if (Rm) { double T[9]; matmul3(Rm, J, T); for (int i = 0; i < 9; ++i) J[i] = T[i]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[3 * i + j] = dot3(J + 3 * i, J + 3 * j);
}

__device__ __forceinline__ float maha_f32(const float* vi, const float* Oi, const float* vj, const float* Oj) {
  float v0 = vi[0] - vj[0], v1 = vi[1] - vj[1], v2 = vi[2] - vj[2];
  float O00 = Oi[0] + Oj[0], O01 = Oi[1] + Oj[1], O02 = Oi[2] + Oj[2];
  float O10 = Oi[3] + Oj[3], O11 = Oi[4] + Oj[4], O12 = Oi[5] + Oj[5];
  float O20 = Oi[6] + Oj[6], O21 = Oi[7] + Oj[7], O22 = Oi[8] + Oj[8];
  float dinv = 1 / (O00 * (O11 * O22 - O12 * O21) - O10 * (O01 * O22 - O02 * O21) + O20 * (O01 * O12 - O11 * O02));
  return dinv * v0 * (v0 * (O11 * O22 - O12 * O21) - v1 * (O01 * O22 - O02 * O21) + v2 * (O01 * O12 - O11 * O02)) +
         dinv * v1 * (O00 * (v1 * O22 - O12 * v2) - O10 * (v0 * O22 - O02 * v2) + O20 * (v0 * O12 - v1 * O02)) +
         dinv * v2 * (O00 * (O11 * v2 - v1 * O21) - O10 * (O01 * v2 - v0 * O21) + O20 * (O01 * v1 - O11 * v0));
}

// 1-point voting on n matches.  rel (n*3 f64), cov (n*9 f64) prepared by the caller; relf/covf f32
// copies; returns status, writes inlier flags, pose [R|t], info.
__device__ int voting_1pt(const DevCfg& dc, int n, const double* rel, double* cov, const float* relf,
                          const float* covf, const double* R, int* sizes, int* inl, double* pose, double* info,
                          int* n_inl_out) {
  __shared__ int s_maxid, s_maxsize, s_ninl;
  __shared__ double acc[12];
  const float thr = (float)dc.thr_stereo;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sizes[i] = 1;    // coherent with itself
  __syncthreads();
  // every unordered pair once (the reference evaluates (i, j), i < j, and credits both sets): one warp per row i,
  // the lanes run over j > i (coalesced reads of the j side, no index decoding); the counts are order-free
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < n - 1; i += nw) {
      int ci = 0;
      for (int j = i + 1 + lane; j < n; j += 32) {
        const float m = maha_f32(relf + 3 * i, covf + 9 * i, relf + 3 * j, covf + 9 * j);
        if (m < thr) { ++ci; atomicAdd(&sizes[j], 1); }
      }
      ci = warp_sum_i(ci);
      if (lane == 0 && ci) atomicAdd(&sizes[i], ci);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {                       // first maximum (strict >), warp arg-max
    int ms = 0, mi = 0;
    for (int i = threadIdx.x; i < n; i += 32) if (sizes[i] > ms) { ms = sizes[i]; mi = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      int os = __shfl_xor_sync(KVFE_FULL_MASK, ms, o), oi = __shfl_xor_sync(KVFE_FULL_MASK, mi, o);
      if (os > ms || (os == ms && oi < mi)) { ms = os; mi = oi; }
    }
    if (threadIdx.x == 0) { s_maxid = mi; s_maxsize = ms; s_ninl = 0; }
  }
  __syncthreads();
  if (s_maxsize < 2) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) inl[i] = 0;
    if (threadIdx.x < 12) pose[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
    if (threadIdx.x < 9) info[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) *n_inl_out = 0;
    __syncthreads();
    return KVFE_TRK_INVALID;
  }
  const int id = s_maxid;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int f = 1;
    if (i != id) {
      int lo = i < id ? i : id, hi = i < id ? id : i;
      f = maha_f32(relf + 3 * lo, covf + 9 * lo, relf + 3 * hi, covf + 9 * hi) < thr ? 1 : 0;
    }
    inl[i] = f;
    if (f) {                                    // information matrix in place of the covariance
      double im[9];
      inv3_eigen(cov + 9 * i, im);
      for (int q = 0; q < 9; ++q) cov[9 * i + q] = im[q];
    }
  }
  __syncthreads();
  // translation = (sum info)^-1 sum info * rel over inliers in ascending order; 12 lanes, one
  // accumulator component each, sequential order like the reference loop (Tracker.cpp:588-596)
  if (threadIdx.x < 12) {
    double sacc = 0.0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
      if (!inl[i]) continue;
      ++cnt;
      const double* im = cov + 9 * i;
      double term;
      if (threadIdx.x < 3) term = dot3(im + 3 * threadIdx.x, rel + 3 * i);
      else term = im[threadIdx.x - 3];
      sacc = sacc + term;
    }
    acc[threadIdx.x] = sacc;
    if (threadIdx.x == 0) s_ninl = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ti[9], t[3];
    inv3_eigen(acc + 3, ti);
    matvec3(ti, acc, t);
    for (int r = 0; r < 3; ++r) {
      pose[4 * r] = R[3 * r]; pose[4 * r + 1] = R[3 * r + 1]; pose[4 * r + 2] = R[3 * r + 2]; pose[4 * r + 3] = t[r];
    }
    for (int i = 0; i < 9; ++i) info[i] = acc[3 + i];
    *n_inl_out = s_ninl;
  }
  __syncthreads();
  return s_ninl < dc.min_stereo_inl ? KVFE_TRK_FEW_MATCHES : KVFE_TRK_VALID;
}

// outlierRejectionStereo (VisionImuFrontend.cpp:115-144)
__global__ void __launch_bounds__(RS_THREADS) stereo_ransac_kernel(DevCfg dc, DevBuf db, int mode_mask) {
  const int b = blockIdx.x;
  StreamState& s = db.st[b];
  if (!mode_on(s.mode, mode_mask) || !dc.use_ransac) return;
  if (!dc.use_stereo_tracking) { if (threadIdx.x == 0) s.stereo_status = KVFE_TRK_INVALID; return; }
  const int fs_ref = b * 3 + s.slot_lkf, fs_cur = b * 3 + s.slot_k;
  int* m_ref = db.m_ref + (size_t)b * dc.cap;
  int* m_cur = db.m_cur + (size_t)b * dc.cap;
  int* inl = db.inl + (size_t)b * dc.cap;
  __shared__ double pose[12], info[9];
  __shared__ int s_ninl;
  const int n = block_find_matches(dc, db, fs_ref, fs_cur, true, m_ref, m_cur, db.scratch_i + (size_t)b * db.scratch_stride);
  const bool one_pt = dc.use_1pt && s.given_rot;
  int status;
  if (one_pt) {
    double* rel = ws_a(dc, db, b);                    // n*3
    double* cov = ws_tmp(dc, db, b);                  // n*9 (tmp has cap*16)
    float* relf = reinterpret_cast<float*>(ws_b(dc, db, b));           // n*3 floats
    float* covf = reinterpret_cast<float*>(cov + 9 * (size_t)dc.cap);  // tail of tmp: 7*cap doubles >= 9*cap floats
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      size_t kr = (size_t)fs_ref * dc.cap + m_ref[i], kc = (size_t)fs_cur * dc.cap + m_cur[i];
      double cr[9], cc[9], pc[3];
      point_cov(dc, db.fr.lrx[kr], db.fr.rrx[kr], db.fr.lry[kr], nullptr, cr);
      point_cov(dc, db.fr.lrx[kc], db.fr.rrx[kc], db.fr.lry[kc], s.kf_R_cur, cc);
      matvec3(s.kf_R_cur, db.fr.p3d + 3 * kc, pc);
      for (int k = 0; k < 3; ++k) {
        double v = db.fr.p3d[3 * kr + k] - pc[k];
        rel[3 * i + k] = v; relf[3 * i + k] = (float)v;
      }
      for (int k = 0; k < 9; ++k) { double m = cc[k] + cr[k]; cov[9 * i + k] = m; covf[9 * i + k] = (float)m; }
    }
    __syncthreads();
    int* sizes = db.scratch_i + (size_t)b * db.scratch_stride;
    status = voting_1pt(dc, n, rel, cov, relf, covf, s.kf_R_cur, sizes, inl, pose, info, &s_ninl);
    __syncthreads();
    // removeOutliersStereo is applied unconditionally on this path (Tracker.cpp:658-660)
  } else {
    double* a = ws_a(dc, db, b);
    double* bb = ws_b(dc, db, b);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      for (int k = 0; k < 3; ++k) {
        a[3 * i + k] = db.fr.p3d[3 * ((size_t)fs_ref * dc.cap + m_ref[i]) + k];
        bb[3 * i + k] = db.fr.p3d[3 * ((size_t)fs_cur * dc.cap + m_cur[i]) + k];
      }
    __syncthreads();
    int* wi = db.scratch_i + (size_t)b * db.scratch_stride;
    SacResult r = sac_run<1>(a, bb, n, nullptr, dc.thr_stereo, dc.ransac_iters, dc.ransac_prob, db.rnd_table,
                             db.rnd_n, wi, ws_models(dc, db, b), pose, inl);
    if (!r.success) status = KVFE_TRK_INVALID;
    else status = r.n_inl < dc.min_stereo_inl ? KVFE_TRK_FEW_MATCHES : KVFE_TRK_VALID;
    if (threadIdx.x < 9) info[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) s_ninl = r.n_inl;
    __syncthreads();
  }
  const bool remove = one_pt ? true : (status != KVFE_TRK_INVALID);
  if (remove) {
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (!inl[i]) {
        size_t kr = (size_t)fs_ref * dc.cap + m_ref[i], kc = (size_t)fs_cur * dc.cap + m_cur[i];
        db.fr.rstat[kr] = KVFE_KP_FAILED_ARUN; db.fr.depth[kr] = 0.0;
        db.fr.p3d[3 * kr] = 0; db.fr.p3d[3 * kr + 1] = 0; db.fr.p3d[3 * kr + 2] = 0;
        db.fr.rstat[kc] = KVFE_KP_FAILED_ARUN; db.fr.depth[kc] = 0.0;
        db.fr.p3d[3 * kc] = 0; db.fr.p3d[3 * kc + 1] = 0; db.fr.p3d[3 * kc + 2] = 0;
      }
  }
  if (threadIdx.x == 0) {
    s.stereo_status = status;
    s.nr_stereo_put = n; s.nr_stereo_inl = s_ninl;
    for (int i = 0; i < 9; ++i) s.info_stereo[i] = info[i];
    if (status == KVFE_TRK_VALID) for (int i = 0; i < 12; ++i) s.pose_stereo[i] = pose[i];
  }
}

// ------------------------------------------------------------------------------------------------
// stage-level ("raw") kernels: one problem, plain arrays
// ------------------------------------------------------------------------------------------------
template <int problem>
__global__ void __launch_bounds__(RS_THREADS) sac_raw_kernel(DevCfg dc, DevBuf db, const double* a,
                                                             const double* b, int n, const double* R12, double thr,
                                                             int min_inl, int* inl, int* n_inl, double* pose,
                                                             int* status) {
  __shared__ double model[12];
  __shared__ double Rs[9];
  if (threadIdx.x < 9) Rs[threadIdx.x] = R12 ? R12[threadIdx.x] : ((threadIdx.x % 4 == 0) ? 1.0 : 0.0);
  __syncthreads();
  int* wi = db.scratch_i;
  SacResult r = sac_run<problem>(a, b, n, Rs, thr, dc.ransac_iters, dc.ransac_prob, db.rnd_table, db.rnd_n, wi,
                                 ws_models(dc, db, 0), model, inl);
  if (threadIdx.x == 0) {
    *n_inl = r.n_inl;
    *status = !r.success ? KVFE_TRK_INVALID : (r.n_inl < min_inl ? KVFE_TRK_FEW_MATCHES : KVFE_TRK_VALID);
    for (int i = 0; i < 12; ++i) pose[i] = model[i];
  }
}

__global__ void __launch_bounds__(RS_THREADS) vote_raw_kernel(DevCfg dc, DevBuf db, const float* rl, const float* rr,
                                                              const float* cl, const float* cr, const double* p_ref,
                                                              const double* p_cur, int n, const double* R, int* inl,
                                                              int* n_inl, double* pose, double* info, int* status) {
  __shared__ double Rs[9], sp[12], si[9];
  __shared__ int s_n;
  if (threadIdx.x < 9) Rs[threadIdx.x] = R[threadIdx.x];
  __syncthreads();
  double* rel = ws_a(dc, db, 0);
  double* cov = ws_tmp(dc, db, 0);
  float* relf = reinterpret_cast<float*>(ws_b(dc, db, 0));
  float* covf = reinterpret_cast<float*>(cov + 9 * (size_t)dc.cap);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double c1[9], c2[9], pc[3];
    point_cov(dc, rl[2 * i], rr[2 * i], rl[2 * i + 1], nullptr, c1);
    point_cov(dc, cl[2 * i], cr[2 * i], cl[2 * i + 1], Rs, c2);
    matvec3(Rs, p_cur + 3 * i, pc);
    for (int k = 0; k < 3; ++k) { double v = p_ref[3 * i + k] - pc[k]; rel[3 * i + k] = v; relf[3 * i + k] = (float)v; }
    for (int k = 0; k < 9; ++k) { double m = c2[k] + c1[k]; cov[9 * i + k] = m; covf[9 * i + k] = (float)m; }
  }
  __syncthreads();
  int st = voting_1pt(dc, n, rel, cov, relf, covf, Rs, db.scratch_i, inl, sp, si, &s_n);
  __syncthreads();
  if (threadIdx.x == 0) {
    *status = st; *n_inl = s_n;
    for (int i = 0; i < 12; ++i) pose[i] = sp[i];
    for (int i = 0; i < 9; ++i) info[i] = si[i];
  }
}

int launch_ransac_mono(const DevCfg& dc, const DevBuf& db, int mode_mask, cudaStream_t s) {
  mono_ransac_kernel<<<dc.B, RS_THREADS, 0, s>>>(dc, db, mode_mask);
  return 1;
}
int launch_ransac_stereo(const DevCfg& dc, const DevBuf& db, int mode_mask, cudaStream_t s) {
  stereo_ransac_kernel<<<dc.B, RS_THREADS, 0, s>>>(dc, db, mode_mask);
  return 1;
}
int launch_ransac_mono_raw(const DevCfg& dc, const DevBuf& db, const double* f_ref, const double* f_cur,
                           int n, const double* R12, int use_2pt, int* inl, int* n_inl, double* pose,
                           int* status, cudaStream_t s) {
  if (use_2pt) sac_raw_kernel<0><<<1, RS_THREADS, 0, s>>>(dc, db, f_ref, f_cur, n, R12, dc.thr_mono, dc.min_mono_inl, inl,
                                                          n_inl, pose, status);
  else sac_raw_kernel<2><<<1, RS_THREADS, 0, s>>>(dc, db, f_ref, f_cur, n, R12, dc.thr_mono, dc.min_mono_inl, inl,
                                                   n_inl, pose, status);
  return 1;
}
int launch_ransac_3pt_raw(const DevCfg& dc, const DevBuf& db, const double* p_ref, const double* p_cur,
                          int n, int* inl, int* n_inl, double* pose, int* status, cudaStream_t s) {
  sac_raw_kernel<1><<<1, RS_THREADS, 0, s>>>(dc, db, p_ref, p_cur, n, nullptr, dc.thr_stereo, dc.min_stereo_inl, inl,
                                             n_inl, pose, status);
  return 1;
}
int launch_ransac_1pt_raw(const DevCfg& dc, const DevBuf& db, const float* rl, const float* rr,
                          const float* cl, const float* cr, const double* p_ref, const double* p_cur,
                          int n, const double* R, int* inl, int* n_inl, double* pose, double* info,
                          int* status, cudaStream_t s) {
  vote_raw_kernel<<<1, RS_THREADS, 0, s>>>(dc, db, rl, rr, cl, cr, p_ref, p_cur, n, R, inl, n_inl, pose, info,
                                           status);
  return 1;
}


// ---- stage-level statics of the Tracker (include/kvfe.h, "boundary completion") -------------------------
// Tracker::computeMedianDisparity (Tracker.cpp:991-1018) over frame slots 1 (ref) / 0 (cur) of stream 0 and the
// match list in m_ref / m_cur; out[0] = median, out[1] = 1 when there is at least one match
__global__ void __launch_bounds__(256) median_disparity_raw_kernel(DevCfg dc, DevBuf db, int m, double* out) {
  const double med = block_median_disparity(dc, db, 1, 0, db.m_ref, db.m_cur, nullptr, m, db.rs_d);
  if (threadIdx.x == 0) { out[0] = med < 0.0 ? 0.0 : med; out[1] = med < 0.0 ? 0.0 : 1.0; }
}
// Tracker::getPoint3AndCovariance (Tracker.cpp:772-818) with stereo_point_covariance = I (its only caller,
// Tracker.cpp:560-563): p' = R p, cov = (R J)(R J)^T, J = d backproject2 / d (uL, uR, v)
__global__ void __launch_bounds__(128) point3_cov_raw_kernel(DevCfg dc, const float* __restrict__ ul, const float* __restrict__ ur,
                                                             const float* __restrict__ v, const double* __restrict__ p3d, int n,
                                                             const double* __restrict__ Rm, double* __restrict__ op,
                                                             double* __restrict__ ocov) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double c[9];
  point_cov(dc, ul[i], ur[i], v[i], Rm, c);
  for (int k = 0; k < 9; ++k) ocov[9 * (size_t)i + k] = c[k];
  double p[3] = {p3d[3 * (size_t)i], p3d[3 * (size_t)i + 1], p3d[3 * (size_t)i + 2]};
  if (Rm) { double q[3]; matvec3(Rm, p, q); p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; }
  for (int k = 0; k < 3; ++k) op[3 * (size_t)i + k] = p[k];
}
int launch_median_disparity_raw(const DevCfg& dc, const DevBuf& db, int m, double* out, cudaStream_t s) {
  median_disparity_raw_kernel<<<1, 256, 0, s>>>(dc, db, m, out);
  return 1;
}
int launch_point3_cov_raw(const DevCfg& dc, const float* ul, const float* ur, const float* v, const double* p3d, int n,
                          const double* Rm, double* op, double* ocov, cudaStream_t s) {
  point3_cov_raw_kernel<<<(n + 127) / 128, 128, 0, s>>>(dc, ul, ur, v, p3d, n, Rm, op, ocov);
  return 1;
}
