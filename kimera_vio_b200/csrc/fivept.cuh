// fivept.cuh -- 5-point relative pose (Nister) for the mono RANSAC of rigs that switch the IMU-aided
// 2-point variant off (reference src/frontend/Tracker.cpp:266-276 -> opengv
// CentralRelativePoseSacProblem, algorithm NISTER: 5 points + 3 disambiguation points; only
// params/D455 among the shipped rigs).  One thread computes one hypothesis:
//   null space of the 5x9 epipolar system -> E = xX + yY + zZ + W -> the 10 cubic constraints
//   (det E = 0, 2 E E^T E - tr(E E^T) E = 0) -> Gauss-Jordan on the 10x20 coefficient matrix ->
//   Nister's 3x3 polynomial matrix B(z) -> degree-10 polynomial, real roots by Sturm bracketing +
//   bisection/Newton -> (x, y) -> essential matrices -> decomposition into 4 (R, t) each -> the
//   candidate with the smallest summed reprojection score over the 8 sample points.
// OpenGV itself is not available anywhere in this image: hypothesis-level parity is not claimed,
// the oracle (oracle/ransac.py, action-matrix eigenvalues) and this kernel agree at the level the
// reference's own tests check -- the final inlier mask and status.
#pragma once
#include "common.cuh"

__device__ double relpose_score(const double* R, const double* t, const double* f1, const double* f2);

namespace fivept {

// dense trivariate polynomial of total degree <= 3: c[i][j][k] multiplies x^i y^j z^k
struct P3 { double c[4][4][4]; };

__device__ inline void p3_zero(P3& p) { for (int i = 0; i < 64; ++i) (&p.c[0][0][0])[i] = 0.0; }
__device__ inline void p3_lin(P3& p, double x, double y, double z, double w) {
  p3_zero(p); p.c[1][0][0] = x; p.c[0][1][0] = y; p.c[0][0][1] = z; p.c[0][0][0] = w;
}
__device__ inline void p3_mul_acc(P3& o, const P3& a, const P3& b, double s) {   // o += s * a * b (degree <= 3 kept)
  for (int i1 = 0; i1 < 4; ++i1) for (int j1 = 0; i1 + j1 < 4; ++j1) for (int k1 = 0; i1 + j1 + k1 < 4; ++k1) {
    double av = a.c[i1][j1][k1];
    if (av == 0.0) continue;
    for (int i2 = 0; i1 + i2 < 4; ++i2) for (int j2 = 0; i1 + i2 + j1 + j2 < 4; ++j2)
      for (int k2 = 0; i1 + i2 + j1 + j2 + k1 + k2 < 4; ++k2)
        o.c[i1 + i2][j1 + j2][k1 + k2] += s * av * b.c[i2][j2][k2];
  }
}
__device__ inline void p3_add(P3& o, const P3& a, double s) { for (int i = 0; i < 64; ++i) (&o.c[0][0][0])[i] += s * (&a.c[0][0][0])[i]; }

// univariate polynomials in z, coefficient index = power
__device__ inline void pz_mul(const double* a, int da, const double* b, int db, double* o) {
  for (int i = 0; i <= da + db; ++i) o[i] = 0.0;
  for (int i = 0; i <= da; ++i) for (int j = 0; j <= db; ++j) o[i + j] += a[i] * b[j];
}
__device__ inline double pz_eval(const double* p, int d, double z) { double v = p[d]; for (int i = d - 1; i >= 0; --i) v = v * z + p[i]; return v; }

// number of sign changes of the Sturm chain at z
__device__ inline int sturm_changes(const double (*ch)[12], const int* deg, int n, double z) {
  int cnt = 0; double prev = 0.0;
  for (int i = 0; i < n; ++i) {
    double v = pz_eval(ch[i], deg[i], z);
    if (v == 0.0) continue;
    if (prev != 0.0 && ((v < 0) != (prev < 0))) ++cnt;
    prev = v;
  }
  return cnt;
}

// real roots of a degree <= 10 polynomial; returns the count
__device__ int real_roots(const double* pin, int d, double* roots) {
  double p[12];
  while (d > 0 && fabs(pin[d]) < 1e-300) --d;
  if (d < 1) return 0;
  for (int i = 0; i <= d; ++i) p[i] = pin[i] / pin[d];
  double ch[12][12]; int deg[12]; int n = 0;
  for (int i = 0; i <= d; ++i) ch[0][i] = p[i];
  deg[0] = d;
  for (int i = 0; i < d; ++i) ch[1][i] = p[i + 1] * (i + 1);
  deg[1] = d - 1; n = 2;
  while (deg[n - 1] > 0 && n < 12) {                    // remainder sequence with negation
    const double* a = ch[n - 2]; const double* b = ch[n - 1];
    int da = deg[n - 2], db = deg[n - 1];
    double r[12];
    for (int i = 0; i <= da; ++i) r[i] = a[i];
    for (int k = da; k >= db; --k) {
      double q = r[k] / b[db];
      for (int j = 0; j <= db; ++j) r[k - db + j] -= q * b[j];
      r[k] = 0.0;
    }
    int dr = db - 1;
    double mx = 0.0;
    for (int i = 0; i <= dr; ++i) mx = fmax(mx, fabs(r[i]));
    while (dr > 0 && fabs(r[dr]) <= 1e-14 * mx) --dr;
    if (mx == 0.0) break;
    for (int i = 0; i <= dr; ++i) ch[n][i] = -r[i] / mx;
    deg[n] = dr; ++n;
  }
  double bound = 0.0;                                    // Cauchy bound
  for (int i = 0; i < d; ++i) bound = fmax(bound, fabs(p[i]));
  bound += 1.0;
  // iterative bisection over a small interval stack
  double lo[24], hi[24]; int clo[24], chi[24]; int sp = 0, nr = 0;
  lo[0] = -bound; hi[0] = bound; clo[0] = sturm_changes(ch, deg, n, -bound); chi[0] = sturm_changes(ch, deg, n, bound); sp = 1;
  while (sp > 0 && nr < 10) {
    --sp;
    double a = lo[sp], b = hi[sp]; int ca = clo[sp], cb = chi[sp];
    int k = ca - cb;
    if (k <= 0) continue;
    if (k == 1 || (b - a) < 1e-13 * (1.0 + fabs(a))) {
      // isolate by bisection on the sign of p, then polish with Newton
      double fa = pz_eval(p, d, a), fb = pz_eval(p, d, b);
      double x = 0.5 * (a + b);
      if ((fa < 0) != (fb < 0)) {
        for (int it = 0; it < 80 && (b - a) > 1e-15 * (1.0 + fabs(x)); ++it) {
          x = 0.5 * (a + b);
          double fx = pz_eval(p, d, x);
          if (fx == 0.0) break;
          if ((fx < 0) == (fa < 0)) { a = x; fa = fx; } else b = x;
        }
        x = 0.5 * (a + b);
      }
      for (int it = 0; it < 3; ++it) {
        double f = pz_eval(p, d, x), df = pz_eval(ch[1], d - 1, x);
        if (df == 0.0) break;
        double xn = x - f / df;
        if (!(xn >= lo[sp] - 1e-9 * (1 + fabs(x)) && xn <= hi[sp] + 1e-9 * (1 + fabs(x)))) break;
        x = xn;
      }
      roots[nr++] = x;
      continue;
    }
    double m = 0.5 * (a + b);
    int cm = sturm_changes(ch, deg, n, m);
    if (sp + 2 > 24) continue;
    lo[sp] = a; hi[sp] = m; clo[sp] = ca; chi[sp] = cm; ++sp;
    lo[sp] = m; hi[sp] = b; clo[sp] = cm; chi[sp] = cb; ++sp;
  }
  return nr;
}

// symmetric 3x3 Jacobi eigen-decomposition: A = V diag(w) V^T
__device__ inline void jacobi3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) {
      if (fabs(A[p][q]) < 1e-300) continue;
      double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
    }
  }
}

__device__ inline double det3(const double* R) {
  return R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
}

// fa/fb: bearing arrays (n x 3), idx: 8 sample indices (5 solver points + 3 disambiguation points).
// Writes the 3x4 model; returns false when no candidate exists.
__device__ bool solve(const double* fa, const double* fb, const int* idx, double* model) {
  // ---- null space of the 5x9 system  f1^T E f2 = 0  (row = outer(f1, f2) flattened row-major)
  double Q[5][9];
  for (int r = 0; r < 5; ++r) {
    const double* f1 = fa + 3 * idx[r]; const double* f2 = fb + 3 * idx[r];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Q[r][3 * a + b] = f1[a] * f2[b];
  }
  int pivcol[5]; bool ispiv[9];
  for (int c = 0; c < 9; ++c) ispiv[c] = false;
  for (int r = 0; r < 5; ++r) {                            // Gauss-Jordan with full pivoting
    int br = r, bc = -1; double best = 0.0;
    for (int i = r; i < 5; ++i) for (int c = 0; c < 9; ++c) if (!ispiv[c] && fabs(Q[i][c]) > best) { best = fabs(Q[i][c]); br = i; bc = c; }
    if (bc < 0 || best < 1e-14) return false;
    if (br != r) for (int c = 0; c < 9; ++c) { double t = Q[r][c]; Q[r][c] = Q[br][c]; Q[br][c] = t; }
    double inv = 1.0 / Q[r][bc];
    for (int c = 0; c < 9; ++c) Q[r][c] *= inv;
    for (int i = 0; i < 5; ++i) if (i != r) { double f = Q[i][bc]; if (f != 0.0) for (int c = 0; c < 9; ++c) Q[i][c] -= f * Q[r][c]; }
    pivcol[r] = bc; ispiv[bc] = true;
  }
  double N[4][9]; int nb = 0;
  for (int c = 0; c < 9 && nb < 4; ++c) {
    if (ispiv[c]) continue;
    for (int k = 0; k < 9; ++k) N[nb][k] = 0.0;
    N[nb][c] = 1.0;
    for (int r = 0; r < 5; ++r) N[nb][pivcol[r]] = -Q[r][c];
    double nn = 0.0; for (int k = 0; k < 9; ++k) nn += N[nb][k] * N[nb][k];
    nn = sqrt(nn); for (int k = 0; k < 9; ++k) N[nb][k] /= nn;
    ++nb;
  }
  // ---- the 10 cubic constraints
  P3 E[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) p3_lin(E[r][c], N[0][3 * r + c], N[1][3 * r + c], N[2][3 * r + c], N[3][3 * r + c]);
  P3 EEt[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { p3_zero(EEt[r][c]); for (int k = 0; k < 3; ++k) p3_mul_acc(EEt[r][c], E[r][k], E[c][k], 1.0); }
  P3 tr; p3_zero(tr); p3_add(tr, EEt[0][0], 1.0); p3_add(tr, EEt[1][1], 1.0); p3_add(tr, EEt[2][2], 1.0);
  // monomial order (Nister): x^3 y^3 x^2y xy^2 x^2z x^2 y^2z y^2 xyz xy | xz^2 xz x yz^2 yz y z^3 z^2 z 1
  const int mi[20] = {3, 0, 2, 1, 2, 2, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
  const int mj[20] = {0, 3, 1, 2, 0, 0, 2, 2, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0};
  const int mk[20] = {0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 2, 1, 0, 2, 1, 0, 3, 2, 1, 0};
  double A[10][20];
  {
    P3 d, t2; p3_zero(d);
    // det E
    for (int s = 0; s < 3; ++s) {
      int c1 = (s + 1) % 3, c2 = (s + 2) % 3;
      p3_zero(t2); p3_mul_acc(t2, E[1][c1], E[2][c2], 1.0); p3_mul_acc(t2, E[1][c2], E[2][c1], -1.0);
      p3_mul_acc(d, E[0][s], t2, 1.0);
    }
    for (int m = 0; m < 20; ++m) A[0][m] = d.c[mi[m]][mj[m]][mk[m]];
    int row = 1;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
      P3 acc; p3_zero(acc);
      for (int k = 0; k < 3; ++k) p3_mul_acc(acc, EEt[r][k], E[k][c], 2.0);
      p3_mul_acc(acc, tr, E[r][c], -1.0);
      for (int m = 0; m < 20; ++m) A[row][m] = acc.c[mi[m]][mj[m]][mk[m]];
      ++row;
    }
  }
  // ---- Gauss-Jordan on the first 10 columns (partial pivoting)
  for (int c = 0; c < 10; ++c) {
    int br = c; double best = fabs(A[c][c]);
    for (int r = c + 1; r < 10; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); br = r; }
    if (best < 1e-14) return false;
    if (br != c) for (int k = 0; k < 20; ++k) { double t = A[c][k]; A[c][k] = A[br][k]; A[br][k] = t; }
    double inv = 1.0 / A[c][c];
    for (int k = 0; k < 20; ++k) A[c][k] *= inv;
    for (int r = 0; r < 10; ++r) if (r != c) { double f = A[r][c]; if (f != 0.0) for (int k = 0; k < 20; ++k) A[r][k] -= f * A[c][k]; }
  }
  // rows: 4 = <x^2z>, 5 = <x^2>, 6 = <y^2z>, 7 = <y^2>, 8 = <xyz>, 9 = <xy>; right block columns
  // 10..19 = [xz^2 xz x yz^2 yz y z^3 z^2 z 1].   k = e - z f, l = g - z h, m = i - z j.
  double B[3][3][5];   // B[row][{x, y, 1}][power of z]
  for (int q = 0; q < 3; ++q) {
    const double* e = A[4 + 2 * q]; const double* f = A[5 + 2 * q];
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 5; ++k) B[q][a][k] = 0.0;
    // x-coefficient: e: xz^2 (10), xz (11), x (12);  -z f: xz^3, xz^2, xz
    B[q][0][0] = e[12]; B[q][0][1] = e[11] - f[12]; B[q][0][2] = e[10] - f[11]; B[q][0][3] = -f[10];
    B[q][1][0] = e[15]; B[q][1][1] = e[14] - f[15]; B[q][1][2] = e[13] - f[14]; B[q][1][3] = -f[13];
    B[q][2][0] = e[19]; B[q][2][1] = e[18] - f[19]; B[q][2][2] = e[17] - f[18]; B[q][2][3] = e[16] - f[17]; B[q][2][4] = -f[16];
  }
  // det B(z): degree 10
  double poly[12];
  for (int i = 0; i < 12; ++i) poly[i] = 0.0;
  {
    double t1[8], t2[8], t3[12];
    // cofactor expansion along the third column (degree-4 entries)
    for (int q = 0; q < 3; ++q) {
      int r1 = (q + 1) % 3, r2 = (q + 2) % 3;
      pz_mul(B[r1][0], 3, B[r2][1], 3, t1);
      pz_mul(B[r1][1], 3, B[r2][0], 3, t2);
      for (int i = 0; i <= 6; ++i) t1[i] -= t2[i];
      pz_mul(t1, 6, B[q][2], 4, t3);
      for (int i = 0; i <= 10; ++i) poly[i] += t3[i];
    }
  }
  double roots[10];
  int nr = real_roots(poly, 10, roots);
  // ---- candidates
  double bestq = 1000000.0; bool found = false;
  for (int ri = 0; ri < nr; ++ri) {
    const double z = roots[ri];
    double Bz[3][3];
    for (int q = 0; q < 3; ++q) { Bz[q][0] = pz_eval(B[q][0], 3, z); Bz[q][1] = pz_eval(B[q][1], 3, z); Bz[q][2] = pz_eval(B[q][2], 4, z); }
    // [x y 1] is orthogonal to every row: cross product of the best-conditioned row pair
    double bx = 0, by = 0, bw = 0, bn = -1.0;
    for (int a = 0; a < 3; ++a) {
      int b = (a + 1) % 3;
      double cx = Bz[a][1] * Bz[b][2] - Bz[a][2] * Bz[b][1];
      double cy = Bz[a][2] * Bz[b][0] - Bz[a][0] * Bz[b][2];
      double cw = Bz[a][0] * Bz[b][1] - Bz[a][1] * Bz[b][0];
      if (fabs(cw) > bn) { bn = fabs(cw); bx = cx; by = cy; bw = cw; }
    }
    if (bn < 1e-300) continue;
    const double x = bx / bw, y = by / bw;
    double Em[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
      Em[r][c] = x * N[0][3 * r + c] + y * N[1][3 * r + c] + z * N[2][3 * r + c] + N[3][3 * r + c];
    // SVD E = U S V^T via the eigen-decomposition of E^T E
    double M[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = Em[0][i] * Em[0][j] + Em[1][i] * Em[1][j] + Em[2][i] * Em[2][j];
    jacobi3(M, V);
    int o0 = 0, o1 = 1, o2 = 2;                      // sort eigenvalues descending
    if (M[o0][o0] < M[o1][o1]) { int t = o0; o0 = o1; o1 = t; }
    if (M[o1][o1] < M[o2][o2]) { int t = o1; o1 = o2; o2 = t; }
    if (M[o0][o0] < M[o1][o1]) { int t = o0; o0 = o1; o1 = t; }
    double s0 = sqrt(fmax(M[o0][o0], 0.0)), s1 = sqrt(fmax(M[o1][o1], 0.0));
    if (s0 < 1e-300 || s1 < 1e-300) continue;
    double Vs[3][3], U[3][3];
    for (int k = 0; k < 3; ++k) { Vs[k][0] = V[k][o0]; Vs[k][1] = V[k][o1]; }
    Vs[0][2] = Vs[1][0] * Vs[2][1] - Vs[2][0] * Vs[1][1];
    Vs[1][2] = Vs[2][0] * Vs[0][1] - Vs[0][0] * Vs[2][1];
    Vs[2][2] = Vs[0][0] * Vs[1][1] - Vs[1][0] * Vs[0][1];
    for (int k = 0; k < 3; ++k) {
      U[k][0] = (Em[k][0] * Vs[0][0] + Em[k][1] * Vs[1][0] + Em[k][2] * Vs[2][0]) / s0;
      U[k][1] = (Em[k][0] * Vs[0][1] + Em[k][1] * Vs[1][1] + Em[k][2] * Vs[2][1]) / s1;
    }
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    // Ra = U W V^T, Rb = U W^T V^T, W = [[0,-1,0],[1,0,0],[0,0,1]]; ta = s0 * U[:,2]
    double Ra[9], Rb[9], ta[3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
      // (U W)[r] = [U[r][1], -U[r][0], U[r][2]];  (U W^T)[r] = [-U[r][1], U[r][0], U[r][2]]
      Ra[3 * r + c] = U[r][1] * Vs[c][0] - U[r][0] * Vs[c][1] + U[r][2] * Vs[c][2];
      Rb[3 * r + c] = -U[r][1] * Vs[c][0] + U[r][0] * Vs[c][1] + U[r][2] * Vs[c][2];
    }
    if (det3(Ra) < 0) for (int i = 0; i < 9; ++i) Ra[i] = -Ra[i];
    if (det3(Rb) < 0) for (int i = 0; i < 9; ++i) Rb[i] = -Rb[i];
    for (int k = 0; k < 3; ++k) ta[k] = s0 * U[k][2];
    for (int cand = 0; cand < 4; ++cand) {
      const double* R = (cand < 2) ? Ra : Rb;
      double t[3] = {(cand & 1) ? -ta[0] : ta[0], (cand & 1) ? -ta[1] : ta[1], (cand & 1) ? -ta[2] : ta[2]};
      double q = 0.0;
      for (int k = 0; k < 8; ++k) q += relpose_score(R, t, fa + 3 * idx[k], fb + 3 * idx[k]);
      if (q < bestq) {
        bestq = q; found = true;
        for (int r = 0; r < 3; ++r) { model[4 * r] = R[3 * r]; model[4 * r + 1] = R[3 * r + 1]; model[4 * r + 2] = R[3 * r + 2]; model[4 * r + 3] = t[r]; }
      }
    }
  }
  return found;
}

}  // namespace fivept
