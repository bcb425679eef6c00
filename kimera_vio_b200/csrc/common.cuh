// common.cuh -- device helpers shared by all kernels.  Compiled with -fmad=false: every float /
// double operation below rounds exactly like the scalar x86 code it mirrors; fused multiply-adds
// appear only where the reference arithmetic itself is fused (explicit fmaf / fma calls).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kvfe_internal.h"

#define KVFE_FULL_MASK 0xffffffffu

__device__ __forceinline__ int reflect101(int i, int n) {
  // cv::BORDER_REFLECT_101 for |overshoot| < n (single bounce)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cvRound(float): round half to even (SSE cvtss2si)
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }
// C round(): half away from zero, then (int)
__device__ __forceinline__ int c_round(float v) { return (int)roundf(v); }
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KVFE_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KVFE_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KVFE_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(KVFE_FULL_MASK, v, o);
  return v;
}

// order-preserving float <-> uint encoding for atomicMax on floats of any sign
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Eigen fixed-size-3 reductions: a0 + (a1 + a2)
__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
}
__device__ __forceinline__ double dot3s(const double* a, int sa, const double* b, int sb) {
  return a[0] * b[0] + (a[sa] * b[sb] + a[2 * sa] * b[2 * sb]);
}
__device__ __forceinline__ void matvec3(const double* M, const double* v, double* o) {
  double r0 = dot3(M, v), r1 = dot3(M + 3, v), r2 = dot3(M + 6, v);
  o[0] = r0; o[1] = r1; o[2] = r2;
}
__device__ __forceinline__ void mattvec3(const double* M, const double* v, double* o) {
  double r0 = dot3s(M, 3, v, 1), r1 = dot3s(M + 1, 3, v, 1), r2 = dot3s(M + 2, 3, v, 1);
  o[0] = r0; o[1] = r1; o[2] = r2;
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void matmul3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = dot3s(A + 3 * i, 1, B + j, 3);
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}

// cv::fisheye::undistortPoints (OpenCV 4.13, default criteria COUNT + EPS, 10 iterations, 1e-8) -- the equidistant
// model of UndistorterRectifier::UndistortRectifyKeypoints (UndistorterRectifier.cpp:49-56): Newton iterations on
// theta_d = theta (1 + k1 theta^2 + k2 theta^4 + k3 theta^6 + k4 theta^8), f64; a point that does not converge or
// whose theta changes sign comes back as (-1e6, -1e6).  Restated against cv2 in scratch/fisheye_probe.py (bit-exact
// on 16 000 points incl. non-convergent ones).  Modes as below.
__device__ __forceinline__ void undistort_point_fisheye(const CamModel& c, float u, float v, int mode, float* ox, float* oy) {
  const double pw0 = ((double)u - c.cx) / c.fx, pw1 = ((double)v - c.cy) / c.fy;
  double theta_d = sqrt(pw0 * pw0 + pw1 * pw1);
  const double half_pi = 1.5707963267948966;
  theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
  bool converged = false;
  double theta = theta_d, scale = 0.0;
  if (fabs(theta_d) > 1e-8) {
#pragma unroll 1
    for (int j = 0; j < 10; ++j) {
      const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
      const double a = c.k1 * t2, b = c.k2 * t4, cc = c.p1 * t6, d = c.p2 * t8;
      const double fix = (theta * (1 + a + b + cc + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * cc + 9 * d);
      theta = theta - fix;
      if (fabs(fix) < 1e-8) { converged = true; break; }
    }
    scale = tan(theta) / theta_d;
  } else {
    converged = true;
  }
  const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
  if (!converged || flipped) { *ox = -1000000.0f; *oy = -1000000.0f; return; }
  double x = pw0 * scale, y = pw1 * scale;
  if (mode != 0) {
    const double* M = (mode == 1) ? c.R : (mode == 2 ? c.RP : c.PP);
    const double xx = (M[0] * x + M[1] * y) + M[2];
    const double yy = (M[3] * x + M[4] * y) + M[5];
    const double ww = (M[6] * x + M[7] * y) + M[8];
    x = xx / ww;
    y = yy / ww;
  }
  *ox = (float)x;
  *oy = (float)y;
}

// cv::undistortPoints (cvUndistortPointsInternal), radial-tangential 4-coefficient model, default
// criteria = exactly 5 fixed-point iterations, f64.  mode 0: no R, no P; 1: R only; 2: R and P
// (OpenCV folds P into the rotation first: RR = P[:, :3] * R, CamModel::RP); 3: P only.
__device__ __forceinline__ void undistort_point(const CamModel& c, float u, float v, int mode, float* ox,
                                                float* oy) {
  if (c.model == 1) { undistort_point_fisheye(c, u, v, mode, ox, oy); return; }
  const double ifx = 1.0 / c.fx, ify = 1.0 / c.fy;
  double x0 = ((double)u - c.cx) * ifx;
  double y0 = ((double)v - c.cy) * ify;
  double x = x0, y = y0;
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    double r2 = x * x + y * y;
    double icdist = 1.0 / (1 + ((0.0 * r2 + c.k2) * r2 + c.k1) * r2);
    if (icdist < 0) { x = x0; y = y0; break; }
    double dX = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x);
    double dY = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
    x = (x0 - dX) * icdist;
    y = (y0 - dY) * icdist;
  }
  if (mode != 0) {
    const double* M = (mode == 1) ? c.R : (mode == 2 ? c.RP : c.PP);
    double xx = M[0] * x + M[1] * y + M[2];
    double yy = M[3] * x + M[4] * y + M[5];
    double ww = 1. / (M[6] * x + M[7] * y + M[8]);
    x = xx * ww;
    y = yy * ww;
  }
  *ox = (float)x;
  *oy = (float)y;
}

// f32 map value of cv::fisheye::initUndistortRectifyMap (CV_32FC1; UndistorterRectifier.cpp:260-268) at integer pixel
// (u, v).  OpenCV walks a row with three running f64 sums (_x += iR(0,0) per column), so column u is reached by u
// sequential additions from the row start -- reproduced as such (one-off table build and a few hundred sparse
// look-ups per keyframe).  0 mismatches against cv2 in 2 x 307 200 values (scratch/fisheye_probe.py).
__device__ __forceinline__ void rect_map_at_fisheye(const CamModel& c, int u, int v, float* mx, float* my) {
  const double vd = (double)v;
  double X = vd * c.iR[1] + c.iR[2], Y = vd * c.iR[4] + c.iR[5], Wd = vd * c.iR[7] + c.iR[8];
#pragma unroll 1
  for (int j = 0; j < u; ++j) { X += c.iR[0]; Y += c.iR[3]; Wd += c.iR[6]; }
  double uu, vv;
  if (Wd <= 0) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    uu = (X > 0) ? -inf : inf;
    vv = (Y > 0) ? -inf : inf;
  } else {
    const double x = X / Wd, y = Y / Wd;
    const double r = sqrt(x * x + y * y);
    const double theta = atan(r);
    const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    const double theta_d = theta * (1 + c.k1 * t2 + c.k2 * t4 + c.p1 * t6 + c.p2 * t8);
    const double scale = (r == 0) ? 1.0 : theta_d / r;
    uu = c.fx * x * scale + c.cx;
    vv = c.fy * y * scale + c.cy;
  }
  *mx = (float)uu;
  *my = (float)vv;
}

// f32 map value of cv::initUndistortRectifyMap (CV_32FC1) at integer pixel (u, v), recomputed in
// f64 exactly as validated against cv2 (scratch prototype: 0 mismatches in 4 x 360 960 values).
__device__ __forceinline__ void rect_map_at(const CamModel& c, int u, int v, float* mx, float* my) {
  if (c.model == 1) { rect_map_at_fisheye(c, u, v, mx, my); return; }
  double ud = (double)u, vd = (double)v;
  // OpenCV's row loop runs in an FMA-enabled translation unit (the AVX2 dispatch of initUndistortRectifyMap): the row
  // start i * ir[1] + ir[2], the column term and the final fx * xd + u0 are fused.  Where the result is far from zero the
  // fusions are invisible after the rounding to f32; they decide the last bit where it crosses zero or sits on a f32 tie
  // (zero-distortion rigs: column 0 / row 0 of the map, row 15 of params/uHumans2).  Pinned on the CPU against cv2 over the
  // full maps of Euroc L/R, uHumans1/2, D455 L and the RGB-D test camera: 0 mismatches (one 1e-13 px residue on D455 R).
  double X = fma(ud, c.iR[0], fma(vd, c.iR[1], c.iR[2]));
  double Y = fma(ud, c.iR[3], fma(vd, c.iR[4], c.iR[5]));
  double Wd = fma(ud, c.iR[6], fma(vd, c.iR[7], c.iR[8]));
  double w = 1.0 / Wd, x = X * w, y = Y * w;
  double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
  double kr = 1 + ((0.0 * r2 + c.k2) * r2 + c.k1) * r2;
  double xd = (x * kr + c.p1 * _2xy) + c.p2 * (r2 + 2 * x2);
  double yd = (y * kr + c.p1 * (r2 + 2 * y2)) + c.p2 * _2xy;
  *mx = (float)fma(c.fx, xd, c.cx);
  *my = (float)fma(c.fy, yd, c.cy);
}

__device__ __forceinline__ bool mode_on(int mode, int mask) { return (mask >> mode) & 1; }
