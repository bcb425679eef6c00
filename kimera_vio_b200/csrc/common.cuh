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

// cv::undistortPoints (cvUndistortPointsInternal), radial-tangential 4-coefficient model, default
// criteria = exactly 5 fixed-point iterations, f64.  mode 0: no R, no P; 1: R only; 2: R and P
// (OpenCV folds P into the rotation first: RR = P[:, :3] * R, CamModel::RP); 3: P only.
__device__ __forceinline__ void undistort_point(const CamModel& c, float u, float v, int mode, float* ox,
                                                float* oy) {
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

// f32 map value of cv::initUndistortRectifyMap (CV_32FC1) at integer pixel (u, v), recomputed in
// f64 exactly as validated against cv2 (scratch prototype: 0 mismatches in 4 x 360 960 values).
__device__ __forceinline__ void rect_map_at(const CamModel& c, int u, int v, float* mx, float* my) {
  double ud = (double)u, vd = (double)v;
  double X = (c.iR[0] * ud + c.iR[1] * vd) + c.iR[2];
  double Y = (c.iR[3] * ud + c.iR[4] * vd) + c.iR[5];
  double Wd = (c.iR[6] * ud + c.iR[7] * vd) + c.iR[8];
  double w = 1.0 / Wd, x = X * w, y = Y * w;
  double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
  double kr = 1 + ((0.0 * r2 + c.k2) * r2 + c.k1) * r2;
  double xd = (x * kr + c.p1 * _2xy) + c.p2 * (r2 + 2 * x2);
  double yd = (y * kr + c.p1 * (r2 + 2 * y2)) + c.p2 * _2xy;
  *mx = (float)(c.fx * xd + c.cx);
  *my = (float)(c.fy * yd + c.cy);
}

__device__ __forceinline__ bool mode_on(int mode, int mask) { return (mask >> mode) & 1; }
