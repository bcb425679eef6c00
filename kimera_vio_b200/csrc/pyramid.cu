// pyramid.cu -- image pyramid of cv::calcOpticalFlowPyrLK (reference src/frontend/Tracker.cpp:137-146
// -> cv::buildOpticalFlowPyramid): level l+1 = cv::pyrDown(level l): separable [1 4 6 4 1] at even
// pixels, BORDER_REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2).  Exact integer arithmetic.
//
// One launch per level, tiled over the whole batch: thread (x, 8 output rows) marches down the
// source rows keeping the last five horizontal [1 4 6 4 1] row sums in registers (each source row is
// filtered once per output column: 2 x 5 loads + 5 multiply-adds per output instead of 25 + 25).
// Level 0 -> 1 is the only level with real HBM traffic (reads W*H, writes W*H/4).
#include "common.cuh"

#define PYR_ROWS 8
// grid (ceil(w/128), ceil(h/PYR_ROWS), nimg), block 128
__global__ void __launch_bounds__(128) pyr_level_kernel(DevCfg dc, unsigned char* __restrict__ pyr, int level) {
  unsigned char* base = pyr + (size_t)blockIdx.z * dc.pyr_stride;
  const unsigned char* __restrict__ src = base + dc.lvl_off[level - 1];
  unsigned char* __restrict__ dst = base + dc.lvl_off[level];
  const int sw = dc.lvl_w[level - 1], sh = dc.lvl_h[level - 1], sp = dc.lvl_pitch[level - 1];
  const int w = dc.lvl_w[level], h = dc.lvl_h[level], dp = dc.lvl_pitch[level];
  const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * PYR_ROWS;
  if (x >= w) return;
  const int c0 = reflect101(2 * x - 2, sw), c1 = reflect101(2 * x - 1, sw), c2 = 2 * x,
            c3 = reflect101(2 * x + 1, sw), c4 = reflect101(2 * x + 2, sw);
  auto hsum = [&](int r) {
    const unsigned char* p = src + (size_t)reflect101(r, sh) * sp;
    return (int)p[c0] + 4 * (int)p[c1] + 6 * (int)p[c2] + 4 * (int)p[c3] + (int)p[c4];
  };
  // vertical then horizontal order is irrelevant: all integer, no intermediate rounding
  int a = hsum(2 * y0 - 2), b = hsum(2 * y0 - 1), c = hsum(2 * y0);
#pragma unroll
  for (int u = 0; u < PYR_ROWS; ++u) {
    const int y = y0 + u;
    if (y >= h) break;
    const int d = hsum(2 * y + 1), e = hsum(2 * y + 2);
    dst[(size_t)y * dp + x] = (unsigned char)((a + 4 * b + 6 * c + 4 * d + e + 128) >> 8);
    a = c; b = d; c = e;
  }
}

int launch_pyramid(const DevCfg& dc, unsigned char* pyr, int nimg, cudaStream_t s) {
  int n = 0;
  for (int l = 1; l < dc.n_levels; ++l) {
    dim3 grid((dc.lvl_w[l] + 127) / 128, (dc.lvl_h[l] + PYR_ROWS - 1) / PYR_ROWS, nimg);
    pyr_level_kernel<<<grid, 128, 0, s>>>(dc, pyr, l);
    ++n;
  }
  return n;
}
