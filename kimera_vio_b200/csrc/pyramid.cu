// pyramid.cu -- image pyramid of cv::calcOpticalFlowPyrLK (reference src/frontend/Tracker.cpp:137-146
// -> cv::buildOpticalFlowPyramid): level l+1 = cv::pyrDown(level l): separable [1 4 6 4 1] at even
// pixels, BORDER_REFLECT_101, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2).  Exact integer arithmetic.
//
// Two launches per batch: (1) level 0 -> 1, tiled over the whole batch (the only level with real
// HBM traffic: reads W*H, writes W*H/4); (2) levels 1 -> 2 -> ... -> L inside ONE CTA per image
// (<= 90 KB of pixels, L2/L1 resident, __syncthreads between levels).
#include "common.cuh"

__device__ __forceinline__ int pyr_px(const unsigned char* __restrict__ src, int sp, int sw, int sh, int x,
                                      int y) {
  // vertical then horizontal is arithmetically identical to OpenCV's horizontal-then-vertical
  // (all integer, no intermediate rounding).
  int acc = 0;
  const int wgt[5] = {1, 4, 6, 4, 1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    int yy = reflect101(2 * y + j - 2, sh);
    const unsigned char* r = src + (size_t)yy * sp;
    int rowacc = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) rowacc += wgt[i] * r[reflect101(2 * x + i - 2, sw)];
    acc += wgt[j] * rowacc;
  }
  return (acc + 128) >> 8;
}

__global__ void __launch_bounds__(256) pyr_level1_kernel(DevCfg dc, unsigned char* __restrict__ pyr) {
  unsigned char* base = pyr + (size_t)blockIdx.z * dc.pyr_stride;
  const unsigned char* src = base + dc.lvl_off[0];
  unsigned char* dst = base + dc.lvl_off[1];
  int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dc.lvl_w[1] || y >= dc.lvl_h[1]) return;
  dst[(size_t)y * dc.lvl_pitch[1] + x] =
      (unsigned char)pyr_px(src, dc.lvl_pitch[0], dc.lvl_w[0], dc.lvl_h[0], x, y);
}

__global__ void __launch_bounds__(1024) pyr_upper_kernel(DevCfg dc, unsigned char* __restrict__ pyr) {
  unsigned char* base = pyr + (size_t)blockIdx.x * dc.pyr_stride;
  for (int l = 2; l < dc.n_levels; ++l) {
    const unsigned char* src = base + dc.lvl_off[l - 1];
    unsigned char* dst = base + dc.lvl_off[l];
    int w = dc.lvl_w[l], h = dc.lvl_h[l];
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
      int y = i / w, x = i - y * w;
      dst[(size_t)y * dc.lvl_pitch[l] + x] =
          (unsigned char)pyr_px(src, dc.lvl_pitch[l - 1], dc.lvl_w[l - 1], dc.lvl_h[l - 1], x, y);
    }
    __threadfence_block();
    __syncthreads();
  }
}

int launch_pyramid(const DevCfg& dc, unsigned char* pyr, int nimg, cudaStream_t s) {
  int n = 0;
  if (dc.n_levels > 1) {
    dim3 grid((dc.lvl_w[1] + 31) / 32, (dc.lvl_h[1] + 7) / 8, nimg);
    pyr_level1_kernel<<<grid, 256, 0, s>>>(dc, pyr);
    ++n;
  }
  if (dc.n_levels > 2) {
    pyr_upper_kernel<<<nimg, 1024, 0, s>>>(dc, pyr);
    ++n;
  }
  return n;
}
