// rectify.cu -- row a2: UndistorterRectifier::undistortRectifyImage
// (reference src/frontend/UndistorterRectifier.cpp:115-128; cv::remap INTER_LINEAR,
// BORDER_REPLICATE, CV_32FC1 maps from cv::initUndistortRectifyMap, :248-258).
//
// The float maps are never stored: a one-time kernel recomputes map_x/map_y per pixel in f64
// (bit-identical to cv2's f32 maps) and converts them to cv::remap's fixed point (integer source
// pixel + 5 fractional bits per axis) into an 8 B/px table per camera; the per-frame kernel gathers
// the four taps and blends in exact integer arithmetic (15-bit weights).  HBM/L2 traffic per image:
// 8 B/px table read (coalesced) + 1 B/px gather (L2-friendly) + 1 B/px write.
#include "common.cuh"

// cv::remap: sx = cvRound(mx * INTER_TAB_SIZE), ix = sx >> 5, fx = sx & 31; saturate_cast<short>
// of the integer coordinates
__device__ __forceinline__ uint2 remap_entry(float mx, float my) {
  int sx = cv_round(mx * 32.0f), sy = cv_round(my * 32.0f);
  int ix = clampi(sx >> 5, -32768, 32767), iy = clampi(sy >> 5, -32768, 32767);
  return make_uint2((unsigned)(ix & 0xffff) | ((unsigned)(iy & 0xffff) << 16), (unsigned)((sx & 31) | ((sy & 31) << 5)));
}

__device__ __forceinline__ unsigned char remap_px(const unsigned char* __restrict__ src, int pitch, int W,
                                                  int H, uint2 e) {
  const int ix = (short)(e.x & 0xffff), iy = (short)(e.x >> 16), fx = e.y & 31, fy = (e.y >> 5) & 31;
  int x0 = clampi(ix, 0, W - 1), x1 = clampi(ix + 1, 0, W - 1);
  int y0 = clampi(iy, 0, H - 1), y1 = clampi(iy + 1, 0, H - 1);
  // 32x32 bilinear table: w = cvRound(((32-fy)/32 * (32-fx)/32) * 32768) -- exact integers
  int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32;
  int w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  const unsigned char* r0 = src + (size_t)y0 * pitch;
  const unsigned char* r1 = src + (size_t)y1 * pitch;
  int v = w00 * r0[x0] + w01 * r0[x1] + w10 * r1[x0] + w11 * r1[x1];
  return (unsigned char)((v + (1 << 14)) >> 15);
}

// grid: (ceil(W/4/64), ceil(H/RECT_ROWS), nimg); 64 threads x 4 px per row, RECT_ROWS rows per block
#define RECT_ROWS 8
__global__ void __launch_bounds__(64) rectify_kernel(DevCfg dc, const uint2* __restrict__ rmap,
                                                     const unsigned char* __restrict__ src, size_t src_stride,
                                                     unsigned char* __restrict__ dst, size_t dst_stride,
                                                     const StreamState* __restrict__ st, int mode_mask) {
  const int img = blockIdx.z;
  if (st && !mode_on(st[img].mode, mode_mask)) return;
  const int u0 = (blockIdx.x * 64 + threadIdx.x) * 4;
  if (u0 >= dc.W) return;
  const unsigned char* s = src + (size_t)img * src_stride;
  const bool full = u0 + 3 < dc.W && (dc.W & 3) == 0;
#pragma unroll 2
  for (int rr = 0; rr < RECT_ROWS; ++rr) {
    const int v = blockIdx.y * RECT_ROWS + rr;
    if (v >= dc.H) break;
    const uint2* m = rmap + (size_t)v * dc.W + u0;
    unsigned char* d = dst + (size_t)img * dst_stride + (size_t)v * dc.pitch + u0;
    if (full) {
      const uint4 m01 = __ldg(reinterpret_cast<const uint4*>(m));
      const uint4 m23 = __ldg(reinterpret_cast<const uint4*>(m) + 1);
      *reinterpret_cast<uchar4*>(d) = make_uchar4(remap_px(s, dc.pitch, dc.W, dc.H, make_uint2(m01.x, m01.y)),
                                                   remap_px(s, dc.pitch, dc.W, dc.H, make_uint2(m01.z, m01.w)),
                                                   remap_px(s, dc.pitch, dc.W, dc.H, make_uint2(m23.x, m23.y)),
                                                   remap_px(s, dc.pitch, dc.W, dc.H, make_uint2(m23.z, m23.w)));
    } else {
      for (int k = 0; k < 4 && u0 + k < dc.W; ++k) d[k] = remap_px(s, dc.pitch, dc.W, dc.H, m[k]);
    }
  }
}

__global__ void rmap_table_kernel(DevCfg dc, const CamModel* __restrict__ cams, int cam, uint2* __restrict__ rmap) {
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= dc.W) return;
  float a, b;
  rect_map_at(cams[cam], u, v, &a, &b);
  rmap[(size_t)v * dc.W + u] = remap_entry(a, b);
}

__global__ void maps_kernel(DevCfg dc, const CamModel* __restrict__ cams, int cam, float* __restrict__ mx,
                            float* __restrict__ my) {
  int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= dc.W) return;
  float a, b;
  rect_map_at(cams[cam], u, v, &a, &b);
  mx[(size_t)v * dc.W + u] = a;
  my[(size_t)v * dc.W + u] = b;
}

int launch_rectify(const DevCfg& dc, const uint2* rmap, const unsigned char* src, size_t src_stride,
                   unsigned char* dst, size_t dst_stride, int nimg, const StreamState* st, int mode_mask,
                   cudaStream_t s) {
  dim3 grid(((dc.W + 3) / 4 + 63) / 64, (dc.H + RECT_ROWS - 1) / RECT_ROWS, nimg);
  rectify_kernel<<<grid, 64, 0, s>>>(dc, rmap, src, src_stride, dst, dst_stride, st, mode_mask);
  return 1;
}

int launch_rmap_table(const DevCfg& dc, const CamModel* d_cam, int cam, uint2* rmap, cudaStream_t s) {
  dim3 grid((dc.W + 127) / 128, dc.H);
  rmap_table_kernel<<<grid, 128, 0, s>>>(dc, d_cam, cam, rmap);
  return 1;
}

int launch_maps(const DevCfg& dc, const CamModel* d_cam, int cam, float* mx, float* my, cudaStream_t s) {
  dim3 grid((dc.W + 127) / 128, dc.H);
  maps_kernel<<<grid, 128, 0, s>>>(dc, d_cam, cam, mx, my);
  return 1;
}
