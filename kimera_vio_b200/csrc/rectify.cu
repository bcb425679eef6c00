// rectify.cu -- row a2: UndistorterRectifier::undistortRectifyImage
// (reference src/frontend/UndistorterRectifier.cpp:115-128; cv::remap INTER_LINEAR,
// BORDER_REPLICATE, CV_32FC1 maps from cv::initUndistortRectifyMap, :248-258).
//
// The float maps are never stored: each thread recomputes map_x/map_y for its pixel in f64
// (bit-identical to cv2's f32 maps), converts to cv::remap's fixed point (5 fractional bits,
// 15-bit weights) and blends in exact integer arithmetic.  HBM traffic: 1 B/px read (gather,
// L2-friendly: neighbouring outputs map to neighbouring sources) + 1 B/px write.
#include "common.cuh"

__device__ __forceinline__ unsigned char remap_px(const unsigned char* __restrict__ src, int pitch, int W,
                                                  int H, float mx, float my) {
  // cv::remap: sx = cvRound(mx * INTER_TAB_SIZE), ix = sx >> 5, fx = sx & 31
  int sx = cv_round(mx * 32.0f), sy = cv_round(my * 32.0f);
  int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
  // saturate_cast<short> of the integer coordinates
  ix = clampi(ix, -32768, 32767);
  iy = clampi(iy, -32768, 32767);
  int x0 = clampi(ix, 0, W - 1), x1 = clampi(ix + 1, 0, W - 1);
  int y0 = clampi(iy, 0, H - 1), y1 = clampi(iy + 1, 0, H - 1);
  // 32x32 bilinear table: w = cvRound(((32-fy)/32 * (32-fx)/32) * 32768) -- exact integers
  int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32;
  int w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  int v = w00 * src[(size_t)y0 * pitch + x0] + w01 * src[(size_t)y0 * pitch + x1] +
          w10 * src[(size_t)y1 * pitch + x0] + w11 * src[(size_t)y1 * pitch + x1];
  return (unsigned char)((v + (1 << 14)) >> 15);
}

// grid: (ceil(W/4/64), ceil(H/RECT_ROWS), nimg); 64 threads x 4 px per row, RECT_ROWS rows per block
#define RECT_ROWS 8
__global__ void __launch_bounds__(64) rectify_kernel(DevCfg dc, const CamModel* __restrict__ cams, int cam,
                                                     const unsigned char* __restrict__ src, size_t src_stride,
                                                     unsigned char* __restrict__ dst, size_t dst_stride,
                                                     const StreamState* __restrict__ st, int mode_mask) {
  const int img = blockIdx.z;
  if (st && !mode_on(st[img].mode, mode_mask)) return;
  const int u0 = (blockIdx.x * 64 + threadIdx.x) * 4;
  if (u0 >= dc.W) return;
  // every thread reads the model from global (L1-broadcast); cheap and avoids a barrier
  const CamModel& cm = cams[cam];
  const unsigned char* s = src + (size_t)img * src_stride;
  for (int rr = 0; rr < RECT_ROWS; ++rr) {
    const int v = blockIdx.y * RECT_ROWS + rr;
    if (v >= dc.H) break;
    unsigned char out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int u = u0 + k;
      float mx, my;
      rect_map_at(cm, u < dc.W ? u : dc.W - 1, v, &mx, &my);
      out[k] = remap_px(s, dc.pitch, dc.W, dc.H, mx, my);
    }
    unsigned char* d = dst + (size_t)img * dst_stride + (size_t)v * dc.pitch + u0;
    if (u0 + 3 < dc.W) {
      *reinterpret_cast<uchar4*>(d) = make_uchar4(out[0], out[1], out[2], out[3]);
    } else {
      for (int k = 0; k < 4 && u0 + k < dc.W; ++k) d[k] = out[k];
    }
  }
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

int launch_rectify(const DevCfg& dc, const CamModel* d_cam, int cam, const unsigned char* src,
                   size_t src_stride, unsigned char* dst, size_t dst_stride, int nimg,
                   const StreamState* st, int mode_mask, cudaStream_t s) {
  dim3 grid((dc.W / 4 + 63) / 64, (dc.H + RECT_ROWS - 1) / RECT_ROWS, nimg);
  rectify_kernel<<<grid, 64, 0, s>>>(dc, d_cam, cam, src, src_stride, dst, dst_stride, st, mode_mask);
  return 1;
}

int launch_maps(const DevCfg& dc, const CamModel* d_cam, int cam, float* mx, float* my, cudaStream_t s) {
  dim3 grid((dc.W + 127) / 128, dc.H);
  maps_kernel<<<grid, 128, 0, s>>>(dc, d_cam, cam, mx, my);
  return 1;
}
