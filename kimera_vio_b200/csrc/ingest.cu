// ingest.cu -- row f3 (the step before the path): the grayscale conditioning of
// UtilsOpenCV::ReadAndConvertToGrayScale (reference src/utils/UtilsOpenCV.cpp:390-403) that the data providers
// apply to every image when stereo_matching_params.equalize_image is set (EurocDataProvider.cpp:146-200,
// StereoMatchingParams.cpp:80-90: YAML key equalizeImage): cv::equalizeHist, in place, on the device.
//   hist[256] -> first non-empty bin i0 -> scale = 255.f / (total - hist[i0]) -> lut[i] = saturate_cast<uchar>(
//   cvRound(float(sum_{i0 < j <= i} hist[j]) * scale)), lut[i0] = 0 -> dst = lut[src]; a constant image stays as it is.
// Integer histogram + one float multiply per level: bit-exact with cv2.equalizeHist (tests/test_gpu_ingest.py).
// (PNG decoding stays on the host: the pipeline takes decoded 8-bit images from pageable, pinned or device memory.)
#include "common.cuh"

// grid (nimg), block 1024: one CTA per image
__global__ void __launch_bounds__(1024) equalize_hist_kernel(DevCfg dc, unsigned char* __restrict__ imgs, size_t img_stride,
                                                             const StreamState* __restrict__ st, int mode_mask) {
  __shared__ int hist[256];
  __shared__ unsigned char lut[256];
  __shared__ int s_const;
  const int b = blockIdx.x;
  if (st && !mode_on(st[b].mode, mode_mask)) return;
  unsigned char* img = imgs + (size_t)b * img_stride;
  const int W = dc.W, H = dc.H, pitch = dc.pitch;
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  __syncthreads();
  const int wq = W >> 2;
  for (int y = threadIdx.x / 32; y < H; y += blockDim.x / 32) {          // one warp per row: 32-bit loads
    const unsigned char* row = img + (size_t)y * pitch;
    for (int q = threadIdx.x & 31; q < wq; q += 32) {
      const unsigned int v = reinterpret_cast<const unsigned int*>(row)[q];
      atomicAdd(&hist[v & 255u], 1); atomicAdd(&hist[(v >> 8) & 255u], 1);
      atomicAdd(&hist[(v >> 16) & 255u], 1); atomicAdd(&hist[v >> 24], 1);
    }
    for (int x = (wq << 2) + (threadIdx.x & 31); x < W; x += 32) atomicAdd(&hist[row[x]], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = W * H;
    int i = 0;
    while (!hist[i]) ++i;
    s_const = hist[i] == total;
    if (!s_const) {
      const float scale = 255.f / (float)(total - hist[i]);
      int sum = 0;
      for (int k = 0; k <= i; ++k) lut[k] = 0;
      for (++i; i < 256; ++i) {
        sum += hist[i];
        const int v = cv_round((float)sum * scale);
        lut[i] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    }
  }
  __syncthreads();
  if (s_const) return;
  for (int y = threadIdx.x / 32; y < H; y += blockDim.x / 32) {
    unsigned char* row = img + (size_t)y * pitch;
    for (int q = threadIdx.x & 31; q < wq; q += 32) {
      const unsigned int v = reinterpret_cast<const unsigned int*>(row)[q];
      reinterpret_cast<unsigned int*>(row)[q] = (unsigned int)lut[v & 255u] | ((unsigned int)lut[(v >> 8) & 255u] << 8) |
                                                ((unsigned int)lut[(v >> 16) & 255u] << 16) | ((unsigned int)lut[v >> 24] << 24);
    }
    for (int x = (wq << 2) + (threadIdx.x & 31); x < W; x += 32) row[x] = lut[row[x]];
  }
}

int launch_equalize(const DevCfg& dc, unsigned char* imgs, size_t img_stride, int nimg, const StreamState* st, int mode_mask,
                    cudaStream_t s) {
  equalize_hist_kernel<<<nimg, 1024, 0, s>>>(dc, imgs, img_stride, st, mode_mask);
  return 1;
}
