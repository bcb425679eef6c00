// Probe 2: the CUDA programming guide's TMA example (libcu++ API), u8 / i32, to compare against tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <typename T, int BW, int BH>
__global__ void kern(const __grid_constant__ CUtensorMap tm, int x, int y, T* out) {
  __shared__ alignas(128) T buf[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &tm, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&buf[0][0])[i];
}

template <typename T, int BW, int BH>
int run(CUtensorMapDataType dt, int W, int H, int P, int x, int y) {
  std::vector<T> h((size_t)P * H);
  for (int yy = 0; yy < H; ++yy) for (int xx = 0; xx < P; ++xx) h[(size_t)yy * P + xx] = (T)(yy * 3 + xx);
  T *d, *dout; cudaMalloc(&d, h.size() * sizeof(T)); cudaMalloc(&dout, BW * BH * sizeof(T));
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)P * sizeof(T)};
  cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d; ", (int)r);
  kern<T, BW, BH><<<1, 128>>>(tm, x, y, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s; ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<T> o(BW * BH); cudaMemcpy(o.data(), dout, o.size() * sizeof(T), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
    const int xx = x + c, yy = y + r2;
    const T want = (xx < 0 || xx >= W || yy < 0 || yy >= H) ? (T)0 : h[(size_t)yy * P + xx];
    bad += o[r2 * BW + c] != want;
  }
  printf("mismatches=%d\n", bad);
  return 0;
}

int main(int argc, char** argv) {
  const int t = argc > 1 ? atoi(argv[1]) : 0;
  if (t == 0) return run<int, 32, 8>(CU_TENSOR_MAP_DATA_TYPE_INT32, 256, 64, 256, 32, 16);
  if (t == 1) return run<unsigned char, 32, 32>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 256, 64, 256, 32, 16);
  if (t == 2) return run<unsigned char, 32, 32>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 94, 60, 96, -3, 40);
  if (t == 3) return run<unsigned char, 64, 16>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 94, 60, 96, 5, 7);
  if (t == 9) return run<unsigned char, 32, 32>(CU_TENSOR_MAP_DATA_TYPE_UINT8, atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
  return 0;
}
