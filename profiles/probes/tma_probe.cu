// Standalone probe: one cp.async.bulk.tensor.3d box load of a u8 (x, y, image) tensor, map passed as
// __grid_constant__ or from global memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../kimera_vio_b200/csrc/tma.cuh"

struct __align__(64) Maps { unsigned char m[2][128]; };

namespace tma {
__device__ __forceinline__ void tensor_g2s_2d(void* dst_smem, const void* tmap, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
}
__global__ void probe(const __grid_constant__ Maps maps, const void* gmaps, int which, int x, int y, int z, unsigned char* out, int rank, int bytes) {
  __shared__ __align__(128) unsigned char box[1024];
  __shared__ __align__(8) unsigned long long bar;
  uint64_t* b = reinterpret_cast<uint64_t*>(&bar);
  if (threadIdx.x == 0) { tma::mbar_init(b, 1); tma::fence_barrier_init(); }
  __syncwarp();
  const void* mp = which == 0 ? (const void*)&maps.m[1][0] : (const void*)((const char*)gmaps + 128);
  if (threadIdx.x == 0) {
    tma::mbar_expect_tx(b, bytes);
    if (rank == 3) tma::tensor_g2s_3d(box, mp, x, y, z, b);
    else tma::tensor_g2s_2d(box, mp, x, y, b);
  }
  tma::mbar_wait(b, 0);
  for (int i = threadIdx.x; i < 1024; i += 32) out[i] = box[i];
}

int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 3, bw = argc > 2 ? atoi(argv[2]) : 32, bh = argc > 3 ? atoi(argv[3]) : 32;
  const int only = argc > 4 ? atoi(argv[4]) : -1;
  const int W = 94, H = 60, P = 96, B = 3;
  const size_t stride = 8192;
  std::vector<unsigned char> h(stride * B);
  for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < P; ++x) h[b * stride + y * P + x] = (unsigned char)(b * 50 + y * 3 + x);
  unsigned char *d, *dout; cudaMalloc(&d, h.size()); cudaMalloc(&dout, 1024);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry %p q=%d\n", fn, (int)q);
  Maps maps; memset(&maps, 0, sizeof(maps));
  for (int k = 0; k < 2; ++k) {
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B}, strides[2] = {(cuuint64_t)P, (cuuint64_t)stride};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(reinterpret_cast<CUtensorMap*>(maps.m[k]), CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d -> %d\n", k, (int)r);
  }
  void* gm; cudaMalloc(&gm, sizeof(maps)); cudaMemcpy(gm, &maps, sizeof(maps), cudaMemcpyHostToDevice);
  for (int which = 0; which < 2; ++which) {
    if (only >= 0 && which != only) continue;
    const int x = -3, y = 40, z = rank == 3 ? 2 : 0;
    probe<<<1, 32>>>(maps, gm, which, x, y, z, dout, rank, bw * bh);
    cudaError_t e = cudaDeviceSynchronize();
    printf("which=%d: %s\n", which, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<unsigned char> o(1024); cudaMemcpy(o.data(), dout, 1024, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) {
      const int xx = x + c, yy = y + r;
      const unsigned char want = (xx < 0 || xx >= W || yy < 0 || yy >= H) ? 0 : h[z * stride + yy * P + xx];
      bad += o[r * bw + c] != want;
    }
    printf("which=%d mismatches=%d\n", which, bad);
  }
  return 0;
}
