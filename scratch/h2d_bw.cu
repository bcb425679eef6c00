#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void spin(long long cycles) { long long t = clock64(); while (clock64() - t < cycles) {} }
int main() {
  const size_t sz = 752 * 480; const int N = 64, NS = 32;
  unsigned char *h, *d, *hp, *dp;
  CK(cudaMallocHost(&h, N * sz)); CK(cudaMalloc(&d, N * sz));
  CK(cudaMallocHost(&hp, 32 * 104544)); CK(cudaMalloc(&dp, 32 * 104544));
  std::vector<cudaStream_t> ss(NS);
  for (auto& s : ss) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  auto run = [&](const char* name, auto fn) {
    fn(); cudaDeviceSynchronize();
    const int reps = 50;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) fn();
    auto t1 = std::chrono::steady_clock::now();
    cudaDeviceSynchronize();
    auto t2 = std::chrono::steady_clock::now();
    double enq = std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
    double tot = std::chrono::duration<double, std::milli>(t2 - t0).count() / reps;
    printf("%-44s enqueue %.3f ms  total %.3f ms per batch\n", name, enq, tot);
  };
  run("H2D 64 x 360KB, 32 streams", [&] { for (int i = 0; i < N; ++i) cudaMemcpyAsync(d + i * sz, h + i * sz, sz, cudaMemcpyHostToDevice, ss[i % NS]); });
  run("H2D 64 x 360KB, 1 stream", [&] { for (int i = 0; i < N; ++i) cudaMemcpyAsync(d + i * sz, h + i * sz, sz, cudaMemcpyHostToDevice, ss[0]); });
  run("H2D 8 x 2.9MB, 8 streams", [&] { for (int i = 0; i < 8; ++i) cudaMemcpyAsync(d + i * 8 * sz, h + i * 8 * sz, 8 * sz, cudaMemcpyHostToDevice, ss[i]); });
  run("H2D 1 x 23MB", [&] { cudaMemcpyAsync(d, h, N * sz, cudaMemcpyHostToDevice, ss[0]); });
  run("D2H 32 x 105KB, 32 streams", [&] { for (int i = 0; i < 32; ++i) cudaMemcpyAsync(hp + i * 104544, dp + i * 104544, 104544, cudaMemcpyDeviceToHost, ss[i]); });
  run("H2D 64x360KB + D2H 32x105KB", [&] {
    for (int i = 0; i < 32; ++i) {
      cudaMemcpyAsync(d + 2 * i * sz, h + 2 * i * sz, sz, cudaMemcpyHostToDevice, ss[i]);
      cudaMemcpyAsync(d + (2 * i + 1) * sz, h + (2 * i + 1) * sz, sz, cudaMemcpyHostToDevice, ss[i]);
      cudaMemcpyAsync(hp + i * 104544, dp + i * 104544, 104544, cudaMemcpyDeviceToHost, ss[i]);
    } });
  // same with a 0.3 ms kernel chain per stream between H2D and D2H (30 kernels of 10 us)
  run("per stream: 2 H2D + 30 x 10us kernels + D2H", [&] {
    for (int i = 0; i < 32; ++i) {
      cudaMemcpyAsync(d + 2 * i * sz, h + 2 * i * sz, sz, cudaMemcpyHostToDevice, ss[i]);
      cudaMemcpyAsync(d + (2 * i + 1) * sz, h + (2 * i + 1) * sz, sz, cudaMemcpyHostToDevice, ss[i]);
      for (int k = 0; k < 30; ++k) spin<<<1, 32, 0, ss[i]>>>(19650);
      cudaMemcpyAsync(hp + i * 104544, dp + i * 104544, 104544, cudaMemcpyDeviceToHost, ss[i]);
    } });
  run("per stream: 30 x 10us kernels only", [&] {
    for (int i = 0; i < 32; ++i) for (int k = 0; k < 30; ++k) spin<<<1, 32, 0, ss[i]>>>(19650); });
  return 0;
}
