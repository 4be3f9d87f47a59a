// Floor of ONE kernel round trip on the box (launch -> tiny kernel -> completion seen by the host), with
// cudaStreamSynchronize and with the host spinning on a flag the kernel writes into mapped memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe tools/latency_floor_probe.cu && ./lat_probe
// (profiles/r02_s21_latency_floor.txt; DESIGN section 8, "Single calls")
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(volatile unsigned* flag, unsigned v, int spin) {
  // a little dependent work to look like a small kernel
  float x = threadIdx.x;
  for (int i = 0; i < spin; i++) x = x * 1.0001f + 0.5f;
  if (threadIdx.x == 0) { if (x == 12345.f) v++; *flag = v; }
}
int main() {
  unsigned* h; unsigned* d;
  cudaHostAlloc(&h, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&d, h, 0);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  for (int spin : {0, 2000, 8000}) {
    for (int mode = 0; mode < 2; mode++) {
      *h = 0;
      for (int w = 0; w < 200; w++) { k<<<1, 32, 0, st>>>(d, w + 1, spin); cudaStreamSynchronize(st); }
      const int N = 5000;
      auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < N; i++) {
        const unsigned v = 1000 + i;
        k<<<1, 32, 0, st>>>(d, v, spin);
        if (mode == 0) cudaStreamSynchronize(st);
        else while (*(volatile unsigned*)h != v) { __builtin_ia32_pause(); }
      }
      cudaStreamSynchronize(st);
      double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / N;
      printf("spin %5d  %s: %.2f us per call\n", spin, mode == 0 ? "cudaStreamSynchronize" : "spin on mapped flag  ", us);
    }
  }
  return 0;
}
