// MUFU.EX2 issue rate on sm_100a: f32 vs packed f16x2 / bf16x2 (does a packed op deliver two results per MUFU slot?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ex2_rate tools/ex2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned seed) {
  unsigned a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned)(t1 - t0);
}

template <int MODE>
void run(const char* name) {
  unsigned* d;
  cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 2000;
  k<MODE><<<148 * 8, 256>>>(d, iters, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned clk;
  cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
  // per SM: 8 CTAs x 8 warps x iters x 8 warp-instructions
  const double winst = 8.0 * 8 * iters * 8;
  printf("%-10s %8.3f ms  CTA0 %u clk  -> %.2f clk per warp-instruction per SM (4 SMSPs), %s\n", name, ms, clk,
         clk / (winst / 1.0) * 1.0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0>("f32");
  run<1>("f16x2");
  run<2>("bf16x2");
  return 0;
}
