// mma_rate.cu — measures issue-to-retire cycles of back-to-back tcgen05.mma (cta_group::1, kind::f16, bf16) for
// several shapes and operand sources, all SMs busy (one CTA per SM), no loads: operands are whatever is
// in shared memory / TMEM.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I det-sam2_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate -lcuda && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "tc05.cuh"


// mode 0: SS (A and B in smem), 1: TS (A in TMEM), 2: alternating QK-like SS N=n and PV-like TS N=64
template <bool ELECT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 200 * 1024;
  const uint32_t slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 0 && (ELECT ? tc::elect_one() : lane == 0)) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint32_t idesc_pv = tc::make_idesc_bf16(128, 64, 0, 1);
    const uint32_t sa = base, sb = base + 64 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t da = tc::make_desc_sw128(sa + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024);
        const uint64_t db = tc::make_desc_sw128(sb + (k >> 2) * (N * 128) + (k & 3) * 32, 16, 1024);
        if (mode == 0) tc::umma_ss(tmem, da, db, idesc, 1u);
        else tc::umma_ts(tmem, tmem + 384 + k * 8, db, idesc, 1u);
      }
      if (mode == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t db = tc::make_desc_sw128(sb + k * 2048, 128 * 128, 1024);
          tc::umma_ts(tmem + 256, tmem + k * 8, db, idesc_pv, 1u);
        }
      }
    }
    tc::umma_commit(bar);
    tc::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 200 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 200;
  const char* names[3] = {"SS", "TS(A in TMEM)", "TS QK + TS PV(N=64,K=128)"};
  for (int el = 0; el < 2; ++el)
  for (int mode = 0; mode < 3; ++mode) {
    for (int N : {64, 128, 256}) {
      if (mode == 2 && N != 128) continue;
      if (mode != 0 && N == 256) {}
      for (int grid : {148}) {
        if (el) rate_kernel<true><<<grid, 128, smem>>>(N, mode, iters, d);
        else rate_kernel<false><<<grid, 128, smem>>>(N, mode, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("error %s\n", cudaGetErrorString(e));
          return 1;
        }
        long long c;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double per16 = double(c) / iters;
        const int n_mma = mode == 2 ? 24 : 16;
        printf("%s %-28s M=128 N=%3d grid=%3d: %8.1f clk per group of %d MMAs (%.1f clk per K=16 QK MMA; formula %d)\n",
               el ? "elect.sync" : "lane==0   ", names[mode], N, grid, per16, n_mma, mode == 2 ? (per16 - 8 * 32) / 16 : per16 / 16, N / 2);
      }
    }
  }
  return 0;
}
