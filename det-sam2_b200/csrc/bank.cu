// bank.cu — table-driven assembly of the step's memory bank (sam2_base.py:564-650).
//
// The stored memories of a session are separate per-frame device buffers (the predictor's dict
// schema); which of them a step attends to changes every frame.  Instead of one gather launch per
// stored frame with the source pointer baked into the launch (which defeats CUDA-graph replay), the
// host writes a small DEVICE-resident table per step — source pointers, temporal-embedding indices,
// pointer distances — and ONE launch pair with static arguments builds
//   kin[b, f*T + t, :] = bf16(mem_f[b,t,:] + pos[t,:] + tpos[idx_f,:])     val[...] = mem_f[b,t,:]
//   kin[b, nf*T + 4*j + q, :] = bf16(ptr_j[b, 64q:64q+64] + W.sine1d(dist_j) + bias),  val = bf16(ptr_j)
#include "common.h"

namespace ds2 {

__global__ void __launch_bounds__(256)
bank_frames_kernel(const __nv_bfloat16* const* __restrict__ src, const int* __restrict__ tpos_idx,
                   const float* __restrict__ pos, const float* __restrict__ tpos_table,
                   __nv_bfloat16* __restrict__ kin, __nv_bfloat16* __restrict__ val, int B, int T, int C,
                   long long dst_bs) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int f = blockIdx.y;
  const __nv_bfloat16* __restrict__ mem = src[f];
  const float* __restrict__ tp_row = tpos_table + static_cast<long long>(tpos_idx[f]) * C;
  const int C8 = C / 8;
  const long long n = static_cast<long long>(B) * T * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    const long long r = i / C8;
    const int t = static_cast<int>(r % T);
    const int b = static_cast<int>(r / T);
    const uint4 m = __ldg(reinterpret_cast<const uint4*>(mem + r * C + c8 * 8));
    const __nv_bfloat16* me = reinterpret_cast<const __nv_bfloat16*>(&m);
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * C + c8 * 8);
    const float4* tp = reinterpret_cast<const float4*>(tp_row + c8 * 8);
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1), t0 = __ldg(tp), t1 = __ldg(tp + 1);
    const float add[8] = {p0.x + t0.x, p0.y + t0.y, p0.z + t0.z, p0.w + t0.w,
                          p1.x + t1.x, p1.y + t1.y, p1.z + t1.z, p1.w + t1.w};
    uint4 ko;
    __nv_bfloat16* ke = reinterpret_cast<__nv_bfloat16*>(&ko);
#pragma unroll
    for (int e = 0; e < 8; ++e) ke[e] = __float2bfloat16(__bfloat162float(me[e]) + add[e]);
    const long long dst = static_cast<long long>(b) * dst_bs + (static_cast<long long>(f) * T + t) * C + c8 * 8;
    *reinterpret_cast<uint4*>(kin + dst) = ko;
    *reinterpret_cast<uint4*>(val + dst) = m;
  }
}

// one block per (pointer j, object b); 256 threads = the 256 pointer channels
__global__ void __launch_bounds__(256)
bank_ptrs_kernel(const float* const* __restrict__ src, const float* __restrict__ dist,
                 const float* __restrict__ w /*[64,256]*/, const float* __restrict__ bias /*[64]*/,
                 __nv_bfloat16* __restrict__ kin, __nv_bfloat16* __restrict__ val, long long dst_bs,
                 long long row0) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  __shared__ float pe[256];
  __shared__ float tp[64];
  const int j = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  {
    // sam2_utils.py:69-79 with dim 256: dim_t[i] = 10000^(2*(i/2)/128); [sin(pos/dim_t) | cos(pos/dim_t)]
    const int i = t & 127;
    const float dim_t = powf(10000.f, static_cast<float>(2 * (i / 2)) / 128.f);
    const float a = dist[j] / dim_t;
    pe[t] = (t < 128) ? sinf(a) : cosf(a);
  }
  __syncthreads();
  {
    const int c = t >> 2, part = t & 3;
    float acc = 0.f;
    for (int k = part * 64; k < part * 64 + 64; ++k) acc = fmaf(w[c * 256 + k], pe[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) tp[c] = acc + bias[c];
  }
  __syncthreads();
  const float pv = src[j][static_cast<long long>(b) * 256 + t];
  const long long dst = static_cast<long long>(b) * dst_bs + (row0 + 4ll * j + t / 64) * 64 + (t % 64);
  kin[dst] = __float2bfloat16(pv + tp[t % 64]);
  val[dst] = __float2bfloat16(pv);
}

}  // namespace ds2

extern "C" int ds2_bank_assemble(const void* const* frame_src, const int32_t* frame_tpos, int32_t nf,
                                 const float* const* ptr_src, const float* ptr_dist, int32_t np,
                                 const float* pos, const float* tpos_table, const float* ptr_w,
                                 const float* ptr_bias, void* kin_bf16, void* val_bf16, int32_t B, int32_t T,
                                 int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(kin_bf16 && val_bf16 && B > 0 && T > 0 && C > 0 && (C % 8) == 0 && nf >= 0 && np >= 0 && nf + np > 0,
              DS2_E_ARG, "ds2_bank_assemble: bad args");
  DS2_REQUIRE(np == 0 || C == 64, DS2_E_ARG, "ds2_bank_assemble: pointer tokens need mem_dim 64");
  const long long dst_bs = (static_cast<long long>(nf) * T + 4ll * np) * C;
  int rc = DS2_OK;
  if (nf > 0) {
    DS2_REQUIRE(frame_src && frame_tpos && pos && tpos_table, DS2_E_ARG, "ds2_bank_assemble: null frame table");
    const long long n = static_cast<long long>(B) * T * (C / 8);
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;  // grid-stride: 8 resident CTAs per SM
    dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(nf));
    DS2_LAUNCH((bank_frames_kernel), grid, 256, 0, as_stream(stream), 
        reinterpret_cast<const __nv_bfloat16* const*>(frame_src), frame_tpos, pos, tpos_table,
        reinterpret_cast<__nv_bfloat16*>(kin_bf16), reinterpret_cast<__nv_bfloat16*>(val_bf16), B, T, C, dst_bs);
    rc = post_launch("bank_frames_kernel");
    if (rc) return rc;
  }
  if (np > 0) {
    DS2_REQUIRE(ptr_src && ptr_dist && ptr_w && ptr_bias, DS2_E_ARG, "ds2_bank_assemble: null pointer table");
    dim3 grid(static_cast<unsigned>(np), static_cast<unsigned>(B));
    DS2_LAUNCH((bank_ptrs_kernel), grid, 256, 0, as_stream(stream), ptr_src, ptr_dist, ptr_w, ptr_bias,
                                                         reinterpret_cast<__nv_bfloat16*>(kin_bf16),
                                                         reinterpret_cast<__nv_bfloat16*>(val_bf16), dst_bs,
                                                         static_cast<long long>(nf) * T);
    rc = post_launch("bank_ptrs_kernel");
  }
  return rc;
}
