// decoder.cu — SAM mask-decoder tail: transposed-conv pixel shuffles fused with skip / LayerNorm2d /
// GELU, the hypernetwork mask product (the 32-channel up-scaled embedding is consumed on the fly and
// never written to HBM), small batched 3-layer MLP heads, and the mask / token selection epilogue.
#include <math.h>

#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.h"

namespace cg = cooperative_groups;

namespace ds2 {

__device__ __forceinline__ float gelu_erf_d(float x) { return gelu_erf_fast(x); }
__device__ __forceinline__ float warp_sum_d(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per OUTPUT pixel; C == 64 (2 channels per lane)
__global__ void __launch_bounds__(256) upscale1_kernel(const float* __restrict__ g, const float* __restrict__ bias,
                                                       const float* __restrict__ skip,
                                                       const float* __restrict__ lnw,
                                                       const float* __restrict__ lnb,
                                                       __nv_bfloat16* __restrict__ y, int B, int Hm, int Wm, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int Ho = Hm * 2, Wo = Wm * 2;
  if (wid >= static_cast<long long>(B) * Ho * Wo) return;
  const int X = static_cast<int>(wid % Wo);
  const int Y = static_cast<int>((wid / Wo) % Ho);
  const int b = static_cast<int>(wid / (static_cast<long long>(Wo) * Ho));
  const long long grow = (static_cast<long long>(b) * Hm + Y / 2) * Wm + X / 2;
  const int sub = (Y & 1) * 2 + (X & 1);
  const float* gp = g + grow * (4LL * C) + sub * C;
  const float* sp = skip + (static_cast<long long>(Y) * Wo + X) * C;
  float v[2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + i * 32;
    v[i] = gp[c] + bias[c] + sp[c];
    s += v[i];
  }
  const float mean = warp_sum_d(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) q += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum_d(q) / C + 1e-6f);
  __nv_bfloat16* yp = y + wid * C;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + i * 32;
    yp[c] = __float2bfloat16(gelu_erf_d((v[i] - mean) * rstd * lnw[c] + lnb[c]));
  }
}

// one thread per OUTPUT pixel; C == 32; M <= 4 hyper vectors per object held in shared memory
__global__ void __launch_bounds__(256) upscale2_masks_kernel(const float* __restrict__ g,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ skip,
                                                             const float* __restrict__ hyper,
                                                             float* __restrict__ masks, int B, int Hm, int Wm,
                                                             int M) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  constexpr int C = 32;
  __shared__ float sh[4 * C + C];
  const int Ho = Hm * 2, Wo = Wm * 2;
  const long long per_obj = static_cast<long long>(Ho) * Wo;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < M * C; i += blockDim.x) sh[i] = hyper[static_cast<long long>(b) * M * C + i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sh[4 * C + i] = bias[i];
  __syncthreads();
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= per_obj) return;
  const int X = static_cast<int>(pix % Wo), Y = static_cast<int>(pix / Wo);
  const long long grow = (static_cast<long long>(b) * Hm + Y / 2) * Wm + X / 2;
  const int sub = (Y & 1) * 2 + (X & 1);
  const float4* gp = reinterpret_cast<const float4*>(g + grow * (4LL * C) + sub * C);
  const float4* sp = reinterpret_cast<const float4*>(skip + pix * C);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    const float4 a = gp[i], s = sp[i];
    float u[4] = {a.x + s.x + sh[4 * C + 4 * i], a.y + s.y + sh[4 * C + 4 * i + 1],
                  a.z + s.z + sh[4 * C + 4 * i + 2], a.w + s.w + sh[4 * C + 4 * i + 3]};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float t = gelu_erf_d(u[e]);
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (m < M) acc[m] = fmaf(t, sh[m * C + 4 * i + e], acc[m]);
    }
  }
  for (int m = 0; m < M; ++m) masks[(static_cast<long long>(b) * M + m) * per_obj + pix] = acc[m];
}

// batched 3-layer MLP, one CTA (256 threads) per row item
struct Mlp3Params {
  const float* x;
  long long ldx;
  const int* gather;
  int rows, nmlp, din, dh, dout;
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int sigmoid_out;
  float* y;
  long long ldy;
  int prefetch;
};
// One dense layer of the batched MLP inside a CTA: y[o] = act(sum_k W[k][o] x[k] + b[o]).  W is INPUT-major
// ([in][out], the transpose of nn.Linear's layout), so the threads of a warp (consecutive outputs o) read
// consecutive weights; when there are fewer outputs than threads the K range is sliced over the spare threads
// and the slices are summed through shared memory.  (The previous version walked nn.Linear rows with one warp
// per output and a shuffle reduction: 96 dependent L2 round trips per CTA, 77 us per launch.)
__device__ __forceinline__ void mlp_layer(const float* __restrict__ W, const float* __restrict__ bias, const float* x,
                                          float* y, float* part, int din, int dout, int act) {
  int P = 1;
  while (P < dout) P <<= 1;              // outputs padded to a power of two (<= blockDim)
  const int G = blockDim.x / P;          // K slices
  const int o = threadIdx.x % P, g = threadIdx.x / P;
  const int per = (din + G - 1) / G;
  const int k0 = g * per, k1 = min(din, k0 + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (o < dout && g < G) {
    int k = k0;
    // 32 independent weight loads in flight per thread before the first FMA needs one (the loop is otherwise
    // a chain of L2 round trips: the weights are read once per CTA)
    for (; k + 31 < k1; k += 32) {
      float wv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) wv[i] = __ldg(W + static_cast<long long>(k + i) * dout + o);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        a0 = fmaf(wv[i], x[k + i], a0);
        a1 = fmaf(wv[i + 1], x[k + i + 1], a1);
        a2 = fmaf(wv[i + 2], x[k + i + 2], a2);
        a3 = fmaf(wv[i + 3], x[k + i + 3], a3);
      }
    }
    for (; k < k1; ++k) a0 = fmaf(__ldg(W + static_cast<long long>(k) * dout + o), x[k], a0);
  }
  part[threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < dout) {
    float a = 0.f;
    for (int gg = 0; gg < G; ++gg) a += part[gg * P + threadIdx.x];
    a += bias[threadIdx.x];
    if (act == 1) a = fmaxf(a, 0.f);
    else if (act == 2) a = 1.f / (1.f + expf(-a));
    y[threadIdx.x] = a;
  }
  __syncthreads();
}

// L2 prefetch of a weight matrix, sliced over the `parts` CTAs that share it.
__device__ __forceinline__ void prefetch_l2_slice(const float* w, long long bytes, int part, int parts) {
  const char* base = reinterpret_cast<const char*>(w);
  for (long long off = (static_cast<long long>(threadIdx.x) * parts + part) * 128; off < bytes;
       off += static_cast<long long>(blockDim.x) * parts * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
}

__global__ void __launch_bounds__(256) mlp3_kernel(const Mlp3Params p) {
  {
    // The weights are constants, so they may be touched BEFORE the programmatic-dependent-launch wait: the frame's 1 GB
    // working set has evicted them from L2 since the last frame, and the three layers are a dependent chain that used to
    // pay a DRAM round trip per 32-load batch (8 batches per layer, 37 us per launch).  Pull all three matrices into L2
    // while the preceding kernel drains.
    const int set = blockIdx.x % p.nmlp, part = blockIdx.x / p.nmlp, parts = (p.rows + p.nmlp - 1) / p.nmlp;
    if (p.prefetch) {
      prefetch_l2_slice(p.w1 + static_cast<long long>(set) * p.dh * p.din, 4LL * p.dh * p.din, part, parts);
      prefetch_l2_slice(p.w2 + static_cast<long long>(set) * p.dh * p.dh, 4LL * p.dh * p.dh, part, parts);
      prefetch_l2_slice(p.w3 + static_cast<long long>(set) * p.dout * p.dh, 4LL * p.dout * p.dh, part, parts);
    }
  }
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  extern __shared__ float sm[];
  float* xin = sm;             // din
  float* h1 = xin + p.din;     // dh
  float* h2 = h1 + p.dh;       // dh
  float* part = h2 + p.dh;     // blockDim partial sums
  float* yout = part + 256;    // dout
  const int item = blockIdx.x;
  const int set = item % p.nmlp;
  const long long src = p.gather ? p.gather[item] : item;
  for (int i = threadIdx.x; i < p.din; i += blockDim.x) xin[i] = p.x[src * p.ldx + i];
  __syncthreads();
  mlp_layer(p.w1 + static_cast<long long>(set) * p.dh * p.din, p.b1 + set * p.dh, xin, h1, part, p.din, p.dh, 1);
  mlp_layer(p.w2 + static_cast<long long>(set) * p.dh * p.dh, p.b2 + set * p.dh, h1, h2, part, p.dh, p.dh, 1);
  mlp_layer(p.w3 + static_cast<long long>(set) * p.dout * p.dh, p.b3 + set * p.dout, h2, yout, part, p.dh, p.dout,
            p.sigmoid_out ? 2 : 0);
  for (int o = threadIdx.x; o < p.dout; o += blockDim.x) p.y[static_cast<long long>(item) * p.ldy + o] = yout[o];
}

// ---- cluster version of the batched 3-layer MLP (the default) -------------------------------------------------------
// mlp3_kernel above gives every row its own CTA, which then streams all three weight matrices (768 KB) through one SM in
// three dependent phases: ~30 us per launch, four launches per frame.  Here a cluster of 8 CTAs owns up to 16 rows of one
// MLP: CTA c computes 1/8 of each layer's outputs for all 16 rows (its 32-column weight slice is 32 KB and is read once),
// the slices are exchanged through distributed shared memory between layers, and the first thing every CTA does — before
// the programmatic-dependent-launch wait, the weights being constants — is to prefetch its three slices into L2.
constexpr int kMcRows = 16, kMcCtas = 8;

// One layer slice: out[r][oo] = act(sum_k W[k][col0 + oo] x[r][k] + bias[col0 + oo]) for this CTA's ncols columns.
// Threads = P column lanes (ncols padded to a power of two) x G k-slices; the k-slices are summed through `part`.
__device__ __forceinline__ void mc_layer(const float* __restrict__ W, int ldw, const float* __restrict__ bias, int col0,
                                         int ncols, const float* xin, int din, float* part, float* out, int act) {
  int P = 1;
  while (P < ncols) P <<= 1;
  const int G = 256 / P;
  const int o = threadIdx.x % P, g = threadIdx.x / P;
  const int per = (din + G - 1) / G;
  const int k0 = g * per, k1 = min(din, k0 + per);
  float acc[kMcRows];
#pragma unroll
  for (int r = 0; r < kMcRows; ++r) acc[r] = 0.f;
  if (o < ncols) {
    for (int k = k0; k < k1; k += 16) {
      float wv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) wv[i] = (k + i < k1) ? __ldg(W + static_cast<long long>(k + i) * ldw + col0 + o) : 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (k + i < k1) {
#pragma unroll
          for (int r = 0; r < kMcRows; ++r) acc[r] = fmaf(wv[i], xin[r * din + k + i], acc[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kMcRows; ++r) part[(g * kMcRows + r) * P + o] = acc[r];
  __syncthreads();
  for (int e = threadIdx.x; e < kMcRows * ncols; e += 256) {
    const int r = e / ncols, oo = e - r * ncols;
    float a = 0.f;
    for (int gg = 0; gg < G; ++gg) a += part[(gg * kMcRows + r) * P + oo];
    a += bias[col0 + oo];
    if (act == 1) a = fmaxf(a, 0.f);
    else if (act == 2) a = 1.f / (1.f + expf(-a));
    out[e] = a;
  }
  __syncthreads();
}

__global__ void __cluster_dims__(kMcCtas, 1, 1) __launch_bounds__(256) mlp3_cluster_kernel(const Mlp3Params p) {
  extern __shared__ __align__(16) float msm[];
  const int wmax = p.din > p.dh ? p.din : p.dh;
  float* full0 = msm;                          // [16][max(din, dh)]  layer input (x, then h2)
  float* full1 = full0 + kMcRows * wmax;       // [16][dh]            h1
  float* part = full1 + kMcRows * p.dh;        // [256][16]
  float* sl0 = part + 256 * kMcRows;           // [16][dh / 8]  this CTA's slice of h1
  float* sl1 = sl0 + kMcRows * (p.dh / kMcCtas);   // [16][dh / 8]  ... of h2
  float* sl2 = sl1 + kMcRows * (p.dh / kMcCtas);   // [16][ceil(dout / 8)]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int set = blockIdx.y;
  const int rows_per_set = p.rows / p.nmlp;
  const int row0 = blockIdx.z * kMcRows;
  const int nh = p.dh / kMcCtas;                       // hidden columns per CTA
  const int n3 = (p.dout + kMcCtas - 1) / kMcCtas;     // output columns per CTA
  const int c3 = rank * n3;
  const int nc3 = max(0, min(n3, p.dout - c3));
  const float* w1 = p.w1 + static_cast<long long>(set) * p.dh * p.din;
  const float* w2 = p.w2 + static_cast<long long>(set) * p.dh * p.dh;
  const float* w3 = p.w3 + static_cast<long long>(set) * p.dout * p.dh;
  if (p.prefetch) {
    // weight slices are constants: pull them into L2 while the preceding kernel drains (one 128-byte line per row and slice)
    for (int k = threadIdx.x; k < p.din; k += 256)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(w1 + static_cast<long long>(k) * p.dh + rank * nh));
    for (int k = threadIdx.x; k < p.dh; k += 256) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(w2 + static_cast<long long>(k) * p.dh + rank * nh));
      if (nc3 > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(w3 + static_cast<long long>(k) * p.dout + c3));
    }
  }
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  for (int e = threadIdx.x; e < kMcRows * p.din; e += 256) {
    const int r = e / p.din, c = e - r * p.din;
    float v = 0.f;
    if (row0 + r < rows_per_set) {
      const int item = (row0 + r) * p.nmlp + set;
      const long long src = p.gather ? p.gather[item] : item;
      v = p.x[src * p.ldx + c];
    }
    full0[e] = v;
  }
  __syncthreads();
  auto gather_slices = [&](float* local_slice, float* dst) {
    cluster.sync();   // every CTA's slice is written
    for (int e = threadIdx.x; e < kMcRows * p.dh; e += 256) {
      const int r = e / p.dh, c = e - r * p.dh;
      const float* remote = cluster.map_shared_rank(local_slice, c / nh);
      dst[e] = remote[r * nh + (c % nh)];
    }
    __syncthreads();
  };
  mc_layer(w1, p.dh, p.b1 + set * p.dh, rank * nh, nh, full0, p.din, part, sl0, 1);
  gather_slices(sl0, full1);
  mc_layer(w2, p.dh, p.b2 + set * p.dh, rank * nh, nh, full1, p.dh, part, sl1, 1);
  gather_slices(sl1, full0);
  mc_layer(w3, p.dout, p.b3 + set * p.dout, c3, nc3, full0, p.dh, part, sl2, p.sigmoid_out ? 2 : 0);
  for (int e = threadIdx.x; e < kMcRows * nc3; e += 256) {
    const int r = e / nc3, oo = e - r * nc3;
    if (row0 + r < rows_per_set) {
      const int item = (row0 + r) * p.nmlp + set;
      p.y[static_cast<long long>(item) * p.ldy + c3 + oo] = sl2[e];
    }
  }
  cluster.sync();   // nobody exits while a peer may still read its h2 slice
}

// mask / token selection: blockIdx.x = object; with multimask output the choice needs no reduction over the mask, so the
// copy of the chosen 256^2 mask is sliced over gridDim.y CTAs (one CTA per object left 132 SMs idle for 64 us); the
// stability test of the single-mask path counts over the whole mask and keeps one CTA per object (gridDim.y == 1)
__global__ void __launch_bounds__(256) sam_select_kernel(const float* __restrict__ all_masks,
                                                         const float* __restrict__ ious,
                                                         const float* __restrict__ obj_score,
                                                         const float* __restrict__ mask_tokens, int B, int S, int C,
                                                         int multimask, float delta, float thresh,
                                                         float* __restrict__ low_res, float* __restrict__ iou_out,
                                                         int* __restrict__ best_idx, float* __restrict__ token_out) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  __shared__ int s_i, s_u, s_idx;
  const int b = blockIdx.x;
  const long long n = static_cast<long long>(S) * S;
  if (threadIdx.x == 0) {
    s_i = 0;
    s_u = 0;
  }
  __syncthreads();
  const float* io = ious + b * 4;
  int best_multi = 1;
  {
    float bv = io[1];
    if (io[2] > bv) { bv = io[2]; best_multi = 2; }
    if (io[3] > bv) { bv = io[3]; best_multi = 3; }
  }
  if (!multimask) {
    const float* m0 = all_masks + static_cast<long long>(b) * 4 * n;
    int ci = 0, cu = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = m0[i];
      ci += v > delta;
      cu += v > -delta;
    }
    for (int o = 16; o > 0; o >>= 1) {
      ci += __shfl_xor_sync(0xffffffffu, ci, o);
      cu += __shfl_xor_sync(0xffffffffu, cu, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_i, ci);
      atomicAdd(&s_u, cu);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int idx;
    if (multimask) {
      idx = best_multi;
    } else {
      const float stab = s_u > 0 ? static_cast<float>(s_i) / static_cast<float>(s_u) : 1.0f;
      idx = (stab >= thresh) ? 0 : best_multi;
    }
    s_idx = idx;
    if (blockIdx.y == 0) {
      best_idx[b] = idx;
      iou_out[b] = io[idx];
    }
  }
  __syncthreads();
  const int idx = s_idx;
  const bool present = obj_score[b] > 0.f;
  const float4* src = reinterpret_cast<const float4*>(all_masks + (static_cast<long long>(b) * 4 + idx) * n);
  float4* dst = reinterpret_cast<float4*>(low_res + static_cast<long long>(b) * n);
  const float4 absent = make_float4(-1024.0f, -1024.0f, -1024.0f, -1024.0f);
  const long long n4 = n / 4;      // S is a multiple of 4 (checked by the host side)
  for (long long i = static_cast<long long>(blockIdx.y) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.y) * blockDim.x)
    dst[i] = present ? src[i] : absent;
  if (blockIdx.y != 0) return;
  // object-pointer token: the matching multimask token, or token 0 in single-mask mode
  const int tok = multimask ? idx : 0;
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    token_out[static_cast<long long>(b) * C + i] = mask_tokens[(static_cast<long long>(b) * 4 + tok) * C + i];
}

__global__ void objptr_mix_kernel(float* __restrict__ ptr, const float* __restrict__ obj_score,
                                  const float* __restrict__ no_obj_ptr, int B, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const float lam = obj_score[i / C] > 0.f ? 1.f : 0.f;
  ptr[i] = lam * ptr[i] + (1.f - lam) * no_obj_ptr[i % C];
}

}  // namespace ds2

extern "C" {

int ds2_upscale1(const float* g, const float* bias, const float* skip, const float* ln_w, const float* ln_b,
                 void* y_bf16, int32_t B, int32_t Hm, int32_t Wm, int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(g && bias && skip && ln_w && ln_b && y_bf16 && B > 0, DS2_E_ARG, "ds2_upscale1: bad args");
  DS2_REQUIRE(C == 64, DS2_E_ARG, "ds2_upscale1: C must be 64 (got %d)", C);
  const long long warps = static_cast<long long>(B) * Hm * 2 * Wm * 2;
  DS2_LAUNCH((upscale1_kernel), static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, as_stream(stream), 
      g, bias, skip, ln_w, ln_b, reinterpret_cast<__nv_bfloat16*>(y_bf16), B, Hm, Wm, C);
  return post_launch("upscale1_kernel");
}

int ds2_upscale2_masks(const float* g, const float* bias, const float* skip, const float* hyper, float* masks,
                       int32_t B, int32_t Hm, int32_t Wm, int32_t C, int32_t M, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(g && bias && skip && hyper && masks && B > 0, DS2_E_ARG, "ds2_upscale2_masks: bad args");
  DS2_REQUIRE(C == 32 && M >= 1 && M <= 4, DS2_E_ARG, "ds2_upscale2_masks: C must be 32 and M <= 4");
  const long long per_obj = static_cast<long long>(Hm) * 2 * Wm * 2;
  dim3 grid(static_cast<unsigned>((per_obj + 255) / 256), B);
  DS2_LAUNCH((upscale2_masks_kernel), grid, 256, 0, as_stream(stream), g, bias, skip, hyper, masks, B, Hm, Wm, M);
  return post_launch("upscale2_masks_kernel");
}

int ds2_mlp3(const ds2_mlp3_args* a, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a && a->x && a->w1 && a->b1 && a->w2 && a->b2 && a->w3 && a->b3 && a->y, DS2_E_ARG,
              "ds2_mlp3: null pointer");
  DS2_REQUIRE(a->rows > 0 && a->nmlp > 0 && a->din > 0 && a->dh > 0 && a->dout > 0, DS2_E_ARG, "ds2_mlp3: bad dims");
  Mlp3Params p;
  p.x = a->x;
  p.ldx = a->ldx;
  p.gather = a->gather;
  p.rows = a->rows;
  p.nmlp = a->nmlp;
  p.din = a->din;
  p.dh = a->dh;
  p.dout = a->dout;
  p.w1 = a->w1;
  p.b1 = a->b1;
  p.w2 = a->w2;
  p.b2 = a->b2;
  p.w3 = a->w3;
  p.b3 = a->b3;
  p.sigmoid_out = a->sigmoid_out;
  p.y = a->y;
  p.ldy = a->ldy;
  static const int prefetch = [] {
    const char* e = getenv("DS2_MLP_PREFETCH");   // A/B switch (tuning only)
    return (e && e[0] == '0') ? 0 : 1;
  }();
  p.prefetch = prefetch;
  DS2_REQUIRE(a->dh <= 256 && a->dout <= 256, DS2_E_ARG, "ds2_mlp3: hidden / output width must be <= 256 (got %d, %d)",
              a->dh, a->dout);
  static const bool use_cluster = [] {
    const char* e = getenv("DS2_MLP_CLUSTER");   // DS2_MLP_CLUSTER=0: one CTA per row (A/B)
    return !(e && e[0] == '0');
  }();
  if (use_cluster && (a->dh % kMcCtas) == 0 && (a->rows % a->nmlp) == 0 && a->nmlp <= 65535) {
    const int wmax = a->din > a->dh ? a->din : a->dh;
    const int n3 = (a->dout + kMcCtas - 1) / kMcCtas;
    const int csmem = (kMcRows * wmax + kMcRows * a->dh + 256 * kMcRows + 2 * kMcRows * (a->dh / kMcCtas) + kMcRows * n3) * 4;
    static int attr_smem = 0;
    if (csmem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(mlp3_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, csmem);
      DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_mlp3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_smem = csmem;
    }
    const int row_blocks = (a->rows / a->nmlp + kMcRows - 1) / kMcRows;
    DS2_LAUNCH((mlp3_cluster_kernel), dim3(kMcCtas, a->nmlp, row_blocks), 256, csmem, as_stream(stream), p);
    return post_launch("mlp3_cluster_kernel");
  }
  const int smem = (a->din + 2 * a->dh + 256 + a->dout) * 4;
  DS2_LAUNCH((mlp3_kernel), a->rows, 256, smem, as_stream(stream), p);
  return post_launch("mlp3_kernel");
}

int ds2_sam_select(const float* all_masks, const float* ious, const float* obj_score, const float* mask_tokens,
                   int32_t B, int32_t S, int32_t C, int32_t multimask, float stab_delta, float stab_thresh,
                   float* low_res, float* iou_out, int32_t* best_idx, float* token_out, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(all_masks && ious && obj_score && mask_tokens && low_res && iou_out && best_idx && token_out && B > 0,
              DS2_E_ARG, "ds2_sam_select: bad args");
  DS2_REQUIRE((S % 2) == 0, DS2_E_ARG, "ds2_sam_select: mask side %d must be even", S);
  const dim3 grid(B, multimask ? 16 : 1);
  DS2_LAUNCH((sam_select_kernel), grid, 256, 0, as_stream(stream), all_masks, ious, obj_score, mask_tokens, B, S, C, multimask,
                                                     stab_delta, stab_thresh, low_res, iou_out, best_idx, token_out);
  return post_launch("sam_select_kernel");
}

int ds2_objptr_mix(float* ptr, const float* obj_score, const float* no_obj_ptr, int32_t B, int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(ptr && obj_score && no_obj_ptr && B > 0 && C > 0, DS2_E_ARG, "ds2_objptr_mix: bad args");
  DS2_LAUNCH((objptr_mix_kernel), (B * C + 255) / 256, 256, 0, as_stream(stream), ptr, obj_score, no_obj_ptr, B, C);
  return post_launch("objptr_mix_kernel");
}

}  // extern "C"
