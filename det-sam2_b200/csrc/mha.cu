// mha.cu — generic multi-head flash attention for the small / oddly-shaped attentions of the path:
// Hiera window + global attention (head_dim 72/56/96, windows 4..16, 2x2 query max-pool, zero-pad
// windows) and the two-way mask-decoder attentions (head_dim 16/32, 7-9 tokens x 4096 pixels).
//
// CTA = 4 warps = 64 query rows of one (batch x window, head); keys are streamed through shared
// memory in chunks of 64; S = Q K^T and O += P V on mma.sync.m16n8k16 (bf16 in, f32 accumulate) with
// the head dim zero-padded to a multiple of 16; online softmax in registers with quad shuffles.
// (These are <2 % of the frame's FLOPs; the d=256 memory attention uses the tcgen05 kernel in
// flash_tc.cu.)
#include <math.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.h"

namespace cg = cooperative_groups;

namespace ds2 {

struct MhaParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* out;
  long long q_tok, k_tok, v_tok, o_tok;
  long long q_bs, k_bs, v_bs, o_bs;
  int B, H, D;
  int Lq, Lk;
  int window, Hm, Wm, q_pool;
  int nwy, nwx;
  int Lk_valid;
  float scale_log2;
  const __nv_bfloat16* pad_q;
  const __nv_bfloat16* pad_k;
  const __nv_bfloat16* pad_v;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// Address of token (row index inside the sequence) for q / k / v, or nullptr for a padded token.
struct SeqGeom {
  int b, wy, wx;
  int Lq, Lk;
};

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
               "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// NW warps = 16*NW query rows per CTA; K/V chunks of 64 keys are double-buffered with cp.async so the
// loads of chunk j+1 overlap the MMAs of chunk j.
template <int DP, int NW>
__global__ void __launch_bounds__(32 * NW, (DP <= 96 ? (NW == 8 ? 2 : 4) : 1)) mha_kernel(const MhaParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  constexpr int QS = DP + 8;   // smem row stride (elements) for Q / K / V tiles (conflict-free for ldmatrix)
  constexpr int QT = 16 * NW;  // query rows per CTA
  constexpr int NT = 32 * NW;  // threads
  extern __shared__ __align__(16) uint8_t smem_mha[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_mha);
  __nv_bfloat16* KV = Qs + QT * QS;   // [2 buffers][K 64 rows | V 64 rows][QS]; V row-major, read via ldmatrix.trans

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y;
  const int w = p.window;
  int b, wy = 0, wx = 0, Lq, Lk;
  if (w > 0) {
    const int per = p.nwy * p.nwx;
    b = blockIdx.x / per;
    const int wi = blockIdx.x % per;
    wy = wi / p.nwx;
    wx = wi % p.nwx;
    Lk = w * w;
    Lq = p.q_pool ? (w / 2) * (w / 2) : Lk;
  } else {
    b = blockIdx.x;
    Lq = p.Lq;
    Lk = p.Lk;
  }
  const int q0 = blockIdx.z * QT;
  if (q0 >= Lq) return;
  const int Lk_valid = (p.Lk_valid > 0 && w == 0) ? p.Lk_valid : Lk;
  const int DV8 = p.D / 8;  // 16-byte vectors per head row

  auto tok_ptr = [&](const __nv_bfloat16* base, long long bs, long long ts, const __nv_bfloat16* pad,
                     int r) -> const __nv_bfloat16* {
    if (w == 0) return base + b * bs + static_cast<long long>(r) * ts + h * p.D;
    const int gy = wy * w + r / w, gx = wx * w + r % w;
    if (gy < p.Hm && gx < p.Wm) return base + b * bs + (static_cast<long long>(gy) * p.Wm + gx) * ts + h * p.D;
    return pad ? pad + h * p.D : nullptr;
  };

  // ---- stage Q tile (rows >= Lq and dims >= D are zero) ----
  for (int i = tid; i < QT * (DP / 8); i += NT) {
    const int r = i / (DP / 8), c = i % (DP / 8);
    uint4 val = make_uint4(0, 0, 0, 0);
    const int qr = q0 + r;
    if (qr < Lq && c < DV8) {
      if (w > 0 && p.q_pool) {
        const int hw = w / 2;
        const int py = qr / hw, px = qr % hw;
        bool first = true;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const __nv_bfloat16* src = tok_ptr(p.q, p.q_bs, p.q_tok, p.pad_q, (2 * py + dy) * w + 2 * px + dx);
            uint4 x = make_uint4(0, 0, 0, 0);
            if (src) x = *reinterpret_cast<const uint4*>(src + c * 8);
            val = first ? x : bf16x8_max(val, x);
            first = false;
          }
      } else {
        const __nv_bfloat16* src = tok_ptr(p.q, p.q_bs, p.q_tok, p.pad_q, qr);
        if (src) val = *reinterpret_cast<const uint4*>(src + c * 8);
      }
    }
    *reinterpret_cast<uint4*>(Qs + r * QS + c * 8) = val;
  }
  __syncthreads();

  // ---- Q fragments ----
  uint32_t qa[DP / 16][4];
  {
    const __nv_bfloat16* qrow0 = Qs + (warp * 16 + g) * QS;
    const __nv_bfloat16* qrow1 = qrow0 + 8 * QS;
#pragma unroll
    for (int kk = 0; kk < DP / 16; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(qrow0 + kk * 16 + 2 * t);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(qrow1 + kk * 16 + 2 * t);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(qrow0 + kk * 16 + 8 + 2 * t);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(qrow1 + kk * 16 + 8 + 2 * t);
    }
  }
  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const bool warp_active = (q0 + warp * 16) < Lq;

  // ---- K/V chunk producer (cp.async, zero-fill for rows >= Lk, padded dims and pad-less tokens) ----
  auto stage = [&](int k0, int buf) {
    __nv_bfloat16* Kb = KV + buf * (128 * QS);
    __nv_bfloat16* Vb = Kb + 64 * QS;
    for (int i = tid; i < 64 * (DP / 8); i += NT) {
      const int r = i / (DP / 8), c = i % (DP / 8);
      const int kr = k0 + r;
      const __nv_bfloat16* ks = nullptr;
      const __nv_bfloat16* vs = nullptr;
      if (kr < Lk && c < DV8) {
        ks = tok_ptr(p.k, p.k_bs, p.k_tok, p.pad_k, kr);
        vs = tok_ptr(p.v, p.v_bs, p.v_tok, p.pad_v, kr);
      }
      cp_async16(Kb + r * QS + c * 8, ks ? static_cast<const void*>(ks + c * 8) : static_cast<const void*>(p.k), ks ? 16 : 0);
      cp_async16(Vb + r * QS + c * 8, vs ? static_cast<const void*>(vs + c * 8) : static_cast<const void*>(p.v), vs ? 16 : 0);
    }
    cp_async_commit();
  };
  const int n_chunks = (Lk + 63) / 64;
  stage(0, 0);
  for (int j = 0; j < n_chunks; ++j) {
    const int k0 = j * 64;
    if (j + 1 < n_chunks) {
      stage(k0 + 64, (j + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* Ks = KV + (j & 1) * (128 * QS);
    const __nv_bfloat16* Vs = Ks + 64 * QS;
    if (warp_active) {
    // ---- S = Q K^T for this warp's 16 rows x 64 keys ----
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* krow = Ks + (nt * 8 + g) * QS;
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + kk * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + kk * 16 + 8 + 2 * t);
        mma_bf16_16816(s[nt], qa[kk], b0, b1);
      }
    }
    // ---- mask + online softmax (rows g and g+8) ----
    const int lim = Lk_valid - k0;  // keys >= lim masked
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      if (c0 >= lim) s[nt][0] = s[nt][2] = -INFINITY;
      if (c0 + 1 >= lim) s[nt][1] = s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = (m0 == -INFINITY) ? 0.f : ex2f_fast((m0 - mn0) * p.scale_log2);
    const float a1 = (m1 == -INFINITY) ? 0.f : ex2f_fast((m1 - mn1) * p.scale_log2);
    m0 = mn0;
    m1 = mn1;
    const float off0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2;
    const float off1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float e0 = ex2f_fast(fmaf(s[nt][0], p.scale_log2, -off0));
      const float e1 = ex2f_fast(fmaf(s[nt][1], p.scale_log2, -off0));
      const float e2 = ex2f_fast(fmaf(s[nt][2], p.scale_log2, -off1));
      const float e3 = ex2f_fast(fmaf(s[nt][3], p.scale_log2, -off1));
      sum0 += e0 + e1;
      sum1 += e2 + e3;
      __nv_bfloat162 p01 = __floats2bfloat162_rn(e0, e1);
      __nv_bfloat162 p23 = __floats2bfloat162_rn(e2, e3);
      pa[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&p23);
    }
    l0 = l0 * a0 + sum0;
    l1 = l1 * a1 + sum1;
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    // ---- O += P V : one ldmatrix.x4.trans yields the B fragments of two 8-wide output tiles ----
    {
      // lanes 0-7 / 8-15: keys kk*16 + 0..7 / 8..15 of dims tile i; lanes 16-31: same keys, tile i+1
      const uint32_t vbase = static_cast<uint32_t>(__cvta_generic_to_shared(
          Vs + ((lane & 7) + ((lane >> 3) & 1) * 8) * QS + (lane >> 4) * 8));
#pragma unroll
      for (int i = 0; i < DP / 8; i += 2) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t b0, b1, b2, b3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(vbase + static_cast<uint32_t>((kk * 16 * QS + i * 8) * 2)));
          mma_bf16_16816(o[i], pa[kk], b0, b1);
          mma_bf16_16816(o[i + 1], pa[kk], b2, b3);
        }
      }
    }
    }  // warp_active
    __syncthreads();  // every warp is done with buffer j&1 before chunk j+2 is staged into it
  }
  if (!warp_active) return;
  // ---- finalize: quad-reduce row sums, normalise, store ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  auto out_ptr = [&](int r) -> __nv_bfloat16* {
    if (r >= Lq) return nullptr;
    if (w == 0) return p.out + b * p.o_bs + static_cast<long long>(r) * p.o_tok + h * p.D;
    const int ww = p.q_pool ? w / 2 : w;
    const int Ho = p.q_pool ? p.Hm / 2 : p.Hm, Wo = p.q_pool ? p.Wm / 2 : p.Wm;
    const int gy = wy * ww + r / ww, gx = wx * ww + r % ww;
    if (gy >= Ho || gx >= Wo) return nullptr;
    return p.out + b * p.o_bs + (static_cast<long long>(gy) * Wo + gx) * p.o_tok + h * p.D;
  };
  __nv_bfloat16* o0 = out_ptr(q0 + warp * 16 + g);
  __nv_bfloat16* o1 = out_ptr(q0 + warp * 16 + g + 8);
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) {
    const int c = i * 8 + 2 * t;
    if (c < p.D) {
      if (o0) *reinterpret_cast<__nv_bfloat162*>(o0 + c) = __floats2bfloat162_rn(o[i][0] * i0, o[i][1] * i0);
      if (o1) *reinterpret_cast<__nv_bfloat162*>(o1 + c) = __floats2bfloat162_rn(o[i][2] * i1, o[i][3] * i1);
    }
  }
}

// ---- Hiera windows of 4x4 and 8x8 tokens (stages 1-2, hieradet.py:57-82, backbones/utils.py:16-63) -----------------
// 16 or 64 tokens per window: far too small for a 128-row tcgen05 tile, and on the generic kernel above a 4x4 window kept
// one warp of four busy on a quarter-filled key chunk (162 us per stage-2 block of a 4-frame pass, 7x the time its 150 MB
// of traffic takes).  Here a CTA owns WPC whole windows with ALL heads: a token's [q | k | v] row is contiguous in the
// fused qkv buffer, so the window is staged with 16-byte cp.async in four (eight) contiguous row segments, one warp runs
// one (window, head, 16-query tile) on mma.sync m16n8k16 with the whole key range in a single softmax pass, writes its
// O tile over its own Q tile in shared memory, and the CTA stores whole output rows.  With q_pool the 2x2 max-pooled
// queries (hieradet.py:65-68) are built in shared memory first.
template <int W, int DP, int WPC>
__global__ void __launch_bounds__(512) win_small_attn_kernel(const MhaParams p, int C, int tasks_per_window) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  constexpr int NTOK = W * W;
  constexpr int NT = NTOK / 8;         // key n-tiles of S
  constexpr int KP = NTOK / 16;        // k-steps of P.V
  extern __shared__ __align__(16) uint8_t smem_ws[];
  const int RS = 3 * C + 8;            // elements per staged token row (16-byte pad: conflict-free fragment reads)
  const int QPS = C + 8;               // row stride of the pooled-query tile
  __nv_bfloat16* tok = reinterpret_cast<__nv_bfloat16*>(smem_ws);              // [WPC][NTOK][RS]
  __nv_bfloat16* qp = tok + WPC * NTOK * RS;                                     // [WPC][16 * ceil(NTOK/64)][QPS] (q_pool only)
  constexpr int NQP = NTOK / 4;        // pooled queries per window
  constexpr int QPR = (NQP + 15) / 16 * 16;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int per = p.nwy * p.nwx;
  const int win0 = blockIdx.x * WPC;   // first window of this CTA (windows of one CTA are neighbours in x)
  const int b = win0 / per;
  const int wy = (win0 % per) / p.nwx, wx0 = (win0 % per) % p.nwx;
  const int D = p.D;
  const int pieces = 3 * C / 8;
  // ---- stage the windows ----
  for (int i = tid; i < WPC * NTOK * pieces; i += nthr) {
    const int c = i % pieces, r = (i / pieces) % NTOK, wi = i / (pieces * NTOK);
    const int gy = wy * W + r / W, gx = (wx0 + wi) * W + r % W;
    const __nv_bfloat16* src = p.q + b * p.q_bs + (static_cast<long long>(gy) * p.Wm + gx) * p.q_tok + c * 8;
    cp_async16(tok + (wi * NTOK + r) * RS + c * 8, src, 16);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (p.q_pool) {
    for (int i = tid; i < WPC * QPR * (C / 8); i += nthr) {
      const int c = i % (C / 8), r = (i / (C / 8)) % QPR, wi = i / ((C / 8) * QPR);
      uint4 val = make_uint4(0, 0, 0, 0);
      if (r < NQP) {
        const int py = r / (W / 2), px = r % (W / 2);
        const __nv_bfloat16* t00 = tok + (wi * NTOK + (2 * py) * W + 2 * px) * RS + c * 8;
        val = bf16x8_max(bf16x8_max(*reinterpret_cast<const uint4*>(t00), *reinterpret_cast<const uint4*>(t00 + RS)),
                         bf16x8_max(*reinterpret_cast<const uint4*>(t00 + W * RS), *reinterpret_cast<const uint4*>(t00 + (W + 1) * RS)));
      }
      *reinterpret_cast<uint4*>(qp + (wi * QPR + r) * QPS + c * 8) = val;
    }
    __syncthreads();
  }
  // ---- one warp = one (window, head, 16-query tile) ----
  const int wi = warp / tasks_per_window;
  const int task = warp % tasks_per_window;
  const int qtiles = p.q_pool ? QPR / 16 : NTOK / 16;
  const int h = task / qtiles, qt = task % qtiles;
  if (wi < WPC && h < p.H) {
    const __nv_bfloat16* wtok = tok + wi * NTOK * RS;
    __nv_bfloat16* qbase = p.q_pool ? qp + (wi * QPR + qt * 16) * QPS + h * D : tok + (wi * NTOK + qt * 16) * RS + h * D;
    const int qs = p.q_pool ? QPS : RS;
    uint32_t qa[DP / 16][4];
    {
      const __nv_bfloat16* qrow0 = qbase + g * qs;
      const __nv_bfloat16* qrow1 = qrow0 + 8 * qs;
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        qa[kk][0] = *reinterpret_cast<const uint32_t*>(qrow0 + kk * 16 + 2 * t);
        qa[kk][1] = *reinterpret_cast<const uint32_t*>(qrow1 + kk * 16 + 2 * t);
        if (D < DP && kk == DP / 16 - 1) {
          // DP - D == 8: columns >= D belong to the next head — whose warp may already be writing its O over its own
          // query columns (compute-sanitizer racecheck flagged the discarded read); they are never read
          qa[kk][2] = qa[kk][3] = 0u;
        } else {
          qa[kk][2] = *reinterpret_cast<const uint32_t*>(qrow0 + kk * 16 + 8 + 2 * t);
          qa[kk][3] = *reinterpret_cast<const uint32_t*>(qrow1 + kk * 16 + 8 + 2 * t);
        }
      }
    }
    float s[NT][4];
    const __nv_bfloat16* kb = wtok + C + h * D;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* krow = kb + (nt * 8 + g) * RS;
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + kk * 16 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + kk * 16 + 8 + 2 * t);
        mma_bf16_16816(s[nt], qa[kk], b0, b1);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float off0 = mx0 * p.scale_log2, off1 = mx1 * p.scale_log2;
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[KP][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float e0 = ex2f_fast(fmaf(s[nt][0], p.scale_log2, -off0));
      const float e1 = ex2f_fast(fmaf(s[nt][1], p.scale_log2, -off0));
      const float e2 = ex2f_fast(fmaf(s[nt][2], p.scale_log2, -off1));
      const float e3 = ex2f_fast(fmaf(s[nt][3], p.scale_log2, -off1));
      l0 += e0 + e1;
      l1 += e2 + e3;
      __nv_bfloat162 p01 = __floats2bfloat162_rn(e0, e1);
      __nv_bfloat162 p23 = __floats2bfloat162_rn(e2, e3);
      pa[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&p23);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // ---- O = P V; the odd last 8-column tile is computed as half of a pair (its twin reads the 16-byte row pad or the
    //      next head's columns and is dropped) ----
    float o[DP / 8][4];
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    const uint32_t vbase = static_cast<uint32_t>(__cvta_generic_to_shared(
        wtok + 2 * C + h * D + ((lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 8));
#pragma unroll
    for (int i = 0; i < DP / 8; i += 2) {
#pragma unroll
      for (int kk = 0; kk < KP; ++kk) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vbase + static_cast<uint32_t>((kk * 16 * RS + i * 8) * 2)));
        mma_bf16_16816(o[i], pa[kk], b0, b1);
        mma_bf16_16816(o[i + 1], pa[kk], b2, b3);
      }
    }
    __syncwarp();   // every lane has read its Q fragments before the tile is overwritten with O
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      const int c = i * 8 + 2 * t;
      if (c < D) {
        *reinterpret_cast<__nv_bfloat162*>(qbase + g * qs + c) = __floats2bfloat162_rn(o[i][0] * i0, o[i][1] * i0);
        *reinterpret_cast<__nv_bfloat162*>(qbase + (g + 8) * qs + c) = __floats2bfloat162_rn(o[i][2] * i1, o[i][3] * i1);
      }
    }
  }
  __syncthreads();
  // ---- store whole output rows (C bf16 per token) ----
  const int nq = p.q_pool ? NQP : NTOK;
  const int ww = p.q_pool ? W / 2 : W;
  const int Wo = p.q_pool ? p.Wm / 2 : p.Wm;
  for (int i = tid; i < WPC * nq * (C / 8); i += nthr) {
    const int c = i % (C / 8), r = (i / (C / 8)) % nq, wi2 = i / ((C / 8) * nq);
    const __nv_bfloat16* src = p.q_pool ? qp + (wi2 * QPR + r) * QPS + c * 8 : tok + (wi2 * NTOK + r) * RS + c * 8;
    const int gy = wy * ww + r / ww, gx = (wx0 + wi2) * ww + r % ww;
    *reinterpret_cast<uint4*>(p.out + b * p.o_bs + (static_cast<long long>(gy) * Wo + gx) * p.o_tok + c * 8) =
        *reinterpret_cast<const uint4*>(src);
  }
}

template <int W, int DP, int WPC>
static int launch_win_small(const MhaParams& p, int C, cudaStream_t st) {
  constexpr int NTOK = W * W;
  const int qtiles = p.q_pool ? (NTOK / 4 + 15) / 16 : NTOK / 16;
  const int tpw = p.H * qtiles;
  // at least 8 warps: with few (pooled) tasks the extra warps only help staging and storing the window
  const int threads = 32 * WPC * tpw < 256 ? 256 : 32 * WPC * tpw;
  const int smem = (WPC * NTOK * (3 * C + 8) + WPC * ((NTOK / 4 + 15) / 16 * 16) * (C + 8)) * 2;
  if (threads > 512 || smem > 200 * 1024) return -1;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(win_small_attn_kernel<W, DP, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_mha: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_smem = smem;
  }
  const int windows = p.B * p.nwy * p.nwx;
  DS2_LAUNCH((win_small_attn_kernel<W, DP, WPC>), windows / WPC, threads, smem, st, p, C, tpw);
  return post_launch("win_small_attn_kernel");
}

// 4x4 / 8x8 windows of a fused [q | k | v] buffer without zero-padded windows; -1 = shape not covered (generic kernel)
static int try_win_small(const ds2_mha_args* a, const MhaParams& p, cudaStream_t st) {
  const int C = a->H * a->D;
  if (a->window != 4 && a->window != 8) return -1;
  if ((a->Hm % a->window) || (a->Wm % a->window) || (C % 8)) return -1;
  const __nv_bfloat16 *q = p.q, *k = p.k, *v = p.v;
  if (k != q + C || v != q + 2 * C || a->q_tok_stride != 3 * C || a->k_tok_stride != 3 * C || a->v_tok_stride != 3 * C) return -1;
  if (a->k_bs != a->q_bs || a->v_bs != a->q_bs || (a->o_tok_stride % 8) || (a->q_bs % 8) || (a->o_bs % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) || (reinterpret_cast<uintptr_t>(a->out) & 15)) return -1;
  const int DP = (a->D + 15) / 16 * 16;
  if (DP - a->D != 0 && DP - a->D != 8) return -1;
  if (a->window == 4) {
    if (p.nwx % 2) return -1;
    if (DP == 64) return launch_win_small<4, 64, 2>(p, C, st);
    if (DP == 80) return launch_win_small<4, 80, 2>(p, C, st);
    if (DP == 96) return launch_win_small<4, 96, 2>(p, C, st);
    return -1;
  }
  if (DP == 64) return launch_win_small<8, 64, 1>(p, C, st);
  if (DP == 80) return launch_win_small<8, 80, 1>(p, C, st);
  if (DP == 96) return launch_win_small<8, 96, 1>(p, C, st);
  return -1;
}

// ---- the mask decoder's two attention shapes (transformer.py:178-211: 8 heads x 16 dims) ---------------------------
// (1) token -> image: 8-9 queries against 4096 keys per (object, head).  On the generic kernel one warp of one CTA walked
//     the 64 key chunks in sequence (33 us, latency).  Here the chunks are dealt round-robin to the 4 warps of the 8 CTAs
//     of a thread-block cluster; every warp runs the same mma.sync online softmax as above on its chunks, the 32 partial
//     (max, sum, O) triples are merged by the cluster's rank-0 CTA through distributed shared memory in a fixed order
//     (no workspace, no atomics, bit-reproducible).
constexpr int kFqSplit = 8;            // CTAs per cluster = key slices
constexpr int kFqRow = 18;             // floats per partial row: 16 outputs + running max + running sum

__global__ void __cluster_dims__(1, 1, kFqSplit) __launch_bounds__(128) fewq_attn_kernel(const MhaParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  constexpr int DP = 16, QS = DP + 8;
  __shared__ __align__(16) __nv_bfloat16 Qs[16 * QS];
  __shared__ __align__(16) __nv_bfloat16 KVs[4][2][64 * QS];   // per warp: K chunk, V chunk
  __shared__ float red[4][16][kFqRow];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x, h = blockIdx.y;
  const int Lq = p.Lq, Lk = p.Lk;
  const int Lk_valid = p.Lk_valid > 0 ? p.Lk_valid : Lk;
  for (int i = tid; i < 16 * 2; i += 128) {
    const int r = i >> 1, c = i & 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < Lq) val = *reinterpret_cast<const uint4*>(p.q + b * p.q_bs + static_cast<long long>(r) * p.q_tok + h * DP + c * 8);
    *reinterpret_cast<uint4*>(Qs + r * QS + c * 8) = val;
  }
  __syncthreads();
  uint32_t qa[4];
  {
    const __nv_bfloat16* qrow0 = Qs + g * QS;
    const __nv_bfloat16* qrow1 = qrow0 + 8 * QS;
    qa[0] = *reinterpret_cast<const uint32_t*>(qrow0 + 2 * t);
    qa[1] = *reinterpret_cast<const uint32_t*>(qrow1 + 2 * t);
    qa[2] = *reinterpret_cast<const uint32_t*>(qrow0 + 8 + 2 * t);
    qa[3] = *reinterpret_cast<const uint32_t*>(qrow1 + 8 + 2 * t);
  }
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  __nv_bfloat16* Ks = KVs[warp][0];
  __nv_bfloat16* Vs = KVs[warp][1];
  const int n_chunks = (Lk + 63) / 64;
  for (int j = blockIdx.z * 4 + warp; j < n_chunks; j += 4 * kFqSplit) {
    const int k0 = j * 64;
    __syncwarp();   // the previous chunk's ldmatrix reads are done before the buffers are overwritten
    for (int i = lane; i < 64 * 2; i += 32) {
      const int r = i >> 1, c = i & 1;
      const int kr = k0 + r;
      const bool ok = kr < Lk;
      const long long tok = ok ? kr : 0;
      cp_async16(Ks + r * QS + c * 8, p.k + b * p.k_bs + tok * p.k_tok + h * DP + c * 8, ok ? 16 : 0);
      cp_async16(Vs + r * QS + c * 8, p.v + b * p.v_bs + tok * p.v_tok + h * DP + c * 8, ok ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* krow = Ks + (nt * 8 + g) * QS;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + 2 * t);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + 8 + 2 * t);
      mma_bf16_16816(s[nt], qa, b0, b1);
    }
    const int lim = Lk_valid - k0;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      if (c0 >= lim) s[nt][0] = s[nt][2] = -INFINITY;
      if (c0 + 1 >= lim) s[nt][1] = s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = (m0 == -INFINITY) ? 0.f : ex2f_fast((m0 - mn0) * p.scale_log2);
    const float a1 = (m1 == -INFINITY) ? 0.f : ex2f_fast((m1 - mn1) * p.scale_log2);
    m0 = mn0;
    m1 = mn1;
    const float off0 = (mn0 == -INFINITY) ? 0.f : mn0 * p.scale_log2;
    const float off1 = (mn1 == -INFINITY) ? 0.f : mn1 * p.scale_log2;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float e0 = ex2f_fast(fmaf(s[nt][0], p.scale_log2, -off0));
      const float e1 = ex2f_fast(fmaf(s[nt][1], p.scale_log2, -off0));
      const float e2 = ex2f_fast(fmaf(s[nt][2], p.scale_log2, -off1));
      const float e3 = ex2f_fast(fmaf(s[nt][3], p.scale_log2, -off1));
      sum0 += e0 + e1;
      sum1 += e2 + e3;
      __nv_bfloat162 p01 = __floats2bfloat162_rn(e0, e1);
      __nv_bfloat162 p23 = __floats2bfloat162_rn(e2, e3);
      pa[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&p23);
    }
    l0 = l0 * a0 + sum0;
    l1 = l1 * a1 + sum1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    const uint32_t vbase = static_cast<uint32_t>(__cvta_generic_to_shared(
        Vs + ((lane & 7) + ((lane >> 3) & 1) * 8) * QS + (lane >> 4) * 8));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t b0, b1, b2, b3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                   : "r"(vbase + static_cast<uint32_t>((kk * 16 * QS) * 2)));
      mma_bf16_16816(o[0], pa[kk], b0, b1);
      mma_bf16_16816(o[1], pa[kk], b2, b3);
    }
  }
  // ---- this warp's partial -> shared memory ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    red[warp][g][i * 8 + 2 * t] = o[i][0];
    red[warp][g][i * 8 + 2 * t + 1] = o[i][1];
    red[warp][g + 8][i * 8 + 2 * t] = o[i][2];
    red[warp][g + 8][i * 8 + 2 * t + 1] = o[i][3];
  }
  if (t == 0) {
    red[warp][g][16] = m0;
    red[warp][g][17] = l0;
    red[warp][g + 8][16] = m1;
    red[warp][g + 8][17] = l1;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  if (cluster.block_rank() == 0) {
    // 256 outputs, 2 per thread; partials are merged in (rank, warp) order
    for (int e = tid; e < 16 * 16; e += 128) {
      const int r = e >> 4, c = e & 15;
      if (r >= Lq) continue;
      float M = -INFINITY;
      for (int rk = 0; rk < kFqSplit; ++rk) {
        const float* rr = cluster.map_shared_rank(&red[0][0][0], rk);
#pragma unroll
        for (int w = 0; w < 4; ++w) M = fmaxf(M, rr[(w * 16 + r) * kFqRow + 16]);
      }
      float L = 0.f, O = 0.f;
      for (int rk = 0; rk < kFqSplit; ++rk) {
        const float* rr = cluster.map_shared_rank(&red[0][0][0], rk);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float* row = rr + (w * 16 + r) * kFqRow;
          const float mw = row[16];
          const float sc = (mw == -INFINITY) ? 0.f : ex2f_fast((mw - M) * p.scale_log2);
          L = fmaf(row[17], sc, L);
          O = fmaf(row[c], sc, O);
        }
      }
      p.out[b * p.o_bs + static_cast<long long>(r) * p.o_tok + h * DP + c] = __float2bfloat16_rn(O / L);
    }
  }
  cluster.sync();   // the other CTAs' shared memory stays alive until rank 0 has read it
}

// (2) image -> token: 4096 queries against the 8-9 token keys per (object, head).  One thread = one (query, head):
//     K / V of the object live in shared memory as f32 (head rows padded 16 -> 20 words: conflict-free float4 reads), the
//     scores, the softmax and P.V stay in f32 registers; a warp reads / writes 4 whole query rows (256 B each).
constexpr int kFkMaxKeys = 16;
__global__ void __launch_bounds__(256) fewk_attn_kernel(const MhaParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  constexpr int DP = 16, HS = 20;
  __shared__ __align__(16) float Ks[kFkMaxKeys][8 * HS];
  __shared__ __align__(16) float Vs[kFkMaxKeys][8 * HS];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int Lk = p.Lk;
  const int Lk_valid = p.Lk_valid > 0 ? min(p.Lk_valid, Lk) : Lk;
  for (int i = tid; i < Lk * 128; i += 256) {
    const int j = i >> 7, c = i & 127;
    Ks[j][(c >> 4) * HS + (c & 15)] = __bfloat162float(p.k[b * p.k_bs + static_cast<long long>(j) * p.k_tok + c]);
    Vs[j][(c >> 4) * HS + (c & 15)] = __bfloat162float(p.v[b * p.v_bs + static_cast<long long>(j) * p.v_tok + c]);
  }
  __syncthreads();
  const int h = tid & 7;
  const int qi = blockIdx.x * 32 + (tid >> 3);
  if (qi >= p.Lq) return;
  float q[DP];
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.q + b * p.q_bs + static_cast<long long>(qi) * p.q_tok + h * DP);
    const uint4 u0 = __ldg(src), u1 = __ldg(src + 1);
    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      q[2 * i] = __uint_as_float(w[i] << 16);
      q[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  float s[kFkMaxKeys];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kFkMaxKeys; ++j) {
    s[j] = -INFINITY;
    if (j < Lk_valid) {
      const float4* kr = reinterpret_cast<const float4*>(&Ks[j][h * HS]);
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 kv = kr[i];
        a = fmaf(q[4 * i], kv.x, a);
        a = fmaf(q[4 * i + 1], kv.y, a);
        a = fmaf(q[4 * i + 2], kv.z, a);
        a = fmaf(q[4 * i + 3], kv.w, a);
      }
      s[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  float acc[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) acc[i] = 0.f;
  float l = 0.f;
  const float off = mx * p.scale_log2;
#pragma unroll
  for (int j = 0; j < kFkMaxKeys; ++j) {
    if (j < Lk_valid) {
      const float e = ex2f_fast(fmaf(s[j], p.scale_log2, -off));
      l += e;
      const float4* vr = reinterpret_cast<const float4*>(&Vs[j][h * HS]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 vv = vr[i];
        acc[4 * i] = fmaf(e, vv.x, acc[4 * i]);
        acc[4 * i + 1] = fmaf(e, vv.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(e, vv.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(e, vv.w, acc[4 * i + 3]);
      }
    }
  }
  const float inv = 1.f / l;
  uint32_t ow[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 v2 = __floats2bfloat162_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
    ow[i] = *reinterpret_cast<uint32_t*>(&v2);
  }
  uint4* dst = reinterpret_cast<uint4*>(p.out + b * p.o_bs + static_cast<long long>(qi) * p.o_tok + h * DP);
  dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
}

template <int DP, int NW>
static int launch_mha_nw(const MhaParams& p, dim3 grid, cudaStream_t st) {
  const int smem = (16 * NW + 2 * 128) * (DP + 8) * 2;
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) {
      cudaFuncSetAttribute(mha_kernel<DP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      set = true;
    }
  }
  DS2_LAUNCH((mha_kernel<DP, NW>), grid, 32 * NW, smem, st, p);
  return post_launch("mha_kernel");
}

// 128-row query tiles when a sequence has at least 128 queries (halves the K/V re-reads of the global
// Hiera blocks and the image->token decoder attention), 64-row tiles otherwise.
template <int DP>
static int launch_mha(const MhaParams& p, int nseq, int H, int Lq, cudaStream_t st) {
  if (Lq >= 128) return launch_mha_nw<DP, 8>(p, dim3(nseq, H, (Lq + 127) / 128), st);
  return launch_mha_nw<DP, 4>(p, dim3(nseq, H, (Lq + 63) / 64), st);
}

int launch_win16_attn_tc(const ds2_mha_args* a, cudaStream_t st);  // win_attn_tc.cu (tcgen05)
int launch_glob_attn_tc(const ds2_mha_args* a, cudaStream_t st);   // win_attn_tc.cu (tcgen05)
int launch_glob_flash(const ds2_mha_args* a, cudaStream_t st, int sp);  // flash_tc.cu (tcgen05, TMA ring, Q in TMEM)

}  // namespace ds2

extern "C" int ds2_mha(const ds2_mha_args* a, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a != nullptr, DS2_E_ARG, "ds2_mha: null args");
  DS2_REQUIRE(a->q && a->k && a->v && a->out, DS2_E_ARG, "ds2_mha: null pointer");
  DS2_REQUIRE(a->B > 0 && a->H > 0 && a->D > 0 && a->D <= 128 && (a->D % 8) == 0, DS2_E_ARG,
              "ds2_mha: bad head spec B=%d H=%d D=%d", a->B, a->H, a->D);
  DS2_REQUIRE((a->q_tok_stride % 8) == 0 && (a->k_tok_stride % 8) == 0 && (a->v_tok_stride % 8) == 0 &&
                  (a->o_tok_stride % 2) == 0,
              DS2_E_ALIGN, "ds2_mha: token strides must be multiples of 8 elements");
  MhaParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.q_tok = a->q_tok_stride;
  p.k_tok = a->k_tok_stride;
  p.v_tok = a->v_tok_stride;
  p.o_tok = a->o_tok_stride;
  p.q_bs = a->q_bs;
  p.k_bs = a->k_bs;
  p.v_bs = a->v_bs;
  p.o_bs = a->o_bs;
  p.B = a->B;
  p.H = a->H;
  p.D = a->D;
  p.Lq = a->Lq;
  p.Lk = a->Lk;
  p.window = a->window;
  p.Hm = a->Hm;
  p.Wm = a->Wm;
  p.q_pool = a->q_pool;
  p.Lk_valid = a->Lk_valid;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.pad_q = reinterpret_cast<const __nv_bfloat16*>(a->pad_q);
  p.pad_k = reinterpret_cast<const __nv_bfloat16*>(a->pad_k);
  p.pad_v = reinterpret_cast<const __nv_bfloat16*>(a->pad_v);
  p.nwy = p.nwx = 1;
  int Lq, nseq;
  if (a->window > 0) {
    DS2_REQUIRE(a->Hm > 0 && a->Wm > 0, DS2_E_ARG, "ds2_mha: window mode needs Hm, Wm");
    DS2_REQUIRE(!a->q_pool || ((a->window % 2) == 0 && (a->Hm % 2) == 0 && (a->Wm % 2) == 0), DS2_E_ARG,
                "ds2_mha: q_pool needs even window and map");
    p.nwy = (a->Hm + a->window - 1) / a->window;
    p.nwx = (a->Wm + a->window - 1) / a->window;
    const bool padded = (a->Hm % a->window) != 0 || (a->Wm % a->window) != 0;
    DS2_REQUIRE(!padded || (a->pad_q && a->pad_k && a->pad_v), DS2_E_ARG,
                "ds2_mha: padded windows need pad_q/pad_k/pad_v");
    Lq = a->q_pool ? (a->window / 2) * (a->window / 2) : a->window * a->window;
    nseq = a->B * p.nwy * p.nwx;
  } else {
    DS2_REQUIRE(a->Lq > 0 && a->Lk > 0, DS2_E_ARG, "ds2_mha: bad Lq/Lk");
    Lq = a->Lq;
    nseq = a->B;
  }
  DS2_REQUIRE((Lq + 63) / 64 <= 65535 && a->H <= 65535, DS2_E_ARG, "ds2_mha: grid too large");
  cudaStream_t st = as_stream(stream);
  {
    // Hiera stage-3 windows (16x16 tokens, head_dim 72) and global blocks run on the tensor-memory kernels; DS2_WIN_TC=0 keeps
    // them on the generic mma.sync kernel (A/B testing)
    static const bool win_tc = [] {
      const char* e = getenv("DS2_WIN_TC");
      return !(e && e[0] == '0');
    }();
    if (win_tc) {
      int rc = -1;
      // DS2_WIN_FLASH=1: 16 x 16 windows through the flash kernel's multi-head variant (two 128-row items per window)
      const char* wf = getenv("DS2_WIN_FLASH");
      if (wf && atoi(wf) > 0 && a->window == 16) {
        rc = launch_glob_flash(a, st, atoi(wf));
        if (rc >= 0) return rc;
      }
      rc = launch_win16_attn_tc(a, st);
      if (rc >= 0) return rc;
      // Global blocks: the flash kernel's multi-head variant (TMA ring, Q in TMEM, S double-buffered): 281 us against 496 us
      // for the 4-frame launch of the large model.  Measured on the same box (profiles/r2_s17_glob_flash_ab.txt): one or two
      // softmax threads per row 281 / 281 us, two alternating softmax groups 269 us, half of the TMA boxes 278 us — the
      // softmax warps are busy ~1 900 of the ~2 100 clk per key tile however the row is split (16 384 exponentials per tile
      // = 1 024 clk of MUFU per SM, plus the TMEM round trips).  DS2_GLOB_FLASH=0: the serial-chain kernel of
      // win_attn_tc.cu (A/B), 2: two softmax threads per row.  (read per call: the tests switch it)
      const char* gf = getenv("DS2_GLOB_FLASH");
      const int glob_flash = gf ? atoi(gf) : 1;
      if (glob_flash > 0 && a->window == 0) {
        rc = launch_glob_flash(a, st, glob_flash);
        if (rc >= 0) return rc;
      }
      rc = launch_glob_attn_tc(a, st);
      if (rc >= 0) return rc;
    }
  }
  if (a->window > 0) {
    static const bool win_small = [] {
      const char* e = getenv("DS2_WIN_SMALL");   // DS2_WIN_SMALL=0: generic kernel (A/B)
      return !(e && e[0] == '0');
    }();
    if (win_small) {
      const int rc = try_win_small(a, p, st);
      if (rc >= 0) return rc;
    }
  }
  const int D = a->D;
  if (a->window == 0 && D == 16 && a->H == 8 && (a->o_tok_stride % 8) == 0) {   // the mask decoder's two shapes
    static const bool dec_fast = [] {
      const char* e = getenv("DS2_DEC_ATTN");   // DS2_DEC_ATTN=0: generic mma.sync kernel (A/B)
      return !(e && e[0] == '0');
    }();
    if (dec_fast && Lq <= 16 && a->Lk >= 256 && a->B <= 65535) {
      DS2_LAUNCH((fewq_attn_kernel), dim3(a->B, a->H, kFqSplit), 128, 0, st, p);
      return post_launch("fewq_attn_kernel");
    }
    if (dec_fast && a->Lk <= kFkMaxKeys && Lq >= 256 && a->B <= 65535) {
      DS2_LAUNCH((fewk_attn_kernel), dim3((Lq + 31) / 32, a->B), 256, 0, st, p);
      return post_launch("fewk_attn_kernel");
    }
  }
  if (D <= 16) return launch_mha<16>(p, nseq, a->H, Lq, st);
  if (D <= 32) return launch_mha<32>(p, nseq, a->H, Lq, st);
  if (D <= 64) return launch_mha<64>(p, nseq, a->H, Lq, st);
  if (D <= 80) return launch_mha<80>(p, nseq, a->H, Lq, st);
  if (D <= 96) return launch_mha<96>(p, nseq, a->H, Lq, st);
  return launch_mha<128>(p, nseq, a->H, Lq, st);
}
