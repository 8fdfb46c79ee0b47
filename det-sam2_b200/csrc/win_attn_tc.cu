// win_attn_tc.cu — Hiera windowed attention (window 16x16 = 256 tokens, head_dim <= 80) on tcgen05.
//
// One CTA = one (window, head): S = Q K^T and O = P V for 256 queries x 256 keys, head_dim 72 zero-padded
// to 80 in shared memory only (the qkv activations keep the reference layout, hieradet.py:57-82).
//   all warps   : stage Q, K, V of the window from the token-major qkv matrix into shared memory with
//                 16-byte loads, directly in the tensor core's canonical layouts:
//                   Q, K : K-major, SWIZZLE_128B, [rows][64 dims] + a second 64-wide block for dims 64..79
//                   V    : MN-major, SWIZZLE_128B, [keys][64 dims] + a second block for dims 64..79
//                 (a window is a 16x16 patch of the token raster: rows are 144-byte segments 3*dim*2
//                 bytes apart, which no TMA box maps onto a 128-byte-swizzled tile).
//   warp 8      : TMEM allocation (512 columns) and MMA issue (elect.sync):
//                   S_h = Q_h K^T   tcgen05.mma SS  M=128 N=256, 5 K-steps      -> TMEM cols [256h, 256h+256)
//                   O_h = P_h V     tcgen05.mma TS  M=128 N=80, 16 K-steps      -> TMEM cols [256h+128, +80)
//                 for the two 128-row query tiles h = 0, 1 (Q K^T of tile 1 overlaps the softmax of tile 0).
//   warps 0..7  : softmax, one thread per query row (TMEM lane): pass 1 row max over the 256 scores,
//                 pass 2 ex2 / row sum / bf16 P written over the first 128 columns of S_h (always behind
//                 the read pointer); O_h re-uses the dead upper half of S_h.  Epilogue O / l -> bf16.
// The whole window fits one pass, so there is no online rescaling.
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "tc05.cuh"

namespace ds2 {

struct WinAttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* out;
  long long q_tok, k_tok, v_tok, o_tok;  // elements between tokens
  long long q_bs, k_bs, v_bs, o_bs;      // elements between batch items
  int nwin;                              // windows per batch item
  int H, D;                              // heads, head_dim (multiple of 8, 64 < D <= 80)
  int Wm;                                // raster width in tokens
  int nwx;                               // windows per raster row
  float scale_log2;
  int dbg;
};

constexpr int kWinTok = 256;                   // tokens per window
constexpr int kWinSide = 16;
constexpr int kBlkBytes = kWinTok * 128;       // one 64-dim block of a 256-row tile (32 KB)
constexpr int kWinSmem = 6 * kBlkBytes + 1024 + 128;
constexpr int kWinThreads = 288;

// Phase timestamps of CTA 0 (tuning aid, DS2_WIN_DBG=1): [0..8] softmax warp 0, [9..13] MMA thread.
__device__ long long g_win_times[16];
#define WIN_T(i)                                                   \
  do {                                                             \
    if (p.dbg && blockIdx.x == 0) g_win_times[i] = clock64() - t_start; \
  } while (0)

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kWinThreads, 1) win16_attn_tc_kernel(const WinAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  // [Q blk0 | Q blk1 | K blk0 | K blk1 | V blk0 | V blk1], 32 KB each
  const uint32_t sq = base, sk = base + 2 * kBlkBytes, sv = base + 4 * kBlkBytes;
  const uint32_t bar_base = base + 6 * kBlkBytes;
  auto bar_s = [&](int h) { return bar_base + 8u * h; };        // S_h complete (tcgen05.commit)
  auto bar_p = [&](int h) { return bar_base + 16u + 8u * h; };  // P_h written (4 softmax warps)
  auto bar_o = [&](int h) { return bar_base + 32u + 8u * h; };  // O_h complete (tcgen05.commit)
  const uint32_t tmem_slot = bar_base + 64u;

  const long long t_start = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.H;
  const int win = blockIdx.x / p.H;
  const int b = win / p.nwin;
  const int wy = (win % p.nwin) / p.nwx, wx = win % p.nwx;

  if (threadIdx.x == 0) {
    for (int h = 0; h < 2; ++h) {
      tc::mbar_init(bar_s(h), 1);
      tc::mbar_init(bar_p(h), 4);
      tc::mbar_init(bar_o(h), 1);
    }
    tc::mbar_init(bar_base + 56u, kWinThreads);
    tc::fence_barrier_init();
  }
  if (warp == 8) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  pdl_sync();  // everything above is independent of the producer of qkv
  if (threadIdx.x == 0) WIN_T(1);

  // ---- stage the window: 3 matrices x 256 tokens x 10 chunks of 8 dims (chunks >= D/8 are zero padding) ----
  // Thread -> (chunk c, row r0) once, then r advances by a fixed step: no divisions in the loops.  All loads
  // of a thread are issued before its first store so that their L2 latencies overlap.
  auto bar_v = bar_base + 56u;  // V staged (all threads arrive)
  {
    constexpr int kRowsPerIter = 28;            // 280 of the 288 threads = 28 rows x 10 chunks
    constexpr int kIters = (kWinTok + kRowsPerIter - 1) / kRowsPerIter;  // 10
    const int nchunk = p.D >> 3;                // 9 for head_dim 72
    const int c = threadIdx.x % 10;
    const int r0 = threadIdx.x / 10;
    const bool active = r0 < kRowsPerIter;
    const long long tok0 = static_cast<long long>(wy * kWinSide) * p.Wm + wx * kWinSide;
    const __nv_bfloat16* srcs[3] = {p.q + b * p.q_bs + head * p.D + c * 8, p.k + b * p.k_bs + head * p.D + c * 8,
                                    p.v + b * p.v_bs + head * p.D + c * 8};
    const long long strides[3] = {p.q_tok, p.k_tok, p.v_tok};
    const uint32_t bases[3] = {sq, sk, sv};
    // 128-byte-swizzled tile: row r at (r/8)*1024 + (r%8)*128, 16-byte chunk j at j ^ (r%8); dims 64.. in block 1
    auto put = [&](int m, int r, const uint4& x) {
      const uint32_t dst = bases[m] + (c >> 3) * kBlkBytes + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
    };
    uint4 val[2][kIters];
    auto load = [&](int m, uint4(&dst)[kIters]) {
#pragma unroll
      for (int t = 0; t < kIters; ++t) {
        const int r = r0 + t * kRowsPerIter;
        dst[t] = make_uint4(0u, 0u, 0u, 0u);
        if (active && r < kWinTok && c < nchunk) {
          const long long tok = tok0 + static_cast<long long>(r >> 4) * p.Wm + (r & 15);
          dst[t] = __ldg(reinterpret_cast<const uint4*>(srcs[m] + tok * strides[m]));
        }
      }
    };
    auto store = [&](int m, const uint4(&src)[kIters]) {
#pragma unroll
      for (int t = 0; t < kIters; ++t) {
        const int r = r0 + t * kRowsPerIter;
        if (active && r < kWinTok) put(m, r, src[t]);
      }
    };
    load(0, val[0]);
    load(1, val[1]);
    if (threadIdx.x == 0) {
      if (val[1][kIters - 1].x == 0x12345678u && p.dbg == 77) g_win_times[15] = 1;  // timestamp after the loads return
      WIN_T(2);
    }
    store(0, val[0]);
    load(2, val[0]);  // V loads in flight while K is stored and Q K^T starts
    store(1, val[1]);
    // Q and K are in place: the MMA warp may start Q K^T while V is still being written
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (threadIdx.x == 0) WIN_T(3);
    // Q K^T goes to the tensor pipe BEFORE the MMA warp stores its share of V (it used to sit behind those stores and the
    // V loads they wait for: ~1 500 clk of the CTA's ~16 000, DS2_WIN_DBG=1)
    uint32_t tb;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tb) : "r"(tmem_slot));
    if (warp == 8 && tc::elect_one()) {
      const uint32_t idesc_qk = tc::make_idesc_bf16(128, 256, 0, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t d = tb + 256u * h;
        const uint32_t sqh = sq + h * (128 * 128);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
          // K-steps 0..3: dims 16ks..16ks+15 of block 0; K-step 4: dims 64..79 = first 32 bytes of block 1
          const uint32_t off = ks < 4 ? ks * 32 : kBlkBytes;
          const uint64_t da = tc::make_desc_sw128(sqh + off, 16, 1024);
          const uint64_t db = tc::make_desc_sw128(sk + off, 16, 1024);
          tc::umma_ss(d, da, db, idesc_qk, ks != 0 ? 1u : 0u);
        }
        tc::umma_commit(bar_s(h));
      }
      WIN_T(9);
    }
    __syncwarp();
    store(2, val[0]);
    tc::fence_proxy_async_smem();
    tc::mbar_arrive(bar_v);
  }
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 8 && tc::elect_one()) {
    // ------------------------------ MMA issuer (P V) ------------------------------
    const uint32_t idesc_pv = tc::make_idesc_bf16(128, 80, 0, 1);
    tc::mbar_wait(bar_v, 0);
    WIN_T(10);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tc::mbar_wait(bar_p(h), 0);
      tc::tc_fence_after();
      WIN_T(11 + h);
      const uint32_t pa = tmem_base + 256u * h;  // P_h: 128 columns of bf16 pairs = 256 keys
      const uint32_t d = pa + 128u;              // O_h
#pragma unroll
      for (int ks = 0; ks < kWinTok / 16; ++ks) {
        // V is MN-major: 8-key groups 1024 B apart (SBO), 64-dim blocks kBlkBytes apart (LBO)
        const uint64_t db = tc::make_desc_sw128(sv + ks * 2048, kBlkBytes, 1024);
        tc::umma_ts(d, pa + ks * 8, db, idesc_pv, ks != 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_o(h));
    }
  } else if (warp < 8) {
    // ------------------------------ softmax / epilogue ------------------------------
    const int h = warp >> 2;
    const int lq = warp & 3;
    const uint32_t ts = tmem_base + 256u * h + (static_cast<uint32_t>(lq * 32) << 16);
    tc::mbar_wait(bar_s(h), 0);
    tc::tc_fence_after();
    if (threadIdx.x == 0) WIN_T(4);
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
    for (int c = 0; c < kWinTok / 32; ++c) {
      uint32_t s[32];
      tc::tmem_ld32(ts + c * 32, s);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(s[i]));
    }
    const float moff = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;
    if (threadIdx.x == 0) WIN_T(5);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < kWinTok / 32; ++c) {
      uint32_t s[32];
      tc::tmem_ld32(ts + c * 32, s);
      tc::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e0 = ex2_fast(fmaf(__uint_as_float(s[2 * i]), p.scale_log2, -moff));
        const float e1 = ex2_fast(fmaf(__uint_as_float(s[2 * i + 1]), p.scale_log2, -moff));
        sum0 += e0;
        sum1 += e1;
        pk[i] = tc::pack_bf16(e0, e1);
      }
      tc::tmem_st16(ts + c * 16, pk);  // columns [16c, 16c+16) <= columns already read ([0, 32c+32))
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar_p(h));
    if (threadIdx.x == 0) WIN_T(6);
    const float inv = 1.0f / (sum0 + sum1);
    tc::mbar_wait(bar_o(h), 0);
    tc::tc_fence_after();
    if (threadIdx.x == 0) WIN_T(7);
    // O rows go through shared memory (the Q tiles are dead once both Q K^T have retired) so that the global
    // stores are whole 144-byte token segments spread over consecutive lanes instead of one 16-byte piece per
    // lane at a 1152-byte stride
    const int nchunk = p.D >> 3;
    const uint32_t stg = sq + warp * (32 * 160);  // 32 rows x 160 B per warp
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // 80 accumulator columns: 32 + 32 + 16
      uint32_t o[32];
      if (c < 2) {
        tc::tmem_ld32(ts + 128 + c * 32, o);
      } else {
        uint32_t o16[16];
        tc::tmem_ld16(ts + 128 + 64, o16);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = o16[i];
#pragma unroll
        for (int i = 16; i < 32; ++i) o[i] = 0u;
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (c * 4 + i < 10) {
          const uint32_t x = tc::pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          const uint32_t y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          const uint32_t z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          const uint32_t w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 160 + (c * 4 + i) * 16), "r"(x), "r"(y),
                       "r"(z), "r"(w)
                       : "memory");
        }
      }
    }
    __syncwarp();
    const int r_base = h * 128 + lq * 32;  // first query row of this warp
    for (int i = lane; i < 32 * nchunk; i += 32) {
      const int rr = i / nchunk, cc = i - rr * nchunk;
      const int rw = r_base + rr;
      uint4 t;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(stg + rr * 160 + cc * 16));
      const long long tok = static_cast<long long>(wy * kWinSide + (rw >> 4)) * p.Wm + wx * kWinSide + (rw & 15);
      reinterpret_cast<uint4*>(p.out + b * p.o_bs + tok * p.o_tok + head * p.D)[cc] = t;
    }
  }

  if (threadIdx.x == 0) WIN_T(8);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Global (non-windowed) Hiera attention, head_dim <= 80: flash loop over 128-key tiles.
//
// One CTA = 128 queries of one head; 160 threads; 96 KB of shared memory and 256 TMEM columns, so that TWO
// CTAs share an SM and one's staging / softmax overlaps the other's MMAs (the per-CTA chain is serial:
// stage K,V -> S = Q K^T -> softmax -> O += P V).  K/V rows of the next tile are prefetched into registers
// while the current tile is being processed.  TMEM: S / P at columns [0,128), O at [128,208).
// ---------------------------------------------------------------------------------------------
struct GlobAttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* out;
  long long q_tok, k_tok, v_tok, o_tok;
  long long q_bs, k_bs, v_bs, o_bs;
  int H, D, Lq, Lk;
  float scale_log2;
  int dbg;
};
#define GLOB_T(i)                                                                          \
  do {                                                                                     \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && j == 4 && threadIdx.x == 0) g_win_times[i] = clock64() - t_start; \
  } while (0)

constexpr int kGlobThreads = 160;
constexpr int kGlobBlk = 128 * 128;  // one 64-dim block of a 128-row tile (16 KB)
constexpr int kGlobSmem = 6 * kGlobBlk + 1024 + 128;

__global__ void __launch_bounds__(kGlobThreads, 2) glob_attn_tc_kernel(const GlobAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sq = base, sk = base + 2 * kGlobBlk, sv = base + 4 * kGlobBlk;
  const uint32_t bar_base = base + 6 * kGlobBlk;
  const uint32_t bar_s = bar_base, bar_p = bar_base + 8u, bar_o = bar_base + 16u;
  const uint32_t tmem_slot = bar_base + 24u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = p.Lk / 128;

  if (threadIdx.x == 0) {
    tc::mbar_init(bar_s, 1);
    tc::mbar_init(bar_p, 4);
    tc::mbar_init(bar_o, 1);
    tc::fence_barrier_init();
  }
  if (warp == 4) {
    tc::tmem_alloc(tmem_slot, 256);
    tc::tmem_relinquish();
  }
  pdl_sync();

  // staging map: thread -> chunk c (16 bytes = 8 dims) and rows r0 + 16 t of a 128-row tile
  const int nchunk = p.D >> 3;
  const int c = threadIdx.x % 10;
  const int r0 = threadIdx.x / 10;  // 0..15
  const bool ld_on = c < nchunk;
  auto put = [&](uint32_t tile, int r, const uint4& x) {
    const uint32_t dst = tile + (c >> 3) * kGlobBlk + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
  };
  auto load8 = [&](const __nv_bfloat16* src, long long tok_stride, int tok0, uint4(&dst)[8]) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      dst[t] = make_uint4(0u, 0u, 0u, 0u);
      if (ld_on) dst[t] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(tok0 + r0 + 16 * t) * tok_stride));
    }
  };
  const __nv_bfloat16* qsrc = p.q + b * p.q_bs + head * p.D + c * 8;
  const __nv_bfloat16* ksrc = p.k + b * p.k_bs + head * p.D + c * 8;
  const __nv_bfloat16* vsrc = p.v + b * p.v_bs + head * p.D + c * 8;
  uint4 kreg[8], vreg[8];
  {
    uint4 qreg[8];
    load8(qsrc, p.q_tok, q0, qreg);
    load8(ksrc, p.k_tok, 0, kreg);
    load8(vsrc, p.v_tok, 0, vreg);
#pragma unroll
    for (int t = 0; t < 8; ++t) put(sq, r0 + 16 * t, qreg[t]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t idesc_qk = tc::make_idesc_bf16(128, 128, 0, 0);
  const uint32_t idesc_pv = tc::make_idesc_bf16(128, 80, 0, 1);
  const long long t_start = clock64();

  const uint32_t ts = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t to = ts + 128;
  float m_ref = -INFINITY, l = 0.f;

  for (int j = 0; j < n_tiles; ++j) {
    // K/V smem and the S/P columns are free once P V of the previous tile has retired
    GLOB_T(0);
    if (j > 0) tc::mbar_wait(bar_o, (j - 1) & 1);
    tc::tc_fence_after();
    GLOB_T(1);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      put(sk, r0 + 16 * t, kreg[t]);
      put(sv, r0 + 16 * t, vreg[t]);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    GLOB_T(2);
    // Q K^T goes to the tensor pipe FIRST; only then does anybody issue the next tile's 16 global loads (they used to sit
    // between the barrier and the MMA issue: 1 600 clk of every 6 500-clk tile on the critical path, DS2_WIN_DBG=1)
    // elect.sync evaluated in place (a cached predicate makes the compiler serialise every tcgen05.mma)
    if (warp == 4 && tc::elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const uint32_t off = ks < 4 ? ks * 32 : kGlobBlk;
        tc::umma_ss(tmem_base, tc::make_desc_sw128(sq + off, 16, 1024), tc::make_desc_sw128(sk + off, 16, 1024), idesc_qk,
                    ks != 0 ? 1u : 0u);
      }
      tc::umma_commit(bar_s);
    }
    __syncwarp();
    if (j + 1 < n_tiles) {  // next tile's rows: in flight during this tile's MMAs and softmax
      load8(ksrc, p.k_tok, (j + 1) * 128, kreg);
      load8(vsrc, p.v_tok, (j + 1) * 128, vreg);
    }
    GLOB_T(3);
    if (warp == 4 && tc::elect_one()) {
      tc::mbar_wait(bar_p, j & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        tc::umma_ts(tmem_base + 128, tmem_base + ks * 8, tc::make_desc_sw128(sv + ks * 2048, kGlobBlk, 1024), idesc_pv,
                    (j | ks) != 0 ? 1u : 0u);
      tc::umma_commit(bar_o);
    }
    if (warp < 4) {
      tc::mbar_wait(bar_s, j & 1);
      tc::tc_fence_after();
      GLOB_T(4);
      // two passes over the 128 scores of the row (TMEM reads are cheap; the prefetched K/V rows of the next
      // tile already hold 64 registers): pass 1 row max, pass 2 ex2 / sum / bf16 P written behind the reads
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t sr[32];
        tc::tmem_ld32(ts + cc * 32, sr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(sr[i]));
      }
      const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      GLOB_T(5);
      float alpha = 1.f;
      bool resc = false;
      if (j == 0) {
        m_ref = mt;
      } else if ((mt - m_ref) * p.scale_log2 > 8.0f) {  // lazy rescaling: P stays below 2^8
        alpha = ex2_fast((m_ref - mt) * p.scale_log2);
        m_ref = mt;
        resc = true;
      }
      const float moff = m_ref * p.scale_log2;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t sr[32];
        tc::tmem_ld32(ts + cc * 32, sr);
        tc::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2_fast(fmaf(__uint_as_float(sr[2 * i]), p.scale_log2, -moff));
          const float e1 = ex2_fast(fmaf(__uint_as_float(sr[2 * i + 1]), p.scale_log2, -moff));
          sum0 += e0;
          sum1 += e1;
          pk[i] = tc::pack_bf16(e0, e1);
        }
        tc::tmem_st16(ts + cc * 16, pk);  // columns [16cc, 16cc+16) lie inside what was already read
      }
      l = l * alpha + (sum0 + sum1);
      if (__any_sync(0xffffffffu, resc)) {
        // O is stable: P V of tile j-1 retired (waited for at the top of the loop), P V of tile j not issued
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          if (cc < 2) {
            uint32_t o[32];
            tc::tmem_ld32(to + cc * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st32(to + cc * 32, o);
          } else {
            uint32_t o[16];
            tc::tmem_ld16(to + 64, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st16(to + 64, o);
          }
        }
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_p);
      GLOB_T(6);
    }
  }
  // ---- epilogue: O / l -> bf16 -> global (through the dead Q tile for whole-segment stores) ----
  tc::mbar_wait(bar_o, (n_tiles - 1) & 1);
  tc::tc_fence_after();
  if (warp < 4) {
    const float inv = 1.0f / l;
    const uint32_t stg = sq + warp * (32 * 160);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      uint32_t o[32];
      if (cc < 2) {
        tc::tmem_ld32(to + cc * 32, o);
      } else {
        uint32_t o16[16];
        tc::tmem_ld16(to + 64, o16);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = o16[i];
#pragma unroll
        for (int i = 16; i < 32; ++i) o[i] = 0u;
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (cc * 4 + i < 10) {
          const uint32_t x = tc::pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          const uint32_t y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          const uint32_t z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          const uint32_t w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 160 + (cc * 4 + i) * 16), "r"(x), "r"(y),
                       "r"(z), "r"(w)
                       : "memory");
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < 32 * nchunk; i += 32) {
      const int rr = i / nchunk, cc = i - rr * nchunk;
      uint4 t;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(stg + rr * 160 + cc * 16));
      const long long tok = q0 + warp * 32 + rr;
      reinterpret_cast<uint4*>(p.out + b * p.o_bs + tok * p.o_tok + head * p.D)[cc] = t;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

int launch_glob_attn_tc(const ds2_mha_args* a, cudaStream_t st) {
  if (a->window != 0 || a->q_pool) return -1;
  if (a->D <= 64 || a->D > 80 || (a->D % 8) != 0) return -1;
  if (a->Lq < 128 || (a->Lq % 128) != 0 || a->Lk < 128 || (a->Lk % 128) != 0) return -1;
  if (a->Lk_valid != 0 && a->Lk_valid != a->Lk) return -1;
  if ((a->q_tok_stride % 8) || (a->k_tok_stride % 8) || (a->v_tok_stride % 8) || (a->o_tok_stride % 8)) return -1;
  if ((a->q_bs % 8) || (a->k_bs % 8) || (a->v_bs % 8) || (a->o_bs % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15)
    return -1;
  if (a->H > 65535 || a->B > 65535) return -1;
  GlobAttnParams p;
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
  p.H = a->H;
  p.D = a->D;
  p.Lq = a->Lq;
  p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  static const bool dbg = [] {
    const char* e = getenv("DS2_WIN_DBG");
    return e && e[0] == '1';
  }();
  p.dbg = dbg ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(glob_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlobSmem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_mha: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  DS2_LAUNCH((glob_attn_tc_kernel), dim3(a->Lq / 128, a->H, a->B), kGlobThreads, kGlobSmem, st, p);
  return post_launch("glob_attn_tc_kernel");
}

// Returns DS2_OK after launching, or -1 when the shape is not one this kernel covers (caller falls back
// to the generic mma.sync kernel in mha.cu).
int launch_win16_attn_tc(const ds2_mha_args* a, cudaStream_t st) {
  if (a->window != kWinSide || a->q_pool) return -1;
  if (a->D <= 64 || a->D > 80 || (a->D % 8) != 0) return -1;
  if ((a->Hm % kWinSide) != 0 || (a->Wm % kWinSide) != 0) return -1;
  if ((a->q_tok_stride % 8) || (a->k_tok_stride % 8) || (a->v_tok_stride % 8) || (a->o_tok_stride % 8)) return -1;
  if ((a->q_bs % 8) || (a->k_bs % 8) || (a->v_bs % 8) || (a->o_bs % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15)
    return -1;
  WinAttnParams p;
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
  p.H = a->H;
  p.D = a->D;
  p.Wm = a->Wm;
  p.nwx = a->Wm / kWinSide;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  static const bool dbg = [] {
    const char* e = getenv("DS2_WIN_DBG");
    return e && e[0] == '1';
  }();
  p.dbg = dbg ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(win16_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWinSmem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_mha: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  p.nwin = (a->Hm / kWinSide) * (a->Wm / kWinSide);
  DS2_LAUNCH((win16_attn_tc_kernel), a->B * p.nwin * a->H, kWinThreads, kWinSmem, st, p);
  return post_launch("win16_attn_tc_kernel");
}

}  // namespace ds2

extern "C" int ds2_debug_win_times(long long* out16) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, ds2::g_win_times, 16 * sizeof(long long));
  return e == cudaSuccess ? DS2_OK : static_cast<int>(e);
}
