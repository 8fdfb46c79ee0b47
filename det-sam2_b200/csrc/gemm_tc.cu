// gemm_tc.cu — persistent, warp-specialised bf16 GEMM on tcgen05 tensor cores (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] · W[N,K]^T )         A, W bf16 K-major; f32 accumulate in TMEM
//
// CTA = 384 threads, one CTA per SM, static round-robin over 128 x BN output tiles.
//   warp 0       : TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle, 4-stage ring)
//   warp 1       : MMA issuer     (tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction)
//   warp 2       : TMEM allocator (512 columns = 2 accumulator stages of <=256 columns)
//   warps 4..11  : epilogue, two warps per TMEM lane quarter, alternating 32-column chunks:
//                  tcgen05.ld 32x32b -> bias / act / gamma / rotary in registers (thread = row) ->
//                  swizzled shared staging -> TMA store (cp.async.bulk.tensor, fully coalesced) or,
//                  for the in-place residual add x += f(x), TMA reduce-add (cp.reduce.async.bulk .add
//                  f32 executed at L2), so no SM ever reads the residual.  Row-per-thread global
//                  stores are kept only as the fallback for odd cases (two outputs, row-periodic
//                  residual tables, unaligned pitches).
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
// tile i+1.  BN is chosen per problem to minimise waves x tile width.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "tc05.cuh"

namespace ds2 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kBBytesMax = 256 * kBK * 2;     // 32 KB
constexpr int kStageBytes = kABytes + kBBytesMax;
constexpr int kEpiWarps = 8;
constexpr int kStagingBytes = 4096;           // 32 rows x 128 B per epilogue warp
constexpr int kGemmSmem = kStages * kStageBytes + kEpiWarps * kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;

enum StoreMode { kStoreDirect = 0, kStoreTmaF32 = 1, kStoreTmaAddF32 = 2, kStoreTmaBf16 = 3 };

struct GemmParams {
  int M, N, K, BN;
  int tiles_m, tiles_n;
  int store_mode;
  const float* bias;
  const float* gamma;
  const float* residual;
  long long ldr;
  int res_row_mod;
  int act;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  long long ldc, ldc_bf16;
  const float* rope_cs;
  int rope_col0, rope_col1, rope_period, rope_rows_per_batch, rope_row_limit;
  // axial form of the rotary table: [64][side] (cos, sin) indexed by ONE grid coordinate — pairs 0..63 rotate by the x
  // coordinate of the position, pairs 64..127 by its y coordinate, with the same 64 frequencies (position_encoding.py:
  // 173-182), so the [128][side^2] table is this 32 KB table read two ways.  It is staged in shared memory (the
  // pipeline runs with 3 stages then: rotary GEMMs have K <= 256) instead of 128 KB of table per 128-row tile coming
  // out of L2 with one global load per pair.
  const float* rope_axial;
  int rope_side;
  int stages;
};

__device__ __forceinline__ float gelu_erf(float x) { return gelu_erf_fast(x); }

// bias / activation / gamma / rotary on NC consecutive columns of one row (registers).
// Rotary state of one output row (= one epilogue thread for a whole tile): computed once per tile, not per chunk — the
// three integer divisions by run-time values cost as many issue slots as the rotation of a 32-column chunk itself.
struct RopeRow {
  uint32_t x_off, y_off;  // byte offsets of this row's x / y coordinate inside one pair's [side] (cos, sin) line
  bool on;                // false: row beyond rope_row_limit (object-pointer tokens) -> no rotation
};
__device__ __forceinline__ RopeRow rope_row(const GemmParams& p, int row) {
  RopeRow r;
  const int rb = row % p.rope_rows_per_batch;
  const int pos = rb % p.rope_period;
  r.on = rb < p.rope_row_limit;
  r.x_off = static_cast<uint32_t>(pos % p.rope_side) * 8u;
  r.y_off = static_cast<uint32_t>(pos / p.rope_side) * 8u;
  return r;
}

template <int NC>
__device__ __forceinline__ void epilogue_math(const GemmParams& p, const uint32_t* acc, int row, int col0,
                                              float (&v)[NC], uint32_t rope_smem = 0, RopeRow rr = RopeRow{0u, 0u, false}) {
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(acc[i]);
  const bool full = (col0 + NC <= p.N);
  if (p.bias) {
    if (full && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
        v[4 * i] += t.x;
        v[4 * i + 1] += t.y;
        v[4 * i + 2] += t.z;
        v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (p.act == 2) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = gelu_erf(v[i]);
  }
  if (p.gamma) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (full || col0 + i < p.N) v[i] *= __ldg(p.gamma + col0 + i);
  }
  if (p.rope_axial) {
    if (rr.on && col0 >= p.rope_col0 && col0 < p.rope_col1) {
      const int pair0 = ((col0 - p.rope_col0) & 255) >> 1;        // 16 | 64: a chunk lies in ONE half of the pairs
      // lanes = consecutive rows = consecutive x (conflict-free 8-byte reads) or one shared y (broadcast)
      const uint32_t pitch = static_cast<uint32_t>(p.rope_side) * 8u;
      const uint32_t base = rope_smem + static_cast<uint32_t>(pair0 & 63) * pitch + (pair0 < 64 ? rr.x_off : rr.y_off);
      float cx[NC / 2], cy[NC / 2];
#pragma unroll
      for (int i = 0; i < NC / 2; ++i)
        // volatile: must stay behind the named barrier that publishes the table (volatile asm keeps its order)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cx[i]), "=f"(cy[i]) : "r"(base + static_cast<uint32_t>(i) * pitch));
#pragma unroll
      for (int i = 0; i < NC / 2; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        v[2 * i] = a * cx[i] - b * cy[i];
        v[2 * i + 1] = a * cy[i] + b * cx[i];
      }
    }
  } else if (p.rope_cs && col0 >= p.rope_col0 && col0 < p.rope_col1) {
    const int rb = row % p.rope_rows_per_batch;
    if (rb < p.rope_row_limit) {
      // table is pair-major [128][period]: the 32 lanes of a warp (= 32 consecutive rows = consecutive
      // positions) read 256 contiguous bytes per pair instead of 32 different table rows
      const int pos = rb % p.rope_period;
      const float2* cs = reinterpret_cast<const float2*>(p.rope_cs) +
                         static_cast<size_t>(((col0 - p.rope_col0) & 255) >> 1) * p.rope_period + pos;
#pragma unroll
      for (int i = 0; i < NC / 2; ++i) {
        const float2 c = __ldg(cs + static_cast<size_t>(i) * p.rope_period);
        const float a = v[2 * i], b = v[2 * i + 1];
        v[2 * i] = a * c.x - b * c.y;
        v[2 * i + 1] = a * c.y + b * c.x;
      }
    }
  }
}

// Fallback store: residual add + row-per-thread global stores.
template <int NC>
__device__ __forceinline__ void epilogue_store_direct(const GemmParams& p, float (&v)[NC], int row, int col0) {
  const bool full = (col0 + NC <= p.N);
  if (p.residual) {
    const long long rr = p.res_row_mod > 0 ? (row % p.res_row_mod) : row;
    const float* r = p.residual + rr * p.ldr + col0;
    if (full && ((p.ldr & 3) == 0)) {
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r) + i);
        v[4 * i] += t.x;
        v[4 * i + 1] += t.y;
        v[4 * i + 2] += t.z;
        v[4 * i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (col0 + i < p.N) v[i] += __ldg(r + i);
    }
  }
  if (p.out_f32) {
    float* o = p.out_f32 + static_cast<long long>(row) * p.ldc + col0;
    if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
      for (int i = 0; i < NC / 4; ++i)
        reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (col0 + i < p.N) o[i] = v[i];
    }
  }
  if (p.out_bf16) {
    __nv_bfloat16* o = p.out_bf16 + static_cast<long long>(row) * p.ldc_bf16 + col0;
    if (full && ((p.ldc_bf16 & 7) == 0)) {
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) {
        uint4 t;
        t.x = tc::pack_bf16(v[8 * i], v[8 * i + 1]);
        t.y = tc::pack_bf16(v[8 * i + 2], v[8 * i + 3]);
        t.z = tc::pack_bf16(v[8 * i + 4], v[8 * i + 5]);
        t.w = tc::pack_bf16(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(o)[i] = t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (col0 + i < p.N) o[i] = __float2bfloat16(v[i]);
    }
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + kStages * kStageBytes;
  const uint32_t bar_base = staging_base + kEpiWarps * kStagingBytes;
  // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
    if (p.store_mode != kStoreDirect) tc::prefetch_tmap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(full_bar(s), 1);
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tfull_bar(s), 1);
      tc::mbar_init(tempty_bar(s), kEpiWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // prologue done (barriers, TMEM, descriptor prefetch: nothing that depends on the previous grid);
  // only now wait for the producer of our operands, and let the next grid start its own prologue
  pdl_sync();

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int k_blocks = (p.K + kBK - 1) / kBK;
  const uint32_t stage_tx = kABytes + static_cast<uint32_t>(p.BN) * kBK * 2;
  const int nstages = p.stages;                                  // 4, or 3 when the rotary table takes the 4th slot
  const uint32_t rope_smem = smem_base + 3u * kStageBytes;

  // Role gates use elect.sync, not `lane == 0`: tcgen05.mma / TMA take their operands from the uniform
  // datapath, and under a lane-id predicate the compiler wraps EVERY such instruction in an
  // ELECT / BRA.U.ANY serialisation loop (~97 clk per MMA issued, measured with tools/mma_rate.cu, against
  // 42-74 clk for the instruction itself) — the single issuing thread, not the tensor pipe, set the pace.
  if (warp == 0 && tc::elect_one()) {
    // ---------------- TMA producer ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(empty_bar(stage), phase ^ 1);
        tc::mbar_expect_tx(full_bar(stage), stage_tx);
        const uint32_t sa = smem_base + stage * kStageBytes;
        tc::tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBK, m_blk * kBM);
        tc::tma_load_2d(sa + kABytes, &tmap_w, full_bar(stage), kb * kBK, n_blk * p.BN);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && tc::elect_one()) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = tc::make_idesc_bf16(kBM, p.BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(full_bar(stage), phase);
        tc::tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t da = tc::make_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = tc::make_desc_sw128(sb + k * 32, 16, 1024);
          tc::umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc::umma_commit(empty_bar(stage));
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc::umma_commit(tfull_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int ew = warp - 4;       // 0..7
    const int lq = warp & 3;       // TMEM lane quarter this warp may touch: lanes [32*lq, 32*lq+32)
    const int half = ew >> 2;      // which alternate 32-column chunks this warp takes
    const uint32_t stg = staging_base + static_cast<uint32_t>(ew) * kStagingBytes;
    const int mode = p.store_mode;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool pending = false;  // lane 0: a bulk store may still be reading the staging buffer
    if (p.rope_axial) {
      // the table is a constant (not produced by the previous grid): nothing here waits for pdl — but every epilogue
      // warp must see all of it, hence the named barrier over the 8 epilogue warps
      const int n4 = 64 * p.rope_side * 2 / 4;
      const float4* src = reinterpret_cast<const float4*>(p.rope_axial);
      for (int i = threadIdx.x - 128; i < n4; i += 32 * kEpiWarps) {
        const float4 t4 = __ldg(src + i);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rope_smem + 16u * i), "f"(t4.x), "f"(t4.y), "f"(t4.z),
                     "f"(t4.w) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::tc_fence_after();
      const int row0 = m_blk * kBM + lq * 32;
      const int row = row0 + lane;
      const uint32_t t_addr =
          tmem_base + static_cast<uint32_t>(acc) * 256u + (static_cast<uint32_t>(lq * 32) << 16);
      const int n0 = n_blk * p.BN;
      // loads 32 (or a 16-wide tail, or no) accumulator columns starting at column c of the tile
      auto load_cols = [&](int c, uint32_t(&r)[32]) {
        const int left = p.BN - c;
        if (left >= 32) {
          tc::tmem_ld32(t_addr + c, r);
        } else if (left >= 16) {
          uint32_t r16[16];
          tc::tmem_ld16(t_addr + c, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            r[i] = r16[i];
            r[16 + i] = 0u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
      };
      auto wait_staging = [&]() {
        // the previous bulk store must have finished READING the staging buffer
        if (pending) {
          tc::bulk_wait_read0();
          pending = false;
        }
        __syncwarp();
      };
      auto issue_store = [&](int c) {
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (mode == kStoreTmaAddF32) tc::tma_reduce_add_2d(&tmap_c, stg, n0 + c, row0);
          else tc::tma_store_2d(&tmap_c, stg, n0 + c, row0);
          tc::bulk_commit();
          pending = true;
        }
      };
      const int mrow = row < p.M ? row : 0;
      const RopeRow rr = p.rope_axial ? rope_row(p, mrow) : RopeRow{0u, 0u, false};
      if (mode == kStoreTmaBf16) {
        // 64-column chunks: 32 rows x 128 B staging, SWIZZLE_128B (16-byte chunk j of row r at j ^ (r & 7)),
        // one TMA store of a full 128-byte line per row
        for (int c = half * 64; c < p.BN; c += 128) {
          uint32_t r0[32], r1[32];
          load_cols(c, r0);
          load_cols(c + 32, r1);
          tc::tmem_ld_wait();
          if (n0 + c >= p.N) continue;  // warp-uniform
          uint32_t pk[32];
          {
            float v[32];
            epilogue_math<32>(p, r0, mrow, n0 + c, v, rope_smem, rr);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = tc::pack_bf16(v[2 * i], v[2 * i + 1]);
          }
          if (n0 + c + 32 < p.N && c + 32 < p.BN) {
            float v[32];
            epilogue_math<32>(p, r1, mrow, n0 + c + 32, v, rope_smem, rr);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[16 + i] = tc::pack_bf16(v[2 * i], v[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[16 + i] = 0u;
          }
          wait_staging();
          const uint32_t rowaddr = stg + static_cast<uint32_t>(lane) * 128u;
          const uint32_t sw = static_cast<uint32_t>(lane) & 7u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(rowaddr + ((static_cast<uint32_t>(j) ^ sw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2],
                         pk[4 * j + 3]);
          issue_store(c);
        }
      } else {
        for (int c = half * 32; c < p.BN; c += 64) {
          uint32_t r[32];
          load_cols(c, r);
          tc::tmem_ld_wait();
          if (n0 + c >= p.N) continue;  // warp-uniform
          if (mode == kStoreDirect) {
            if (row < p.M) {
              float v[32];
              epilogue_math<32>(p, r, row, n0 + c, v, rope_smem, rr);
              epilogue_store_direct<32>(p, v, row, n0 + c);
            }
            continue;
          }
          float v[32];
          epilogue_math<32>(p, r, mrow, n0 + c, v, rope_smem, rr);
          wait_staging();
          // 32 rows x 128 B, SWIZZLE_128B
          const uint32_t rowaddr = stg + static_cast<uint32_t>(lane) * 128u;
          const uint32_t sw = static_cast<uint32_t>(lane) & 7u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(rowaddr + ((static_cast<uint32_t>(j) ^ sw) << 4), __float_as_uint(v[4 * j]),
                         __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          issue_store(c);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (pending) tc::bulk_wait0();  // global writes complete before the CTA retires
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}


// =================================================================================================
// gemm2: the same GEMM on CTA PAIRS (tcgen05 cta_group::2), the default for M >= 512
// =================================================================================================
// A cluster of two CTAs (two SMs of one TPC) owns a 256 x BN output tile: CTA r holds rows [128 r, 128 r + 128) of A
// and rows [r BN/2, (r + 1) BN/2) of W in its shared memory, ONE thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' operands and writes both CTAs' TMEM, and each CTA runs
// the epilogue of its own 128 rows.  Per k-block a CTA loads 16 KB of A + <= 16 KB of W instead of 16 + 32 KB: the
// Hiera / memory-attention GEMMs have short K (256 ... 2304) and were bound by L2 -> SM operand traffic (510 MB per
// 16384 x 2304 x 576 launch = 12.4 TB/s, the L2 fabric limit, at 1.0 PFLOP/s; profiles/r2_s5_gemm_probe.txt).
//   warp 0        : TMA producer (both CTAs; transaction bytes of both land on the LEADER's full barrier)
//   warp 1        : MMA issuer (leader CTA only); commits multicast to both CTAs' barriers
//   warp 2        : TMEM allocation (512 columns, cta_group::2, both CTAs)
//   warps 4..19   : epilogue, FOUR warps per TMEM lane quarter: 16 warps hide the latencies (TMEM load, bias, MUFU of
//                   the GELU, shared-memory staging) that 8 warps exposed — the v1 epilogue issued on 40 % of the
//                   cycles of its two warps per scheduler (ncu source page, profiles/r2_s5_gemm_epilogue_ncu.txt).
//                   A warp owns 32-column chunks ci = cg, cg + 4, ...; bias / gamma of its chunks are fetched BEFORE it
//                   waits for the accumulator and re-read from a warp-private shared-memory line; a warp owns 64
//                   adjacent columns of the tile and sends them out through its 4 KB staging box (128-byte swizzle):
//                   one TMA store of 64 bf16 columns, or two TMA stores / reduce-adds of 32 f32 columns.
constexpr int k2Stages = 4;
constexpr int k2ABytes = kBM * kBK * 2;          // 16 KB: this CTA's 128 rows of A
constexpr int k2BBytesMax = 128 * kBK * 2;       // 16 KB: this CTA's half (<= 128 rows) of the W tile
constexpr int k2StageBytes = k2ABytes + k2BBytesMax;
constexpr int k2EpiWarps = 16;
constexpr int k2StagingBytes = 4096;             // per warp: one box of 32 rows x 128 B
constexpr int k2VecBytes = 512;                  // per warp: bias (64 f32) + gamma (64 f32)
constexpr int k2Smem = k2Stages * k2StageBytes + k2EpiWarps * (k2StagingBytes + k2VecBytes) + 1024 + 256;
constexpr int k2Threads = 128 + 32 * k2EpiWarps;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data into THIS CTA's shared memory, transaction bytes onto the barrier at `bar` (a shared::cluster
// address, here the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// bias / activation / gamma / rotary on 32 consecutive columns of one row, in place; bias and gamma come from the warp's
// shared-memory line `vec` (64 f32 bias, then 64 f32 gamma) at element offset `voff`
__device__ __forceinline__ void epilogue_math_v2(const GemmParams& p, float (&v)[32], int col0, uint32_t vec, int voff,
                                                 uint32_t rope_smem, const RopeRow& rr) {
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float b0, b1, b2, b3;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(vec + 4u * (voff + 4 * i)));
      v[4 * i] += b0;
      v[4 * i + 1] += b1;
      v[4 * i + 2] += b2;
      v[4 * i + 3] += b3;
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (p.act == 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (p.gamma) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float g0, g1, g2, g3;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3) : "r"(vec + 256u + 4u * (voff + 4 * i)));
      v[4 * i] *= g0;
      v[4 * i + 1] *= g1;
      v[4 * i + 2] *= g2;
      v[4 * i + 3] *= g3;
    }
  }
  if (p.rope_axial && rr.on && col0 >= p.rope_col0 && col0 < p.rope_col1) {
    const int pair0 = ((col0 - p.rope_col0) & 255) >> 1;
    const uint32_t pitch = static_cast<uint32_t>(p.rope_side) * 8u;
    const uint32_t base = rope_smem + static_cast<uint32_t>(pair0 & 63) * pitch + (pair0 < 64 ? rr.x_off : rr.y_off);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float cx, cy;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cx), "=f"(cy) : "r"(base + static_cast<uint32_t>(i) * pitch));
      const float a = v[2 * i], b = v[2 * i + 1];
      v[2 * i] = a * cx - b * cy;
      v[2 * i + 1] = a * cy + b * cx;
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                               const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + k2Stages * k2StageBytes;
  const uint32_t vec_base = staging_base + k2EpiWarps * k2StagingBytes;
  const uint32_t bar_base = vec_base + k2EpiWarps * k2VecBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (k2Stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * k2Stages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
    if (p.store_mode != kStoreDirect) tc::prefetch_tmap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      tc::mbar_init(full_bar(s), 1);     // leader: its own producer's arrive.expect_tx (bytes of BOTH CTAs)
      tc::mbar_init(empty_bar(s), 1);    // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tfull_bar(s), 1);                  // one multicast commit per tile
      tc::mbar_init(tempty_bar(s), 2 * k2EpiWarps);    // leader: the epilogue warps of both CTAs
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // both CTAs' barriers are initialised and both TMEM allocations done before any remote access
  tc::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_sync();

  const int num_tiles = p.tiles_m * p.tiles_n;       // tiles of 256 x BN
  const int k_blocks = (p.K + kBK - 1) / kBK;
  const int half_bn = p.BN >> 1;
  const uint32_t cta_tx = k2ABytes + static_cast<uint32_t>(half_bn) * kBK * 2;
  const int nstages = p.stages;
  const uint32_t rope_smem = smem_base + 3u * k2StageBytes;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && tc::elect_one()) {
    // ---------------- TMA producer (both CTAs) ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      const int row_a = m_blk * 256 + static_cast<int>(rank) * kBM;
      const int row_w = n_blk * p.BN + static_cast<int>(rank) * half_bn;
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(empty_bar(stage), phase ^ 1);
        if (leader) tc::mbar_expect_tx(full_bar(stage), 2u * cta_tx);
        const uint32_t sa = smem_base + stage * k2StageBytes;
        const uint32_t lead_full = full_bar(stage) & kPeerMask;
        tma_load_2d_2sm(sa, &tmap_a, lead_full, kb * kBK, row_a);
        tma_load_2d_2sm(sa + k2ABytes, &tmap_w, lead_full, kb * kBK, row_w);
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader && tc::elect_one()) {
    // ---------------- MMA issuer (leader CTA) ----------------
    const uint32_t idesc = tc::make_idesc_bf16(256, p.BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * 256u;
      for (int kb = 0; kb < k_blocks; ++kb) {
        tc::mbar_wait(full_bar(stage), phase);
        tc::tc_fence_after();
        const uint32_t sa = smem_base + stage * k2StageBytes;
        const uint32_t sb = sa + k2ABytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t da = tc::make_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t db = tc::make_desc_sw128(sb + k * 32, 16, 1024);
          umma_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit_2sm(empty_bar(stage));
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_2sm(tfull_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (both CTAs, own 128 rows) ----------------
    const int ew = warp - 4;       // 0..15
    const int lq = warp & 3;       // TMEM lane quarter
    const int cg = ew >> 2;        // column group: 32-column chunks cg, cg + 4, ...
    const uint32_t stg = staging_base + static_cast<uint32_t>(ew) * k2StagingBytes;
    const uint32_t vec = vec_base + static_cast<uint32_t>(ew) * k2VecBytes;
    const int mode = p.store_mode;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool pending = false;          // lane 0: a bulk store may still be reading the staging box
    if (p.rope_axial) {
      const int n4 = 64 * p.rope_side * 2 / 4;
      const float4* src = reinterpret_cast<const float4*>(p.rope_axial);
      for (int i = threadIdx.x - 128; i < n4; i += 32 * k2EpiWarps) {
        const float4 t4 = __ldg(src + i);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rope_smem + 16u * i), "f"(t4.x), "f"(t4.y), "f"(t4.z),
                     "f"(t4.w) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * k2EpiWarps) : "memory");
    }
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int m_blk = t / p.tiles_n, n_blk = t % p.tiles_n;
      const int n0 = n_blk * p.BN;
      const int row0 = m_blk * 256 + static_cast<int>(rank) * kBM + lq * 32;
      const int row = row0 + lane;
      const int mrow = row < p.M ? row : 0;
      // bias / gamma of this warp's (<= 2) chunks: fetched before the accumulator wait, kept in a warp-private line
      float bv[2] = {0.f, 0.f}, gv[2] = {1.f, 1.f};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = n0 + cg * 64 + 32 * j + lane;
        if (cg * 64 + 32 * j < p.BN && c < p.N) {
          if (p.bias) bv[j] = __ldg(p.bias + c);
          if (p.gamma) gv[j] = __ldg(p.gamma + c);
        }
      }
      const RopeRow rr = p.rope_axial ? rope_row(p, mrow) : RopeRow{0u, 0u, false};
      __syncwarp();   // the previous tile's reads of the line are done
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + 4u * (32 * j + lane)), "f"(bv[j]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + 256u + 4u * (32 * j + lane)), "f"(gv[j]) : "memory");
      }
      __syncwarp();
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::tc_fence_after();
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc) * 256u + (static_cast<uint32_t>(lq * 32) << 16);
      // one 4 KB staging box per warp: 32 rows x 128 B, SWIZZLE_128B (16-byte chunk q of row r at q ^ (r & 7)) —
      // 64 bf16 columns (both halves, ONE store) or 32 f32 columns (one store per half)
      const uint32_t rowaddr = stg + static_cast<uint32_t>(lane) * 128u;
      const uint32_t sw = static_cast<uint32_t>(lane) & 7u;
      auto wait_box = [&]() {   // the previous bulk store must have finished READING the box
        if (pending) {
          tc::bulk_wait_read0();
          pending = false;
        }
        __syncwarp();
      };
      auto issue = [&](int col) {
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (mode == kStoreTmaAddF32) tc::tma_reduce_add_2d(&tmap_c, stg, col, row0);
          else tc::tma_store_2d(&tmap_c, stg, col, row0);
          tc::bulk_commit();
          pending = true;
        }
      };
      const int cbase = cg * 64;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = cbase + 32 * j;
        if (c >= p.BN) break;  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld32(t_addr + c, r);
        tc::tmem_ld_wait();
        const bool live = n0 + c < p.N;  // warp-uniform
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (live) epilogue_math_v2(p, v, n0 + c, vec, 32 * j, rope_smem, rr);
        if (mode == kStoreDirect) {
          if (live && row < p.M) epilogue_store_direct<32>(p, v, row, n0 + c);
          continue;
        }
        if (mode == kStoreTmaBf16) {
          if (j == 0) {
            if (!live) break;
            wait_box();
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = live ? tc::pack_bf16(v[2 * i], v[2 * i + 1]) : 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_shared_v4(rowaddr + ((static_cast<uint32_t>(4 * j + q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2],
                         pk[4 * q + 3]);
          // a 64-wide box never reaches into a neighbouring tile: with several n-tiles BN is a multiple of 64
          // (choose_bn2), with one n-tile everything right of N is clipped by the tensor map
          if (j == 1 || c + 32 >= p.BN) issue(n0 + cbase);
        } else {
          if (!live) break;
          wait_box();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            st_shared_v4(rowaddr + ((static_cast<uint32_t>(q) ^ sw) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                         __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
          issue(n0 + c);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc) & kPeerMask);   // the leader's barrier, from either CTA
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (pending) tc::bulk_wait0();  // global writes complete before the CTA retires
  }

  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA of the pair retires while its peer may still address its shared memory / TMEM
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT debug kernel (bring-up cross-check only; selected with args->impl == 1)
// ---------------------------------------------------------------------------------------------
__global__ void gemm_bf16_simt_kernel(const __nv_bfloat16* __restrict__ A, long long lda,
                                      const __nv_bfloat16* __restrict__ W, long long ldw,
                                      const GemmParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  __shared__ float sa[16][17];
  __shared__ float sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty;
  const int col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    sa[ty][tx] = (row < p.M && k0 + tx < p.K) ? __bfloat162float(A[row * lda + k0 + tx]) : 0.f;
    const int wr = blockIdx.x * 16 + ty;
    sb[ty][tx] = (wr < p.N && k0 + tx < p.K) ? __bfloat162float(W[wr * ldw + k0 + tx]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += sa[ty][k] * sb[tx][k];
    __syncthreads();
  }
  // epilogue on column pairs so the rotary path matches the tensor-core kernel
  __shared__ float sc[16][17];
  sc[ty][tx] = acc;
  __syncthreads();
  if ((tx & 1) == 0 && row < p.M && col < p.N) {
    float v[2] = {sc[ty][tx], sc[ty][tx + 1]};
    for (int i = 0; i < 2; ++i) {
      const int c = col + i;
      if (c >= p.N) continue;
      if (p.bias) v[i] += p.bias[c];
      if (p.act == 1) v[i] = fmaxf(v[i], 0.f);
      if (p.act == 2) v[i] = gelu_erf(v[i]);
      if (p.gamma) v[i] *= p.gamma[c];
    }
    if ((p.rope_cs || p.rope_axial) && col >= p.rope_col0 && col < p.rope_col1) {
      const int rb = row % p.rope_rows_per_batch;
      if (rb < p.rope_row_limit) {
        const int pos = rb % p.rope_period;
        const int pair = ((col - p.rope_col0) & 255) >> 1;
        const float2 c = p.rope_axial
            ? reinterpret_cast<const float2*>(p.rope_axial)[(pair & 63) * p.rope_side +
                                                            (pair < 64 ? pos % p.rope_side : pos / p.rope_side)]
            : reinterpret_cast<const float2*>(p.rope_cs)[static_cast<size_t>(pair) * p.rope_period + pos];
        const float a = v[0], b = v[1];
        v[0] = a * c.x - b * c.y;
        v[1] = a * c.y + b * c.x;
      }
    }
    for (int i = 0; i < 2; ++i) {
      const int c = col + i;
      if (c >= p.N) continue;
      if (p.residual) {
        const long long rr = p.res_row_mod > 0 ? (row % p.res_row_mod) : row;
        v[i] += p.residual[rr * p.ldr + c];
      }
      if (p.out_f32) p.out_f32[static_cast<long long>(row) * p.ldc + c] = v[i];
      if (p.out_bf16) p.out_bf16[static_cast<long long>(row) * p.ldc_bf16 + c] = __float2bfloat16(v[i]);
    }
  }
}

// Tile width: one n-tile when N <= 256 (any multiple of 16); otherwise the multiple of 32 that
// minimises (waves x tile width), i.e. the tensor-pipe time of the slowest SM.
static int choose_bn(int M, int N, int sms, int step) {
  if (N <= 256) return ((N + 15) / 16) * 16;
  const int tiles_m = (M + kBM - 1) / kBM;
  int best = 256;
  long long best_cost = -1;
  for (int bn = 256; bn >= 64; bn -= step) {
    const long long tiles = static_cast<long long>(tiles_m) * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    const long long cost = waves * (bn + 24);  // +24: per-tile fixed cost (barriers, accumulator hand-over)
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}


// Tile width of the CTA-pair kernel: multiples of 32 (32-column staging boxes); one n-tile when N <= 256, otherwise the
// width that minimises waves x (width + per-tile fixed cost) over the 74 clusters.
static int choose_bn2(int M, int N, int clusters, int step) {
  if (N <= 256) return ((N + 31) / 32) * 32;
  const int tiles_m = (M + 255) / 256;
  int best = 256;
  long long best_cost = -1;
  for (int bn = 256; bn >= 64; bn -= step) {
    const long long tiles = static_cast<long long>(tiles_m) * ((N + bn - 1) / bn);
    const long long waves = (tiles + clusters - 1) / clusters;
    const long long cost = waves * (bn + 24);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static int launch_gemm2(const ds2_gemm_args* a, GemmParams p, int sms, cudaStream_t st) {
  const int clusters = sms / 2;
  const int step = p.store_mode == kStoreTmaBf16 ? 64 : 32;   // bf16 results leave in 64-column boxes
  p.BN = choose_bn2(a->M, a->N, clusters, step);
  {
    static const int force_bn = [] {
      const char* e = getenv("DS2_GEMM_BN");
      return e ? atoi(e) : 0;
    }();
    if (force_bn >= 64 && force_bn <= 256 && (force_bn % step) == 0 && a->N > 256) p.BN = force_bn;
  }
  p.tiles_m = (a->M + 255) / 256;
  p.tiles_n = (a->N + p.BN - 1) / p.BN;
  p.stages = a->rope_axial ? 3 : k2Stages;    // 64 * side * 8 B <= 32 KB = the fourth stage's slot (side <= 64)
  DS2_REQUIRE(!a->rope_axial || a->rope_side <= 64, DS2_E_ARG, "ds2_gemm: axial rotary table side %d > 64", a->rope_side);
  CUtensorMap ta, tw, tcm;
  memset(&tcm, 0, sizeof(tcm));
  if (p.store_mode == kStoreTmaBf16) {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldc_bf16) * 2};
    const uint32_t box[2] = {64, 32};
    int rc = make_tmap(&tcm, a->out_bf16, 2, 128, 2, dims, strides, box);
    if (rc) return rc;
  } else if (p.store_mode == kStoreTmaF32 || p.store_mode == kStoreTmaAddF32) {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldc) * 4};
    const uint32_t box[2] = {32, 32};
    int rc = make_tmap(&tcm, a->out_f32, 4, 128, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->lda) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(kBM)};
    int rc = make_tmap_bf16(&ta, a->A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldw) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(p.BN / 2)};
    int rc = make_tmap_bf16(&tw, a->W, 2, dims, strides, box);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_tcgen05_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2Smem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = 2 * (tiles < clusters ? tiles : clusters);
  DS2_LAUNCH((gemm2_bf16_tcgen05_2cta_kernel), grid, k2Threads, k2Smem, st, ta, tw, tcm, p);
  return post_launch("gemm2_bf16_tcgen05_2cta_kernel");
}

}  // namespace ds2

extern "C" int ds2_gemm(const ds2_gemm_args* a, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a != nullptr, DS2_E_ARG, "ds2_gemm: null args");
  DS2_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, DS2_E_ARG, "ds2_gemm: bad shape %d %d %d", a->M, a->N,
              a->K);
  DS2_REQUIRE(a->A && a->W, DS2_E_ARG, "ds2_gemm: null operand");
  DS2_REQUIRE(a->out_f32 || a->out_bf16, DS2_E_ARG, "ds2_gemm: no output");
  DS2_REQUIRE(a->lda >= a->K && a->ldw >= a->K, DS2_E_ARG, "ds2_gemm: leading dim < K");
  if (a->rope_cs || a->rope_axial) {
    DS2_REQUIRE(a->rope_period > 0 && a->rope_rows_per_batch > 0 && (a->rope_col0 % 32) == 0 &&
                    (a->rope_col1 % 32) == 0,
                DS2_E_ARG, "ds2_gemm: bad rotary spec");
  }
  if (a->rope_axial) {
    DS2_REQUIRE(a->rope_side > 0 && a->rope_side * a->rope_side == a->rope_period && a->rope_side <= 96 &&
                    (a->rope_side % 2) == 0 && (reinterpret_cast<uintptr_t>(a->rope_axial) & 15) == 0,
                DS2_E_ARG, "ds2_gemm: axial rotary table needs period == side^2, even side <= 96, 16-byte alignment");
  }
  GemmParams p;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.bias = a->bias;
  p.gamma = a->gamma;
  p.residual = a->residual;
  p.ldr = a->ldr;
  p.res_row_mod = a->res_row_mod;
  p.act = a->act;
  p.out_f32 = a->out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.ldc = a->ldc;
  p.ldc_bf16 = a->ldc_bf16;
  p.rope_cs = a->rope_cs;
  p.rope_col0 = a->rope_col0;
  p.rope_col1 = a->rope_col1;
  p.rope_period = a->rope_period;
  p.rope_rows_per_batch = a->rope_rows_per_batch;
  p.rope_row_limit = a->rope_row_limit;
  p.rope_axial = a->rope_axial;
  p.rope_side = a->rope_side;
  p.stages = a->rope_axial ? 3 : kStages;   // 64 * side * 8 B <= 48 KB = the fourth stage's slot
  cudaStream_t st = as_stream(stream);

  if (a->impl == 1) {
    p.BN = 16;
    p.tiles_m = p.tiles_n = 0;
    dim3 grid((a->N + 15) / 16, (a->M + 15) / 16), block(16, 16);
    DS2_LAUNCH((gemm_bf16_simt_kernel), grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(a->A), a->lda,
                                                  reinterpret_cast<const __nv_bfloat16*>(a->W), a->ldw,
                                                  p);
    return post_launch("gemm_bf16_simt_kernel");
  }

  DS2_REQUIRE((a->lda % 8) == 0 && (a->ldw % 8) == 0, DS2_E_ALIGN,
              "ds2_gemm: lda/ldw must be multiples of 8 elements (got %lld, %lld)",
              static_cast<long long>(a->lda), static_cast<long long>(a->ldw));
  const int sms = sm_count();
  DS2_REQUIRE(sms > 0, DS2_E_NODEVICE, "ds2_gemm: no CUDA device");
  // epilogue store path
  p.store_mode = kStoreDirect;
  const bool one_out = (a->out_f32 != nullptr) != (a->out_bf16 != nullptr);
  if (one_out && a->impl != 2) {   // impl 2: tests force the row-per-thread stores
    if (a->out_bf16 && !a->residual && (a->ldc_bf16 % 8) == 0 &&
        (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0) {
      p.store_mode = kStoreTmaBf16;
    } else if (a->out_f32 && (a->ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0) {
      if (!a->residual) p.store_mode = kStoreTmaF32;
      else if (a->residual == a->out_f32 && a->ldr == a->ldc && a->res_row_mod == 0) p.store_mode = kStoreTmaAddF32;
    }
  }

  // ---- CTA-pair kernel (gemm2): the default for tall problems; impl 3 forces it, impl 4 forces the single-CTA kernel ----
  {
    static const int v2_env = [] {
      const char* e = getenv("DS2_GEMM_V2");
      return e ? atoi(e) : 1;
    }();
    const bool full_table = a->rope_cs != nullptr && a->rope_axial == nullptr;   // only the single-CTA kernel reads it
    const bool want_v2 = !full_table && (a->impl == 3 || (a->impl == 0 && v2_env != 0 && a->M >= 512 && a->N >= 32));
    if (want_v2 && (sms % 2) == 0) return launch_gemm2(a, p, sms, st);
  }
  // bf16 stores go out in 64-column boxes, so tiles must not end inside one
  p.BN = choose_bn(a->M, a->N, sms, p.store_mode == kStoreTmaBf16 ? 64 : 32);
  {
    // tuning aid: DS2_GEMM_BN=<n> forces the tile width of every multi-tile problem (N > 256) for A/B sweeps of the
    // cost model in choose_bn; ignored unless n is a legal width for the store path
    static const int force_bn = [] {
      const char* e = getenv("DS2_GEMM_BN");
      return e ? atoi(e) : 0;
    }();
    const int step = p.store_mode == kStoreTmaBf16 ? 64 : 32;
    if (force_bn >= 64 && force_bn <= 256 && (force_bn % step) == 0 && a->N > 256) p.BN = force_bn;
  }
  p.tiles_m = (a->M + kBM - 1) / kBM;
  p.tiles_n = (a->N + p.BN - 1) / p.BN;
  CUtensorMap ta, tw, tcm;
  memset(&tcm, 0, sizeof(tcm));
  if (p.store_mode == kStoreTmaBf16) {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldc_bf16) * 2};
    const uint32_t box[2] = {64, 32};
    int rc = make_tmap(&tcm, a->out_bf16, 2, 128, 2, dims, strides, box);
    if (rc) return rc;
  } else if (p.store_mode == kStoreTmaF32 || p.store_mode == kStoreTmaAddF32) {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldc) * 4};
    const uint32_t box[2] = {32, 32};
    int rc = make_tmap(&tcm, a->out_f32, 4, 128, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->lda) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(kBM)};
    int rc = make_tmap_bf16(&ta, a->A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldw) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(p.BN)};
    int rc = make_tmap_bf16(&tw, a->W, 2, dims, strides, box);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_gemm: cudaFuncSetAttribute: %s",
                cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sms ? tiles : sms;
  DS2_LAUNCH((gemm_bf16_tcgen05_kernel), grid, kGemmThreads, kGemmSmem, st, ta, tw, tcm, p);
  return post_launch("gemm_bf16_tcgen05_kernel");
}
