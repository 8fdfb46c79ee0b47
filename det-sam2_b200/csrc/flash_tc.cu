// flash_tc.cu — single-head (d = 256) flash attention on tcgen05 for SAM 2 memory attention.
//
//   out[b] = softmax(scale * q[b] k[b]^T) v[b]     q [B,Lq,256]  k [B,Lk,256]  v [B,Lk,DV]  (bf16)
//
// One CTA owns NQ query tiles of 128 rows of one object and streams the keys once for all of them.
//   warp 0           : TMA producer — Q tiles once (4 boxes of 128x64 each), then rings of K tiles
//                      (BN keys x 256, K-major) and V tiles (BN keys x DV, MN-major), 128B swizzle.
//                      K slots are released as soon as Q·K^T of that tile retires, V slots after P·V.
//   warp 1           : TMEM allocator + MMA issuer
//                        S_h,j = Q_h K_j^T     tcgen05.mma SS, M=128 N=BN, 16 K-steps -> S[h][j&1]
//                        O_h  += P_h,j V_j     tcgen05.mma TS (A = P in TMEM),  M=128 N=DV
//                      S is double-buffered per query tile so Q·K^T of tile j+1 overlaps softmax j.
//   warps 2..        : softmax, SP warpgroups per query tile.  A row (TMEM lane) of a key tile is split
//                      column-wise over SP threads (one per warpgroup, both on the SMSP that owns the
//                      lane quarter): tcgen05.ld of BN/SP columns, local max, the SP partial maxima of
//                      a row are exchanged through shared memory (bf16, one named barrier per lane
//                      quarter) so all parts use the SAME reference, lazy rescaling (threshold 2^8),
//                      ex2, partial row sum, bf16 P written over S with tcgen05.st; conditional O
//                      rescale and the final O / l epilogue are split over the DV columns.
//                      With one warpgroup the softmax of a 128x128 tile (~2400 clk: 128 dependent
//                      MUFU.EX2 per thread at 4 lanes/clk/SMSP plus load/convert) was twice the MMA time
//                      (1280 clk); two warps per SMSP overlap each other's TMEM loads, barrier waits
//                      and stores with the other's MUFU work.
// TMEM columns: S[h][i] at (2h+i)*BN, O[h] at 2*NQ*BN + h*DV  (<= 512).
// DV = 64 is the cross-attention case: the 64->256 value projection is applied AFTER P·V by the
// caller (softmax rows sum to one), which cuts P·V work and V traffic 4x; NQ = 2 halves the K
// traffic per FLOP (K tiles come out of L2, which is the binding bandwidth here).
// HD = 80 (the kernel keeps its name) is the multi-head variant for Hiera's global-attention blocks: head_dim 72
// padded to 80, H heads side by side in the token-major qkv matrix, optional 16 x 16 window addressing — see FlashCfg.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "tc05.cuh"

namespace ds2 {

constexpr int kHD = 256;
constexpr int kQM = 128;

struct FlashParams {
  int B, Lq, Lk;
  float scale_log2;
  // Two key halves (two_phase != 0): EVERY (query tile, object) item is computed as two independent flash passes over
  // the key tiles [0, T/2) and [T/2, T) whose (O, max, sum) partials go through the workspace `ws` and are combined
  // in part order by the same code — by the one CTA that ran both halves back to back (the first n_full items), or, for
  // the items of the partial last wave, by whichever of the item's two CTAs arrives last (atomic counter per item).
  // The arithmetic of an item is therefore identical however it was scheduled: the grid may split the tail wave (3.46
  // waves of 148 CTAs at 16 objects become 3 + 0.5) and "an object tracked in a batch == tracked alone" still holds
  // bit for bit.  two_phase == 0: one pass over all keys, no workspace (self-attention, short key ranges).
  int n_full, two_phase, q_tiles;
  float* ws;               // [items][2][128 rows][DV + 4] f32
  unsigned int* ws_count;  // [items], zero between launches
  int dbg;                 // timing experiments only (impl 5/6/8/13): 1 = load half of each K tile, 2 = skip the exps, 4 = softmax relays only, 3 = stall
                           // accounting.  (Two more — no TMEM loads of S, softmax warps relaying barriers only — were run
                           // once, profiles/r2_s7_flash_softmax_stage_accounting.txt, and removed: they cost registers.)
  const __nv_bfloat16* q;  // QT kernels read the query rows straight from global memory
  long long ldq, bsq;
  int H, D;                // HD == 80 kernels (Hiera global attention): heads per batch item, head h at column h * D
  // HD == 80, window mode (win_nwin > 0): an item is one 128-row half of a 16 x 16 window of the [Hm, Wm] token raster
  // (Lq = Lk = 256); K / V tiles are boxes {64 columns, 16 x, 8 y} of a rank-4 map {columns, x, y, batch}
  int win_nwin, win_nwx, Wm;
  __nv_bfloat16* out;
  long long ldo, bso;
};

// IL = softmax warpgroups that ALTERNATE over the key tiles of a query tile (tile j belongs to group j % IL), each with
// its own O accumulator and running (max, sum): the per-tile softmax latency of one warp per scheduler (TMEM load, row
// max, 128 dependent-issue exps per thread: ~1.7k clk against ~1.5k clk of MMA work per tile) sat on the same-S-buffer
// chain softmax(j) -> P.V(j) -> Q.K^T(j+2); with two groups the softmax of tile j+1 runs while P.V(j) and Q.K^T(j+2) wait
// for tile j.  The groups' partial results are combined like the two key halves (fixed order, through the workspace).
// TP = 1: the instantiation that runs every item as two key halves through the workspace (see FlashParams); the plain
// instantiation (TP = 0) contains none of that code — it costs the single-pass kernel ~120 registers otherwise.
// HD = 80: the multi-head variant for Hiera's global-attention blocks (hieradet.py:57-82 with window_size == 0): head_dim
// 72 zero-padded to 80 = 5 K-steps, DV = 80, one CTA per (query tile, head, frame).  K and V rows of a head are 144-byte
// segments of the token-major qkv matrix, so a tile is fetched as TWO 64-column TMA boxes: columns [h*D, h*D+64) and
// [h*D+64, h*D+128) — the second box carries 8 real dims and, beyond them, the next head's data (or TMA zero fill past the
// end of the row).  That tail is harmless: Q's columns D..79 are stored as zeros (so it adds nothing to Q.K^T), and the
// output columns D..79 it produces in P.V are never written out.
template <int DV, int BN, int NQ, int KS, int VS, int SP, int QT, int IL = 1, int TP = 0, int HD = 256>
struct FlashCfg {
  static constexpr int kKBox = (HD + 63) / 64;    // 64-column boxes per K tile
  static constexpr int kVBox = (DV + 63) / 64;
  static constexpr int kQCols = HD == 256 ? 128 : 48;   // TMEM columns of a query tile (bf16 pairs; 40 used + 8 zero at HD = 80)
  // IL == 2: three whole warpgroups — warps 0-3 control (TMA producer, MMA issuer, two idle), warps 4-7 / 8-11 the two
  // softmax groups — so that setmaxnreg can move registers from the control warpgroup to the softmax warpgroups
  static constexpr int kSoftmaxWarp0 = IL == 2 ? 4 : 2;
  static constexpr int kThreads = 32 * kSoftmaxWarp0 + 128 * NQ * SP * IL;
  static constexpr int kXchBytes = 2 * SP * NQ * kQM * 2;  // [parity][part][row] bf16 partial maxima
  static constexpr int kQBytes = QT ? 0 : kQM * kKBox * 128;  // 64 KB per query tile when Q is a shared-memory operand
  static constexpr int kKBytes = BN * kKBox * 128;
  static constexpr int kVBytes = BN * kVBox * 128;
  static constexpr int kSmemData = NQ * kQBytes + KS * kKBytes + VS * kVBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmem = kSmemData + 1024 + kBarBytes + (kXchBytes < 1024 ? 1024 : kXchBytes);
  static_assert(BN % (32 * SP) == 0 && (DV % (32 * SP) == 0 || DV == 80), "column split");
  static constexpr int kTmemCols = 2 * NQ * BN + NQ * IL * DV + (QT ? NQ * kQCols : 0);
  static_assert(HD == 256 || (HD == 80 && DV == 80 && QT == 1 && NQ == 1 && TP == 0 && SP <= 2 && IL == 1), "multi-head variant");
  static_assert(IL == 1 || (IL == 2 && SP == 1 && NQ == 1), "alternating softmax groups: one query tile, unsplit rows");
  static_assert(TP == 0 || (SP == 1 && NQ == 1), "two key halves: one query tile, unsplit rows");
  static_assert(kTmemCols <= 512, "TMEM budget");
  static_assert(kSmem <= 232448, "shared memory budget");
};

// Stall accounting for tuning (impl == 8 only): cycles the MMA issuer spent blocked on each barrier class
// and the softmax warps on theirs, summed over CTAs.  [0] kfull [1] vfull [2] pfull [3] total MMA-warp
// cycles, [4] sfull (softmax warp 2) [5] odone [6] total softmax-warp cycles [7] CTAs.
// ([8..15] unused: the per-stage accounting of the softmax loop — S load, row max, exp + sum + pack, P store, arrive — cost
// the product kernel 118 registers per thread and was removed after profiles/r2_s7_flash_softmax_stage_accounting.txt)
__device__ unsigned long long g_flash_stall[16];

__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (!on) {
    tc::mbar_wait(bar, parity);
    return;
  }
  const long long t = clock64();
  tc::mbar_wait(bar, parity);
  acc += clock64() - t;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DV, int BN, int NQ, int KS, int VS, int SP, int QT, int IL = 1, int TP = 0, int HD = 256>
__global__ void __launch_bounds__((IL == 2 ? 128 : 64) + 128 * NQ * SP * IL, 1)
flash_d256_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q,
                          const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const FlashParams p) {
  using Cfg = FlashCfg<DV, BN, NQ, KS, VS, SP, QT, IL, TP, HD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sq = smem_base;
  const uint32_t sk0 = sq + NQ * Cfg::kQBytes;
  const uint32_t sv0 = sk0 + KS * Cfg::kKBytes;
  const uint32_t bar_base = sv0 + VS * Cfg::kVBytes;
  int nb = 0;
  const int o_q = nb;        nb += NQ;
  const int o_kfull = nb;    nb += KS;
  const int o_kempty = nb;   nb += KS;
  const int o_vfull = nb;    nb += VS;
  const int o_vempty = nb;   nb += VS;
  const int o_sfull = nb;    nb += 2 * NQ;
  const int o_pfull = nb;    nb += 2 * NQ;
  const int o_odone = nb;    nb += NQ * IL;
  auto bar = [&](int off, int i) { return bar_base + 8u * static_cast<uint32_t>(off + i); };
  const uint32_t tmem_slot = bar_base + 8u * static_cast<uint32_t>(nb);
  const uint32_t xch_base = bar_base + Cfg::kBarBytes;  // partial row maxima / row sums of the split softmax

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int item = blockIdx.x, kv_part = 0, kv_parts = 1;
  if (item >= p.n_full) {          // only with two_phase: an item of the last wave, one CTA per key half
    const int t = item - p.n_full;
    item = p.n_full + (t >> 1);
    kv_part = t & 1;
    kv_parts = 2;
  }
  const int q0 = (item % p.q_tiles) * (kQM * NQ);
  const int hcol = HD == 256 ? 0 : ((item / p.q_tiles) % p.H) * p.D;   // first column of this CTA's head
  const bool win = HD != 256 && p.win_nwin > 0;
  int b = HD == 256 ? item / p.q_tiles : (item / p.q_tiles) / p.H;
  int wx16 = 0, wy16 = 0;               // first raster column / row of the window
  if (win) {
    const int w = b % p.win_nwin;
    b /= p.win_nwin;
    wy16 = (w / p.win_nwx) * 16;
    wx16 = (w % p.win_nwx) * 16;
  }
  // token (row of the q / out matrices inside batch item b) of query row `r` of this item's sequence
  auto tok_of = [&](int r) -> long long {
    return win ? static_cast<long long>(wy16 + (r >> 4)) * p.Wm + wx16 + (r & 15) : static_cast<long long>(r);
  };
  const int all_tiles = (p.Lk + BN - 1) / BN;
  const int j0 = (all_tiles * kv_part) / kv_parts;                     // first key tile of this CTA
  const int n_tiles = (all_tiles * (kv_part + 1)) / kv_parts - j0;    // its number of key tiles (>= 1)
  // local tile index at which the second key half starts when this CTA runs both halves itself (else never reached)
  const int split_at = (TP && kv_parts == 1) ? all_tiles / 2 : -1;
  // tile j starts a fresh accumulation of its softmax group: the group's first tile of the item or of the second half
  auto starts_group = [&](int j) { return j < IL || (TP && split_at >= 0 && j >= split_at && j - IL < split_at); };

  if (warp == 0 && lane == 0) {
    if (!QT) tc::prefetch_tmap(&tmap_q);
    tc::prefetch_tmap(&tmap_k);
    tc::prefetch_tmap(&tmap_v);
    for (int i = 0; i < NQ; ++i) tc::mbar_init(bar(o_q, i), QT ? (HD == 256 ? 4 * SP : 4) : 1);
    for (int i = 0; i < KS; ++i) {
      tc::mbar_init(bar(o_kfull, i), 1);
      tc::mbar_init(bar(o_kempty, i), 1);
    }
    for (int i = 0; i < VS; ++i) {
      tc::mbar_init(bar(o_vfull, i), 1);
      tc::mbar_init(bar(o_vempty, i), 1);
    }
    for (int i = 0; i < 2 * NQ; ++i) {
      tc::mbar_init(bar(o_sfull, i), 1);
      tc::mbar_init(bar(o_pfull, i), 4 * SP);
    }
    for (int i = 0; i < NQ * IL; ++i) tc::mbar_init(bar(o_odone, i), 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
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
  auto tmem_s = [&](int h, int i) { return tmem_base + static_cast<uint32_t>((2 * h + i) * BN); };
  auto tmem_o = [&](int h, int g) { return tmem_base + static_cast<uint32_t>(2 * NQ * BN + (h * IL + g) * DV); };
  // QT: the query tile lives in TMEM (128 lanes x 128 columns of bf16 pairs) and Q·K^T is a TS MMA, so
  // the tensor core reads only the K tile from shared memory.  With both operands in shared memory a
  // 128x128x16 MMA reads 8 KB per 64 clk = the whole 128 B/clk of the SM's shared memory, on top of the
  // TMA writes of the K/V rings — the kernel was shared-memory-bandwidth bound at ~52 % of the MMA rate.
  auto tmem_q = [&](int h) { return tmem_base + static_cast<uint32_t>(2 * NQ * BN + NQ * IL * DV + h * Cfg::kQCols); };

  // Role gates use elect.sync, not `lane == 0`: tcgen05.mma / TMA take their operands from the uniform
  // datapath, and under a lane-id predicate the compiler wraps EVERY such instruction in an
  // ELECT / BRA.U.ANY serialisation loop (~97 clk per MMA issued, measured with tools/mma_rate.cu, against
  // 42-74 clk for the instruction itself) — the single issuing thread, not the tensor pipe, set the pace.
  constexpr int kSW0 = Cfg::kSoftmaxWarp0;
  if (warp < kSW0) {
  if constexpr (IL == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // control warpgroup: 168 -> 40 registers (40 * 128 + 232 * 256 = the 168 * 384 the CTA was launched with)
  if (warp == 0 && tc::elect_one()) {
    // ------------------------------ TMA producer ------------------------------
    if (!QT) {
      for (int h = 0; h < NQ; ++h) {
        tc::mbar_expect_tx(bar(o_q, h), Cfg::kQBytes);
        for (int kk = 0; kk < Cfg::kKBox; ++kk)
          tc::tma_load_3d(sq + h * Cfg::kQBytes + kk * (kQM * 128), &tmap_q, bar(o_q, h), kk * 64,
                          q0 + h * kQM, b);
      }
    }
    for (int j = 0; j < n_tiles; ++j) {
      {
        const int s = j % KS;
        tc::mbar_wait(bar(o_kempty, s), ((j / KS) & 1) ^ 1);
        const int nkk = (p.dbg == 1) ? Cfg::kKBox / 2 : Cfg::kKBox;   // (timing experiment: half of the boxes)
        tc::mbar_expect_tx(bar(o_kfull, s), nkk * (BN * 128));
        const uint32_t sk = sk0 + s * Cfg::kKBytes;
        for (int kk = 0; kk < nkk; ++kk) {
          if (win) tc::tma_load_4d(sk + kk * (BN * 128), &tmap_k, bar(o_kfull, s), hcol + kk * 64, wx16, wy16 + (j0 + j) * (BN / 16), b);
          else tc::tma_load_3d(sk + kk * (BN * 128), &tmap_k, bar(o_kfull, s), hcol + kk * 64, (j0 + j) * BN, b);
        }
      }
      {
        const int s = j % VS;
        tc::mbar_wait(bar(o_vempty, s), ((j / VS) & 1) ^ 1);
        const int nvb = (HD != 256 && p.dbg == 1) ? 1 : Cfg::kVBox;
        tc::mbar_expect_tx(bar(o_vfull, s), nvb * (BN * 128));
        const uint32_t sv = sv0 + s * Cfg::kVBytes;
        for (int nn = 0; nn < nvb; ++nn) {
          if (win) tc::tma_load_4d(sv + nn * (BN * 128), &tmap_v, bar(o_vfull, s), hcol + nn * 64, wx16, wy16 + (j0 + j) * (BN / 16), b);
          else tc::tma_load_3d(sv + nn * (BN * 128), &tmap_v, bar(o_vfull, s), hcol + nn * 64, (j0 + j) * BN, b);
        }
      }
    }
  } else if (warp == 1 && tc::elect_one()) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc_qk = tc::make_idesc_bf16(kQM, BN, 0, 0);
    const uint32_t idesc_pv = tc::make_idesc_bf16(kQM, DV, 0, 1);
    const bool prof = IL == 1 && p.dbg == 3;   // (the stall counters do not fit the 48 registers of the IL == 2 control warps)
    long long w_k = 0, w_v = 0, w_p = 0;
    const long long t_begin = clock64();
    auto issue_qk = [&](int j) {
      const int s = j % KS;
      timed_wait(bar(o_kfull, s), (j / KS) & 1, prof, w_k);
      tc::tc_fence_after();
      const uint32_t sk = sk0 + s * Cfg::kKBytes;
#pragma unroll
      for (int h = 0; h < NQ; ++h) {
        const uint32_t d = tmem_s(h, j & 1);
        const uint32_t sqh = sq + h * Cfg::kQBytes;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t db = tc::make_desc_sw128(sk + (k >> 2) * (BN * 128) + (k & 3) * 32, 16, 1024);
          if (QT) {
            tc::umma_ts(d, tmem_q(h) + k * 8, db, idesc_qk, k != 0 ? 1u : 0u);
          } else {
            const uint64_t da = tc::make_desc_sw128(sqh + (k >> 2) * (kQM * 128) + (k & 3) * 32, 16, 1024);
            tc::umma_ss(d, da, db, idesc_qk, k != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(bar(o_sfull, 2 * h + (j & 1)));
      }
      tc::umma_commit(bar(o_kempty, s));
    };
    for (int h = 0; h < NQ; ++h) tc::mbar_wait(bar(o_q, h), 0);
    tc::tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int s = j % VS;
      timed_wait(bar(o_vfull, s), (j / VS) & 1, prof, w_v);
      const uint32_t sv = sv0 + s * Cfg::kVBytes;
#pragma unroll
      for (int h = 0; h < NQ; ++h) {
        timed_wait(bar(o_pfull, 2 * h + (j & 1)), (j >> 1) & 1, prof, w_p);
        tc::tc_fence_after();
        const uint32_t pa = tmem_s(h, j & 1);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // V tile is MN-major: 64-channel blocks BN*128 B apart (LBO), 8-key groups 1024 B apart (SBO)
          const uint64_t db = tc::make_desc_sw128(sv + k * 2048, BN * 128, 1024);
          // O restarts from zero at the first tile of each key half (the softmax warps have moved the first half's
          // O out of TMEM before they released P of this tile, see below)
          tc::umma_ts(tmem_o(h, j % IL), pa + k * 8, db, idesc_pv, (!starts_group(j) || k != 0) ? 1u : 0u);
        }
        tc::umma_commit(bar(o_odone, h * IL + j % IL));
      }
      tc::umma_commit(bar(o_vempty, s));
    }
    if (prof) {
      atomicAdd(&g_flash_stall[0], static_cast<unsigned long long>(w_k));
      atomicAdd(&g_flash_stall[1], static_cast<unsigned long long>(w_v));
      atomicAdd(&g_flash_stall[2], static_cast<unsigned long long>(w_p));
      atomicAdd(&g_flash_stall[3], static_cast<unsigned long long>(clock64() - t_begin));
      atomicAdd(&g_flash_stall[7], 1ull);
    }
  }
  } else {
    // ------------------------------ softmax / epilogue ------------------------------
    if constexpr (IL == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");  // softmax warpgroups: 168 -> 232 registers
    constexpr int CW = BN / SP;       // S columns of a row handled by this thread
    constexpr int OW = DV / SP;       // O columns of a row handled by this thread (rescale, epilogue)
    const int sw = warp - kSW0;
    const int h = (sw / (4 * SP)) % NQ;   // query tile
    const int g = sw / (4 * SP * NQ);     // softmax group: owns the key tiles j with j % IL == g
    const int part = (sw >> 2) % SP;      // which column part of the row
    const int lq = warp & 3;          // TMEM lane quarter this warp may access
    const uint32_t lane_off = static_cast<uint32_t>(lq * 32) << 16;
    const int rloc = h * kQM + lq * 32 + lane;  // row inside the CTA
    const int row = q0 + rloc;
    // O columns of a row owned by this thread.  DV % (32 SP) == 0: [part * OW, +OW) as 32-column chunks.  DV == 80 (kOdd):
    // the chunks at 0 and 32 plus the 16-column chunk at 64 for one thread per row; split over two threads, part 0 owns the
    // chunks at 0 and 64, part 1 the chunk at 32 (tcgen05.ld / st are warp-wide: `part` is uniform in a warp).
    constexpr bool kOdd = DV == 80;
    const uint32_t to = tmem_o(h, g) + lane_off + (kOdd ? (SP == 1 ? 0 : part * 32) : part * OW);
    constexpr int N32 = kOdd ? (SP == 1 ? 2 : 1) : OW / 32;
    const bool has16 = kOdd && part == 0;
    const uint32_t to16 = tmem_o(h, g) + lane_off + 64;
    const int bar_id = 1 + h * 4 + lq;  // named barrier of the SP warps that share these 32 rows
    auto xch_max = [&](int par, int pt) {
      return xch_base + 2u * static_cast<uint32_t>((par * SP + pt) * (NQ * kQM) + rloc);
    };
    if (QT && g == 0 && HD != 256) {
      // multi-head variant: one thread per row (part 0) moves the D real columns of its query row, zero-padded to 96
      if (part == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<long long>(b) * p.bsq + tok_of(row) * p.ldq + hcol);
        const int nch = p.D >> 3;
        uint32_t w0[32], w1[16];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (row < p.Lq && i < nch) t = __ldg(src + i);
          if (i < 8) {
            w0[4 * i] = t.x;
            w0[4 * i + 1] = t.y;
            w0[4 * i + 2] = t.z;
            w0[4 * i + 3] = t.w;
          } else {
            w1[4 * (i - 8)] = t.x;
            w1[4 * (i - 8) + 1] = t.y;
            w1[4 * (i - 8) + 2] = t.z;
            w1[4 * (i - 8) + 3] = t.w;
          }
        }
        const uint32_t tq = tmem_q(h) + lane_off;
        tc::tmem_st32(tq, w0);
        tc::tmem_st16(tq + 32, w1);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar(o_q, h));
      }
    } else if (QT && g == 0) {
      // this thread's part of its query row: global -> registers -> TMEM (bf16 pairs, K-major A operand)
      constexpr int QW = (kHD / 2) / SP;  // 32-bit columns per thread
      const uint4* src = reinterpret_cast<const uint4*>(p.q + static_cast<long long>(b) * p.bsq +
                                                        static_cast<long long>(row) * p.ldq + part * (kHD / SP));
      const uint32_t tq = tmem_q(h) + lane_off + part * QW;
#pragma unroll
      for (int c = 0; c < QW / 32; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (row < p.Lq) t = __ldg(src + c * 8 + i);
          w[4 * i] = t.x;
          w[4 * i + 1] = t.y;
          w[4 * i + 2] = t.z;
          w[4 * i + 3] = t.w;
        }
        tc::tmem_st32(tq + c * 32, w);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar(o_q, h));
    }
    float m_ref = -INFINITY;
    float l = 0.f;
    float m_half0 = 0.f, l_half0 = 0.f;   // running max / sum of the first key half (two-phase items)
    constexpr int WS_ROW = DV + 4;         // workspace row: O row + (max, sum), padded to keep rows 16-byte aligned
    // leaves (O, m, l) of key half `part_idx` of this item in the workspace
    constexpr int kParts = 2 * IL;         // partial results per item: key halves x softmax groups
    auto dump_part = [&](int part_idx, float m_scaled, float lsum) {
      if constexpr (!TP) return;
      float* wrow = p.ws + (static_cast<size_t>(item * kParts + part_idx) * (NQ * kQM) + rloc) * WS_ROW;
#pragma unroll
      for (int c = 0; c < OW / 32; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(to + c * 32, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<uint4*>(wrow + part * OW + c * 32)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
      if (part == 0) {
        wrow[DV] = m_scaled;
        wrow[DV + 1] = lsum;
      }
    };
    const bool prof = IL == 1 && p.dbg == 3 && warp == kSW0;
    long long w_s = 0, w_o = 0;
    const long long t_begin = clock64();
    for (int j = g; j < n_tiles; j += IL) {
      timed_wait(bar(o_sfull, 2 * h + (j & 1)), (j >> 1) & 1, prof, w_s);
      tc::tc_fence_after();
      const uint32_t ts = tmem_s(h, j & 1) + lane_off;
      if (p.dbg == 4) {
        // timing experiment (impl 13): the softmax warps only relay the barriers, so the launch time is that of the
        // TMA + MMA pipeline alone (the output is garbage)
        if (j >= IL) tc::mbar_wait(bar(o_odone, h * IL + g), ((j / IL) - 1) & 1);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar(o_pfull, 2 * h + (j & 1)));
        continue;
      }
      float alpha = 1.f;
      bool resc = false;
      bool half1_flag = false;
      {
        uint32_t sraw[CW];
  #pragma unroll
        for (int c = 0; c < CW / 32; ++c) {
          uint32_t(&chunk)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sraw[c * 32]);
          tc::tmem_ld32(ts + part * CW + c * 32, chunk);
        }
        tc::tmem_ld_wait();
        const int valid = p.Lk - (j0 + j) * BN - part * CW;  // columns >= valid are TMA zero fill -> mask
        if (valid < CW) {
  #pragma unroll
          for (int i = 0; i < CW; ++i)
            if (i >= valid) sraw[i] = 0xff800000u;  // -inf
        }
        // four independent chains: the running max is not one CW-long dependency chain
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  #pragma unroll
        for (int i = 0; i < CW; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(sraw[i]));
        float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (SP > 1) {
          // every part of a row must scale by the same reference: exchange the (bf16-rounded) partial
          // maxima; the barrier also orders "all parts have loaded their S columns" before any part
          // overwrites S with P below
          const __nv_bfloat16 mine = __float2bfloat16_rn(mt);
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(xch_max(j & 1, part)), "h"(__bfloat16_as_ushort(mine)) : "memory");
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * SP) : "memory");
          mt = __bfloat162float(mine);
  #pragma unroll
          for (int o = 0; o < SP; ++o) {
            if (o == part) continue;
            unsigned short u;
            asm volatile("ld.shared.b16 %0, [%1];" : "=h"(u) : "r"(xch_max(j & 1, o)) : "memory");
            mt = fmaxf(mt, __bfloat162float(__ushort_as_bfloat16(u)));
          }
        }
        const bool half1_start = TP && split_at >= 0 && j >= split_at && j - IL < split_at;
        if (j < IL) {
          m_ref = mt;
        } else if (half1_start) {
          // second key half: an independent flash pass — remember the first half's statistics, start over
          m_half0 = m_ref * p.scale_log2;
          l_half0 = l;
          m_ref = mt;
          l = 0.f;
        } else if ((mt - m_ref) * p.scale_log2 > 8.0f) {
          alpha = ex2_approx((m_ref - mt) * p.scale_log2);
          m_ref = mt;
          resc = true;
        }
        const float moff = m_ref * p.scale_log2;
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pk[CW / 2];
  #pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          float e0 = fmaf(__uint_as_float(sraw[2 * i]), p.scale_log2, -moff);
          float e1 = fmaf(__uint_as_float(sraw[2 * i + 1]), p.scale_log2, -moff);
          if (p.dbg != 2) {
            e0 = ex2_approx(e0);
            e1 = ex2_approx(e1);
          }
          sum0 += e0;
          sum1 += e1;
          pk[i] = tc::pack_bf16(e0, e1);
        }
        l = l * alpha + (sum0 + sum1);
        // P (bf16 pairs) overwrites the first BN/2 columns of S: this thread's part at [part*CW/2, +CW/2)
        if (CW / 2 >= 32) {
  #pragma unroll
          for (int c = 0; c < CW / 64; ++c) {
            const uint32_t(&chunk)[32] = *reinterpret_cast<const uint32_t(*)[32]>(&pk[c * 32]);
            tc::tmem_st32(ts + part * (CW / 2) + c * 32, chunk);
          }
        } else {
          const uint32_t(&chunk)[16] = *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]);
          tc::tmem_st16(ts + part * (CW / 2), chunk);
        }
        half1_flag = half1_start;
      }
      // Consume every phase of `odone` (P·V of tile j-1 complete) so this waiter is never more than
      // one phase behind the barrier — parity waits alias otherwise.  By now that MMA has long retired.
      if (j >= IL) timed_wait(bar(o_odone, h * IL + g), ((j / IL) - 1) & 1, prof, w_o);
      if (TP && half1_flag) {
        // P·V of the last tile of the first half is complete and P·V of this tile cannot be issued before every
        // softmax warp has arrived on `pfull` below: O is stable — move the first half's result to the workspace
        tc::tc_fence_after();
        dump_part(g, m_half0, l_half0);   // first half, group g (the half starts at tile 0: label = g)
      }
      if (__any_sync(0xffffffffu, resc)) {
        // O is stable here: P·V of tile j-1 is complete and P·V of tile j is not yet issued
        tc::tc_fence_after();
#pragma unroll
        for (int c = 0; c < N32; ++c) {
          uint32_t o[32];
          tc::tmem_ld32(to + c * 32, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tc::tmem_st32(to + c * 32, o);
        }
        if constexpr (kOdd) {
          if (has16) {
            uint32_t o[16];
            tc::tmem_ld16(to16, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st16(to16, o);
          }
        }
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar(o_pfull, 2 * h + (j & 1)));
    }
    if (prof && lane == 0) {
      atomicAdd(&g_flash_stall[4], static_cast<unsigned long long>(w_s));
      atomicAdd(&g_flash_stall[5], static_cast<unsigned long long>(w_o));
      atomicAdd(&g_flash_stall[6], static_cast<unsigned long long>(clock64() - t_begin));
    }
    if (SP > 1) {
      // total row sum = sum of the parts (same reference in all of them); reuse the exchange buffer as f32
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * SP) : "memory");
      const uint32_t xl = xch_base + 4u * static_cast<uint32_t>(part * (NQ * kQM) + rloc);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xl), "f"(l) : "memory");
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * SP) : "memory");
      float lt = 0.f;
#pragma unroll
      for (int o = 0; o < SP; ++o) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(xch_base + 4u * static_cast<uint32_t>(o * (NQ * kQM) + rloc)) : "memory");
        lt += v;
      }
      l = lt;
    }
    // epilogue: O / l -> bf16 -> global
    const bool has_tiles = g < n_tiles;
    if (has_tiles) {
      const int jl = g + ((n_tiles - 1 - g) / IL) * IL;   // this group's last tile (two-phase halves have >= 8 tiles)
      tc::mbar_wait(bar(o_odone, h * IL + g), (jl / IL) & 1);
    }
    tc::tc_fence_after();
    __nv_bfloat16* orow = p.out + static_cast<long long>(b) * p.bso + (kOdd ? tok_of(row) : static_cast<long long>(row)) * p.ldo +
                          (kOdd ? hcol + (SP == 1 ? 0 : part * 32) : part * OW);
    if constexpr (TP) {
      // ---- two key halves: this CTA's (last) half goes to the workspace, then the halves are combined in part
      //      order by this CTA (it ran both) or by the item's CTA that arrives last ----
      const int slot = item;
      // label of this group inside the half it just finished: tiles are numbered from the START of the half, so that a
      // half computed by its own CTA (tiles from 0) and the same half inside a whole-item CTA (tiles from split_at) put
      // the same tiles under the same label
      const int half = kv_parts == 1 ? 1 : kv_part;
      const int label = kv_parts == 1 ? (((g - split_at) % IL) + IL) % IL : g;
      dump_part(half * IL + label, m_ref * p.scale_log2, l);
      __threadfence();
      asm volatile("bar.sync 15, %0;" ::"n"(128 * NQ * SP * IL) : "memory");
      unsigned int last = 1u;
      if (kv_parts > 1) {
        const uint32_t flag = xch_base;  // the exchange buffer is idle now
        if (warp == kSW0 && lane == 0) {
          const unsigned int old = atomicAdd(p.ws_count + slot, 1u);
          const unsigned int lst = (old == 1u) ? 1u : 0u;
          if (lst) p.ws_count[slot] = 0u;  // both halves have arrived: ready for the next launch
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(flag), "r"(lst) : "memory");
        }
        asm volatile("bar.sync 15, %0;" ::"n"(128 * NQ * SP * IL) : "memory");
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(flag) : "memory");
      }
      if (last) {
        __threadfence();
        // this thread combines OWc of the row's DV columns: the SP x IL threads of a row split them
        constexpr int OWc = DV / (SP * IL);
        const int col0 = (part * IL + g) * OWc;
        __nv_bfloat16* ocomb = p.out + static_cast<long long>(b) * p.bso + static_cast<long long>(row) * p.ldo + col0;
        const float* base = p.ws + (static_cast<size_t>(slot * kParts) * (NQ * kQM) + rloc) * WS_ROW;
        const size_t pstride = static_cast<size_t>(NQ * kQM) * WS_ROW;
        float mmax = -INFINITY;
        for (int q = 0; q < kParts; ++q) mmax = fmaxf(mmax, __ldcg(base + q * pstride + DV));
        float lt = 0.f;
        for (int q = 0; q < kParts; ++q) lt += __ldcg(base + q * pstride + DV + 1) * ex2_approx(__ldcg(base + q * pstride + DV) - mmax);
        const float inv = 1.0f / lt;
#pragma unroll
        for (int c = 0; c < OWc / 32; ++c) {
          float acc[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = 0.f;
          for (int q = 0; q < kParts; ++q) {  // fixed part order: the result does not depend on arrival order
            const float w = ex2_approx(__ldcg(base + q * pstride + DV) - mmax);
            const float4* src = reinterpret_cast<const float4*>(base + q * pstride + col0 + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 v4 = __ldcg(src + i);
              acc[4 * i] = fmaf(v4.x, w, acc[4 * i]);
              acc[4 * i + 1] = fmaf(v4.y, w, acc[4 * i + 1]);
              acc[4 * i + 2] = fmaf(v4.z, w, acc[4 * i + 2]);
              acc[4 * i + 3] = fmaf(v4.w, w, acc[4 * i + 3]);
            }
          }
          if (row < p.Lq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 t;
              t.x = tc::pack_bf16(acc[8 * i] * inv, acc[8 * i + 1] * inv);
              t.y = tc::pack_bf16(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
              t.z = tc::pack_bf16(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
              t.w = tc::pack_bf16(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
              reinterpret_cast<uint4*>(ocomb + c * 32)[i] = t;
            }
          }
        }
      }
    } else if constexpr (IL == 2) {
      // ---- two alternating softmax groups, merged inside the CTA: group 1 leaves its (O, max, sum) in the first K stage
      //      (every Q.K^T has retired: the last P.V of either group was committed after it), group 0 folds it into its own
      //      in a fixed order and writes the output ----
      float* xrow = reinterpret_cast<float*>(smem_raw + (sk0 - tc::smem_u32(smem_raw))) + rloc * WS_ROW;
      if (g == 1) {
#pragma unroll
        for (int c = 0; c < OW / 32; ++c) {
          uint32_t o[32];
          tc::tmem_ld32(to + c * 32, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<uint4*>(xrow + c * 32)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
        xrow[DV] = has_tiles ? m_ref * p.scale_log2 : -INFINITY;
        xrow[DV + 1] = has_tiles ? l : 0.f;
      }
      asm volatile("bar.sync 14, 256;" ::: "memory");
      if (g == 0) {
        const float m0 = m_ref * p.scale_log2, m1 = xrow[DV], l1 = xrow[DV + 1];
        const float mm = fmaxf(m0, m1);
        const float w0 = ex2_approx(m0 - mm);
        const float w1 = l1 > 0.f ? ex2_approx(m1 - mm) : 0.f;
        const float inv = 1.0f / fmaf(l, w0, l1 * w1);
#pragma unroll
        for (int c = 0; c < OW / 32; ++c) {
          uint32_t o[32];
          tc::tmem_ld32(to + c * 32, o);
          tc::tmem_ld_wait();
          if (row < p.Lq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = reinterpret_cast<const float4*>(xrow + c * 32)[2 * i];
              const float4 b4 = reinterpret_cast<const float4*>(xrow + c * 32)[2 * i + 1];
              const float x1[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
              float r[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                r[e] = (__uint_as_float(o[8 * i + e]) * w0 + (l1 > 0.f ? x1[e] * w1 : 0.f)) * inv;
              uint4 t;
              t.x = tc::pack_bf16(r[0], r[1]);
              t.y = tc::pack_bf16(r[2], r[3]);
              t.z = tc::pack_bf16(r[4], r[5]);
              t.w = tc::pack_bf16(r[6], r[7]);
              reinterpret_cast<uint4*>(orow + c * 32)[i] = t;
            }
          }
        }
      }
    } else if constexpr (kOdd) {
      // ---- multi-head variant: only the D real columns of the head leave the CTA ----
      const float inv = 1.0f / l;
      const int nch = p.D >> 3;                              // 8-column groups of the head (9 at D = 72)
      const int g0 = SP == 1 ? 0 : part * 4;                 // first group of this thread's 32-column chunks
#pragma unroll
      for (int c = 0; c < N32; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(to + c * 32, o);
        tc::tmem_ld_wait();
        if (row < p.Lq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (g0 + 4 * c + i < nch) {
              uint4 t;
              t.x = tc::pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
              t.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
              t.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
              t.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
              reinterpret_cast<uint4*>(orow + c * 32)[i] = t;
            }
          }
        }
      }
      if (has16) {
        uint32_t o[16];
        tc::tmem_ld16(to16, o);
        tc::tmem_ld_wait();
        if (row < p.Lq) {
          __nv_bfloat16* o64 = p.out + static_cast<long long>(b) * p.bso + tok_of(row) * p.ldo + hcol + 64;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (8 + i < nch) {
              uint4 t;
              t.x = tc::pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
              t.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
              t.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
              t.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
              reinterpret_cast<uint4*>(o64)[i] = t;
            }
          }
        }
      }
    } else {
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < OW / 32; ++c) {
      uint32_t o[32];
      tc::tmem_ld32(to + c * 32, o);
      tc::tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 t;
          t.x = tc::pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          t.y = tc::pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          t.z = tc::pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          t.w = tc::pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          reinterpret_cast<uint4*>(orow + c * 32)[i] = t;
        }
      }
    }
    }  // unsplit item
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT debug kernel (impl == 1): one warp per query row, f32 math.  Bring-up cross-check only.
// ---------------------------------------------------------------------------------------------
__global__ void flash_simt_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                  const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out,
                                  long long ldq, long long ldk, long long ldv, long long ldo,
                                  long long bsq, long long bsk, long long bsv, long long bso, int Lq,
                                  int Lk, int DV, float scale) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= Lq) return;
  const __nv_bfloat16* qr = q + b * bsq + warp * ldq;
  float qv[8];
  for (int i = 0; i < 8; ++i) qv[i] = __bfloat162float(qr[lane * 8 + i]);
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // DV/32 values per lane (<= 8)
  const int per = DV / 32;
  for (int t = 0; t < Lk; ++t) {
    const __nv_bfloat16* kr = k + b * bsk + t * ldk;
    float d = 0.f;
    for (int i = 0; i < 8; ++i) d += qv[i] * __bfloat162float(kr[lane * 8 + i]);
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    d *= scale;
    const float mn = fmaxf(m, d);
    const float a = __expf(m - mn), e = __expf(d - mn);
    l = l * a + e;
    const __nv_bfloat16* vr = v + b * bsv + t * ldv;
    for (int i = 0; i < per; ++i) acc[i] = acc[i] * a + e * __bfloat162float(vr[lane * per + i]);
    m = mn;
  }
  __nv_bfloat16* orow = out + b * bso + warp * ldo;
  for (int i = 0; i < per; ++i) orow[lane * per + i] = __float2bfloat16(acc[i] / l);
}

// workspace layout: [items] arrival counters (u32, zero between launches), padded to 256 bytes, then
// [items][2 halves][128 rows][DV + 4] f32 partials
static inline size_t flash_ws_count_bytes(int items) { return (static_cast<size_t>(items) * 4 + 255) / 256 * 256; }

// true when the 64-wide kernel runs an item as two key halves (and, with them, two alternating softmax groups)
static bool flash_two_phase(const ds2_flash_args* a, int BN, int DV) {
  const int all_tiles = (a->Lk + BN - 1) / BN;
  return DV == 64 && a->impl == 0 && all_tiles >= 16 && a->workspace != nullptr &&
         a->workspace_bytes >= ds2_flash_workspace_bytes(a->B, a->Lq, DV);
}

static inline int all_tiles_of(const ds2_flash_args* a, int bn) { return (a->Lk + bn - 1) / bn; }

template <int DV, int BN, int NQ, int KS, int VS, int SP, int QT, int IL = 1, int TP = 0, int HD = 256>
static int launch_flash(const ds2_flash_args* a, cudaStream_t st, int H = 1, int D = 0, int win_Hm = 0, int win_Wm = 0) {
  using Cfg = FlashCfg<DV, BN, NQ, KS, VS, SP, QT, IL, TP, HD>;
  CUtensorMap tq, tk, tv;
  // multi-head variant: the maps span the H * D columns of all heads (boxes reaching past them are zero-filled)
  const uint64_t kcols = HD == 256 ? 256 : static_cast<uint64_t>(H) * D;
  const uint64_t vcols = HD == 256 ? static_cast<uint64_t>(DV) : static_cast<uint64_t>(H) * D;
  {
    const uint64_t dims[3] = {kcols, static_cast<uint64_t>(a->Lq), static_cast<uint64_t>(a->B)};
    const uint64_t str[2] = {static_cast<uint64_t>(a->ldq) * 2, static_cast<uint64_t>(a->bsq) * 2};
    const uint32_t box[3] = {64, kQM, 1};
    int rc = make_tmap_bf16(&tq, a->q, 3, dims, str, box);
    if (rc) return rc;
  }
  if (win_Wm > 0) {
    // window mode: {columns, x, y, batch}; a key tile is 8 raster rows of the 16-token-wide window
    for (int m = 0; m < 2; ++m) {
      const int64_t ld = m == 0 ? a->ldk : a->ldv, bs = m == 0 ? a->bsk : a->bsv;
      const uint64_t dims[4] = {kcols, static_cast<uint64_t>(win_Wm), static_cast<uint64_t>(win_Hm), static_cast<uint64_t>(a->B)};
      const uint64_t str[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * win_Wm, static_cast<uint64_t>(bs) * 2};
      const uint32_t box[4] = {64, 16, BN / 16, 1};
      int rc = make_tmap_bf16(m == 0 ? &tk : &tv, m == 0 ? a->k : a->v, 4, dims, str, box);
      if (rc) return rc;
    }
  } else {
  {
    const uint64_t dims[3] = {kcols, static_cast<uint64_t>(a->Lk), static_cast<uint64_t>(a->B)};
    const uint64_t str[2] = {static_cast<uint64_t>(a->ldk) * 2, static_cast<uint64_t>(a->bsk) * 2};
    const uint32_t box[3] = {64, BN, 1};
    int rc = make_tmap_bf16(&tk, a->k, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {vcols, static_cast<uint64_t>(a->Lk), static_cast<uint64_t>(a->B)};
    const uint64_t str[2] = {static_cast<uint64_t>(a->ldv) * 2, static_cast<uint64_t>(a->bsv) * 2};
    const uint32_t box[3] = {64, BN, 1};
    int rc = make_tmap_bf16(&tv, a->v, 3, dims, str, box);
    if (rc) return rc;
  }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_d256_tcgen05_kernel<DV, BN, NQ, KS, VS, SP, QT, IL, TP, HD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_flash_attn: cudaFuncSetAttribute: %s",
                cudaGetErrorString(e));
    attr_set = true;
  }
  FlashParams p;
  p.B = a->B;
  p.Lq = a->Lq;
  p.Lk = a->Lk;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.dbg = a->impl == 5 ? 1 : (a->impl == 6 ? 2 : (a->impl == 8 ? 3 : (a->impl == 13 ? 4 : 0)));
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.ldq = a->ldq;
  p.bsq = a->bsq;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.ldo = a->ldo;
  p.bso = a->bso;
  p.H = H;
  p.D = D;
  p.Wm = win_Wm;
  p.win_nwx = win_Wm / 16;
  p.win_nwin = (win_Hm / 16) * (win_Wm / 16);
  // ---- grid: whole items + (two-phase only) the items of the partial last wave as two half-length CTAs each ----
  p.q_tiles = (a->Lq + kQM * NQ - 1) / (kQM * NQ);
  const int items = p.q_tiles * a->B * H * (win_Wm > 0 ? p.win_nwin : 1);
  const int sms = sm_count();
  DS2_REQUIRE(sms > 0, DS2_E_NODEVICE, "ds2_flash_attn: no CUDA device");
  const int all_tiles = (a->Lk + BN - 1) / BN;
  // Two key halves whenever the caller lends a workspace and the key range is long enough to be worth it.  The decision
  // depends on the key count and the query-tile count only (never on the batch size or the SM count), so the
  // arithmetic of an item — and with it "tracked in a batch == tracked alone" — does not depend on the grid.
  p.two_phase = TP;
  DS2_REQUIRE(!TP || flash_two_phase(a, BN, DV), DS2_E_ARG, "ds2_flash_attn: the two-key-halves kernel needs the workspace");
  (void)all_tiles;
  int tail = 0;
  if (p.two_phase) {
    // the partial last wave runs as half-length CTAs when they all fit into one wave (3.46 waves -> 3 + 0.5 at 16
    // objects); which items are split is a scheduling choice without numerical consequences
    tail = items % sms;
    if (2 * tail > sms) tail = 0;
    if (a->impl_flags & 1) tail = 0;          // tests: force whole items
    if (a->impl_flags & 2) tail = items;      // tests: force every item to run as two CTAs
  }
  p.n_full = items - tail;
  p.ws_count = reinterpret_cast<unsigned int*>(a->workspace);
  p.ws = p.two_phase ? reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) + flash_ws_count_bytes(items)) : nullptr;
  const int grid = p.n_full + 2 * tail;
  DS2_LAUNCH((flash_d256_tcgen05_kernel<DV, BN, NQ, KS, VS, SP, QT, IL, TP, HD>), grid, Cfg::kThreads, Cfg::kSmem, st, tq, tk, tv, p);
  return post_launch("flash_d256_tcgen05_kernel");
}

// Hiera global attention (window_size == 0 blocks, hieradet.py:57-82) on the flash kernel: head_dim in (64, 80], H heads side
// by side in the token-major qkv matrix.  Returns -1 for shapes it does not cover (ds2_mha then takes the next kernel).
int launch_glob_flash(const ds2_mha_args* a, cudaStream_t st, int sp) {
  const bool win = a->window == 16;     // 16 x 16 windows: 256 queries x 256 keys per (window, head)
  if ((a->window != 0 && !win) || a->q_pool) return -1;
  if (a->D <= 64 || a->D > 80 || (a->D % 8) != 0) return -1;
  if (win) {
    if ((a->Hm % 16) != 0 || (a->Wm % 16) != 0 || a->Hm <= 0 || a->Wm <= 0) return -1;
  } else {
    if (a->Lq < 128 || (a->Lq % 128) != 0 || a->Lk < 128 || (a->Lk % 128) != 0) return -1;
    if (a->Lk_valid != 0 && a->Lk_valid != a->Lk) return -1;
  }
  if ((a->q_tok_stride % 8) || (a->k_tok_stride % 8) || (a->v_tok_stride % 8) || (a->o_tok_stride % 8)) return -1;
  if ((a->q_bs % 8) || (a->k_bs % 8) || (a->v_bs % 8) || (a->o_bs % 8)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15)
    return -1;
  // a head's columns must lie inside a token's row of every operand (the TMA maps are H * D columns wide)
  if (a->k_tok_stride < static_cast<int64_t>(a->H) * a->D || a->v_tok_stride < static_cast<int64_t>(a->H) * a->D) return -1;
  ds2_flash_args f;
  memset(&f, 0, sizeof(f));
  f.B = a->B;
  f.Lq = win ? 256 : a->Lq;
  f.Lk = win ? 256 : a->Lk;
  f.DV = 80;
  f.q = a->q;
  f.k = a->k;
  f.v = a->v;
  f.out = a->out;
  f.ldq = a->q_tok_stride;
  f.ldk = a->k_tok_stride;
  f.ldv = a->v_tok_stride;
  f.ldo = a->o_tok_stride;
  f.bsq = a->q_bs;
  f.bsk = a->k_bs;
  f.bsv = a->v_bs;
  f.bso = a->o_bs;
  f.scale = a->scale;
  if (const char* e = getenv("DS2_GLOB_DBG")) f.impl = atoi(e);   // timing experiments (5 = first box of every tile only)
  const int Hm = win ? a->Hm : 0, Wm = win ? a->Wm : 0;
  if (sp == 2) return launch_flash<80, 128, 1, 3, 2, 2, 1, 1, 0, 80>(&f, st, a->H, a->D, Hm, Wm);
  return launch_flash<80, 128, 1, 3, 2, 1, 1, 1, 0, 80>(&f, st, a->H, a->D, Hm, Wm);
}

}  // namespace ds2

extern "C" int64_t ds2_flash_workspace_bytes(int32_t B, int32_t Lq, int32_t DV) {
  if (B <= 0 || Lq <= 0 || DV != 64) return 0;     // only the 64-wide (cross-attention) kernel runs in two key halves
  const int items = B * ((Lq + ds2::kQM - 1) / ds2::kQM);
  // [items] counters + [items][2 key halves x 2 softmax groups][128 rows][DV + 4] f32
  return static_cast<int64_t>(ds2::flash_ws_count_bytes(items) +
                              static_cast<size_t>(items) * 4 * ds2::kQM * (DV + 4) * sizeof(float));
}

extern "C" int ds2_debug_flash_stalls(unsigned long long* out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, ds2::g_flash_stall, 16 * sizeof(unsigned long long));
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(ds2::g_flash_stall, z, sizeof(z));
  }
  return e == cudaSuccess ? DS2_OK : static_cast<int>(e);
}

extern "C" int ds2_flash_attn(const ds2_flash_args* a, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a != nullptr, DS2_E_ARG, "ds2_flash_attn: null args");
  DS2_REQUIRE(a->B > 0 && a->Lq > 0 && a->Lk > 0, DS2_E_ARG, "ds2_flash_attn: bad shape");
  DS2_REQUIRE(a->DV == 64 || a->DV == 256, DS2_E_ARG, "ds2_flash_attn: DV must be 64 or 256 (got %d)",
              a->DV);
  DS2_REQUIRE(a->q && a->k && a->v && a->out, DS2_E_ARG, "ds2_flash_attn: null pointer");
  cudaStream_t st = as_stream(stream);
  if (a->impl == 1) {
    dim3 grid((a->Lq * 32 + 255) / 256, a->B);
    DS2_LAUNCH((flash_simt_kernel), grid, 256, 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(a->q), reinterpret_cast<const __nv_bfloat16*>(a->k),
        reinterpret_cast<const __nv_bfloat16*>(a->v), reinterpret_cast<__nv_bfloat16*>(a->out), a->ldq,
        a->ldk, a->ldv, a->ldo, a->bsq, a->bsk, a->bsv, a->bso, a->Lq, a->Lk, a->DV, a->scale);
    return post_launch("flash_simt_kernel");
  }
  DS2_REQUIRE((a->ldq % 8) == 0 && (a->ldk % 8) == 0 && (a->ldv % 8) == 0 && (a->ldo % 8) == 0 &&
                  (a->bsq % 8) == 0 && (a->bsk % 8) == 0 && (a->bsv % 8) == 0 && (a->bso % 8) == 0,
              DS2_E_ALIGN, "ds2_flash_attn: pitches must be multiples of 8 elements");
  // impl: 0 = default: one query tile per CTA, Q resident in TMEM (TS MMA), one softmax warpgroup
  //       9 = same with every softmax row split over two warpgroups (same speed once the kernel became
  //           tensor-bound; its bf16-rounded running max changes P by rounding noise, so not the default)
  //       2 = Q as a shared-memory operand (SS MMA), 3 = two query tiles per CTA / 64-key tiles (SS)
  //       5, 6, 8 = timing experiments (half K loads, no exps, barrier-stall accounting)
  // With a workspace (ds2_flash_workspace_bytes) the default 64-wide kernel computes every item as two key halves and
  // may run the items of the partial last wave as two CTAs each (see FlashParams); impl_flags bit 0 / bit 1 force
  // "never split" / "always split" (tests: both must give the same bits).
  if (a->DV == 64) {
    if (a->impl == 3) return launch_flash<64, 64, 2, 2, 3, 1, 0>(a, st);
    if (a->impl == 2) return launch_flash<64, 128, 1, 2, 2, 1, 0>(a, st);
    if (a->impl == 9) return launch_flash<64, 128, 1, 3, 2, 2, 1>(a, st);
    // Measured on one B200, same box, B=16 N=28736 (profiles/r2_s9_flash_variants_ab.txt): single pass 1.13-1.16 ms; two key
    // halves 1.20 (tail wave split) / 1.27 ms (whole items); + alternating softmax groups 1.27 / 1.33 ms (its 320 threads
    // leave 168 registers: the 128-column row spills; reading S twice in chunks instead: 1.97 ms).  The single pass is the
    // default; DS2_FLASH_IL=1 / 2 select the other two for measurements (they need the caller's workspace).
    // DS2_FLASH_IL=3: the two alternating groups with their results merged INSIDE the CTA (no workspace) and the
    // registers moved from the control warpgroup to the softmax warpgroups with setmaxnreg (40 / 232: no spill in the
    // softmax loop): 3-5 % faster per launch in isolation, nothing inside the frame (12.75 / 12.83 against 12.79 / 12.57 ms
    // per step, SM clock 1940 against 1855 MHz: the frame runs at the power cap, a busier kernel is clocked lower) — opt-in.
    // Also measured and dropped: every 2nd / 4th exponential of a row as a cubic on the FMA pipe instead of MUFU
    // (+74 % / +27 % per launch: the softmax warp is bound by its dependent-issue latency, not by the MUFU rate).
    static const int il_env = [] {
      const char* e = getenv("DS2_FLASH_IL");
      return e ? atoi(e) : 0;
    }();
    // il_env: 2 = two key halves + two alternating softmax groups (default with a workspace), 1 = two key halves, one
    // group, 0 = ignore the workspace (single pass)
    // il_env 3: two alternating softmax groups merged inside the CTA (no workspace), registers moved to them with setmaxnreg
    if (il_env == 3 && all_tiles_of(a, 128) >= 4) return launch_flash<64, 128, 1, 3, 2, 1, 1, 2, 0>(a, st);
    if (il_env == 2 && flash_two_phase(a, 128, 64)) return launch_flash<64, 128, 1, 3, 2, 1, 1, 2, 1>(a, st);
    if (il_env == 1 && flash_two_phase(a, 128, 64)) return launch_flash<64, 128, 1, 3, 2, 1, 1, 1, 1>(a, st);
    return launch_flash<64, 128, 1, 3, 2, 1, 1>(a, st);
  }
  if (a->impl == 2) return launch_flash<256, 64, 1, 2, 2, 1, 0>(a, st);
  if (a->impl == 9) return launch_flash<256, 64, 1, 3, 3, 2, 1>(a, st);
  return launch_flash<256, 64, 1, 3, 3, 1, 1>(a, st);
}
