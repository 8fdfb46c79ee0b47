// conv.cu — convolution front-ends: im2col producers for the tcgen05 GEMM (patch embedding,
// 3x3/s2 mask-downsampler stages), the depth-wise 7x7 of the ConvNeXt fuser, and the fused first
// stages of the mask down-sampler (bilinear x4 + sigmoid + conv + LayerNorm2d + GELU in one pass,
// so the 1024^2 high-resolution mask is never written to HBM).
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include "common.h"

namespace ds2 {

__device__ __forceinline__ float gelu_erf_c(float x) { return gelu_erf_fast(x); }

// ---- patch embedding im2col: fp16 [3,S,S] -> bf16 [(S/4)^2, Kpad], k = c*49 + ky*7 + kx ----------
__global__ void im2col_patch_kernel(const __half* __restrict__ img, __nv_bfloat16* __restrict__ out, int S,
                                    int Kpad) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int So = S / 4;
  const int K8 = Kpad / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(So) * So * K8) return;
  const int k8 = static_cast<int>(i % K8);
  const int pix = static_cast<int>(i / K8);
  const int oy = pix / So, ox = pix % So;
  uint4 o;
  __nv_bfloat16* oe = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k8 * 8 + e;
    float v = 0.f;
    if (k < 147) {
      const int c = k / 49, r = k % 49, ky = r / 7, kx = r % 7;
      const int iy = oy * 4 - 3 + ky, ix = ox * 4 - 3 + kx;
      if (iy >= 0 && iy < S && ix >= 0 && ix < S) v = __half2float(img[(static_cast<long long>(c) * S + iy) * S + ix]);
    }
    oe[e] = __float2bfloat16(v);
  }
  *reinterpret_cast<uint4*>(out + static_cast<long long>(pix) * Kpad + k8 * 8) = o;
}

// ---- im2col k3 s2 p1, channels-last bf16: [B,Hi,Wi,C] -> [B*Ho*Wo, 9*C], k = (ky*3+kx)*C + c ----
__global__ void im2col_k3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B,
                                   int Hi, int Wi, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int Ho = Hi / 2, Wo = Wi / 2, C8 = C / 8;
  const long long n = static_cast<long long>(B) * Ho * Wo * 9 * C8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c8 = static_cast<int>(i % C8);
  long long r = i / C8;
  const int tap = static_cast<int>(r % 9);
  r /= 9;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi)
    v = *reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * Hi + iy) * Wi + ix) * C + c8 * 8);
  const long long row = (static_cast<long long>(b) * Ho + oy) * Wo + ox;
  *reinterpret_cast<uint4*>(out + row * (9LL * C) + tap * C + c8 * 8) = v;
}

// ---- depth-wise 7x7, pad 3, channels-last f32 ---------------------------------------------------
// One thread = one channel x (DWY rows x DWX columns) of outputs: each of the DWY+6 input rows of the patch is
// loaded once (DWX+6 values) and feeds every output row it overlaps, so a thread issues
// (DWY+6)(DWX+6)/(DWY*DWX) = 3.4 loads per output at 4 x 16 (9.6 with one output row, 49 naively); threads of
// a warp are consecutive channels, so every load / store is one coalesced 128-byte line.
template <int DWX, int DWY>
__global__ void __launch_bounds__(256) dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int B,
                                                      int Hm, int Wm, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int tiles_x = (Wm + DWX - 1) / DWX;
  const int tiles_y = (Hm + DWY - 1) / DWY;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * DWX, y0 = ty * DWY;
  float wr[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) wr[i] = __ldg(w + c * 49 + i);
  float acc[DWY][DWX];
  const float bv = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
  for (int r = 0; r < DWY; ++r)
#pragma unroll
    for (int o = 0; o < DWX; ++o) acc[r][o] = bv;
#pragma unroll
  for (int ri = 0; ri < DWY + 6; ++ri) {
    const int iy = y0 - 3 + ri;
    // no early-out on rows outside the map: with straight-line code the scheduler hoists the loads of row
    // ri + 1 above the FMAs of row ri (zero rows contribute nothing)
    const bool rv = iy >= 0 && iy < Hm;
    const float* rowp = x + (static_cast<long long>(b) * Hm + (rv ? iy : 0)) * Wm * C + c;
    float row[DWX + 6];
#pragma unroll
    for (int j = 0; j < DWX + 6; ++j) {
      const int ix = x0 - 3 + j;
      row[j] = (rv && ix >= 0 && ix < Wm) ? __ldg(rowp + static_cast<long long>(ix) * C) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < DWY; ++r) {
      const int ky = ri - r;  // input row ri is tap row ky of output row r
      if (ky < 0 || ky > 6) continue;
#pragma unroll
      for (int o = 0; o < DWX; ++o)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc[r][o] = fmaf(row[o + kx], wr[ky * 7 + kx], acc[r][o]);
    }
  }
#pragma unroll
  for (int r = 0; r < DWY; ++r) {
    if (y0 + r >= Hm) continue;
    float* yp = y + ((static_cast<long long>(b) * Hm + y0 + r) * Wm + x0) * C + c;
#pragma unroll
    for (int o = 0; o < DWX; ++o)
      if (x0 + o < Wm) yp[static_cast<long long>(o) * C] = acc[r][o];
  }
}

// Shared-memory version (the default).  The register-tiled kernel above is 4096 straight-line instructions (64 KB of
// code, 220 registers, one CTA per SM): it ran at 1/6 of the FMA rate, stalled on instruction fetch and on the 22 global
// loads per input row.  Here a CTA stages the (16+6)^2 x 32-channel input patch and the 49 x 32 weights in shared memory
// with cp.async (zero-filled outside the map), a warp owns two output rows, a lane one channel: the loop over the 8 input
// rows that feed the two output rows is rolled (~300 instructions), every shared-memory access is 32 consecutive words,
// and three CTAs fit an SM.  Per output: acc = bias, then taps in (ky, kx) order with fmaf — the same sequence as the
// register-tiled kernel, so the two are bit-identical.
constexpr int kDwT = 16, kDwP = kDwT + 6, kDwC = 32;
constexpr int kDwSmem = (kDwP * kDwP * kDwC + 49 * kDwC) * 4;

__global__ void __launch_bounds__(256, 3) dwconv7_smem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y,
                                                              int B, int Hm, int Wm, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  extern __shared__ __align__(16) float dsm[];
  float* patch = dsm;                            // [kDwP][kDwP][32]
  float* wsm = dsm + kDwP * kDwP * kDwC;         // [49][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * kDwC;
  const int tiles_x = (Wm + kDwT - 1) / kDwT, tiles_y = (Hm + kDwT - 1) / kDwT;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * kDwT, y0 = ty * kDwT;
  for (int i = threadIdx.x; i < 49 * kDwC; i += 256) wsm[i] = __ldg(w + (c0 + (i & 31)) * 49 + (i >> 5));
  for (int i = threadIdx.x; i < kDwP * kDwP * (kDwC / 4); i += 256) {
    const int q = i & 7, pix = i >> 3;
    const int py = pix / kDwP, px = pix - py * kDwP;
    const int iy = y0 - 3 + py, ix = x0 - 3 + px;
    float* dst = patch + pix * kDwC + q * 4;
    if (iy >= 0 && iy < Hm && ix >= 0 && ix < Wm) {
      const float* src = x + ((static_cast<long long>(b) * Hm + iy) * Wm + ix) * C + c0 + q * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
                   "l"(src));
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int r0 = warp * 2;
  const float bv = bias ? __ldg(bias + c0 + lane) : 0.f;
  float acc0[kDwT], acc1[kDwT];
#pragma unroll
  for (int o = 0; o < kDwT; ++o) acc0[o] = acc1[o] = bv;
#pragma unroll 1
  for (int ri = 0; ri < 8; ++ri) {
    const float* prow = patch + (r0 + ri) * kDwP * kDwC + lane;
    float row[kDwP];
#pragma unroll
    for (int j = 0; j < kDwP; ++j) row[j] = prow[j * kDwC];
    if (ri < 7) {   // tap row ri of output row r0
      float wk[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) wk[kx] = wsm[(ri * 7 + kx) * kDwC + lane];
#pragma unroll
      for (int o = 0; o < kDwT; ++o)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc0[o] = fmaf(row[o + kx], wk[kx], acc0[o]);
    }
    if (ri >= 1) {  // tap row ri - 1 of output row r0 + 1
      float wk[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) wk[kx] = wsm[((ri - 1) * 7 + kx) * kDwC + lane];
#pragma unroll
      for (int o = 0; o < kDwT; ++o)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc1[o] = fmaf(row[o + kx], wk[kx], acc1[o]);
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int oy = y0 + r0 + rr;
    if (oy >= Hm) continue;
    float* yp = y + ((static_cast<long long>(b) * Hm + oy) * Wm + x0) * C + c0 + lane;
#pragma unroll
    for (int o = 0; o < kDwT; ++o)
      if (x0 + o < Wm) yp[static_cast<long long>(o) * C] = rr ? acc1[o] : acc0[o];
  }
}

// ---- mask down-sampler stage 1 (fused) -----------------------------------------------------------
// PyTorch bilinear (align_corners=False) source index / weights for scale 1/4.
__device__ __forceinline__ float hires_mask_value(const float* __restrict__ lr, int Sl, int Y, int X) {
  const float sy = fmaxf(0.25f * (Y + 0.5f) - 0.5f, 0.f);
  const float sx = fmaxf(0.25f * (X + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Sl - 1), x0 = min(static_cast<int>(sx), Sl - 1);
  const int y1 = y0 + (y0 < Sl - 1 ? 1 : 0), x1 = x0 + (x0 < Sl - 1 ? 1 : 0);
  const float ly = fminf(fmaxf(sy - y0, 0.f), 1.f), lx = fminf(fmaxf(sx - x0, 0.f), 1.f);
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * lr[y0 * Sl + x0] + lx * lr[y0 * Sl + x1]) + ly * (hx * lr[y1 * Sl + x0] + lx * lr[y1 * Sl + x1]);
}

// One CTA = 32 x 8 outputs.  The (2*32+1) x (2*8+1) patch of the never-materialised 1024^2 mask that these outputs read
// is evaluated ONCE per CTA into shared memory (4.3 bilinear + sigmoid evaluations per output instead of 9; positions
// outside the mask hold the conv's zero padding), then every thread convolves its 3x3 window from there.
constexpr int kS1TX = 32, kS1TY = 8;
constexpr int kS1PW = 2 * kS1TX + 1, kS1PH = 2 * kS1TY + 1;

__global__ void __launch_bounds__(kS1TX * kS1TY)
maskds_stage1_kernel(const float* __restrict__ lowres, int B, int Sl, int binarize, float scale, float bias,
                     const float* __restrict__ w, const float* __restrict__ cb, const float* __restrict__ lnw,
                     const float* __restrict__ lnb, __nv_bfloat16* __restrict__ out) {
  __shared__ float sh[kS1PH][kS1PW + 1];
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int So = Sl * 2, Sh = Sl * 4;
  const int tiles_x = (So + kS1TX - 1) / kS1TX, tiles_y = (So + kS1TY - 1) / kS1TY;
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * kS1TX;
  t /= tiles_x;
  const int ty0 = (t % tiles_y) * kS1TY;
  const int b = t / tiles_y;
  const float* lr = lowres + static_cast<long long>(b) * Sl * Sl;
  const int X0 = 2 * tx0 - 1, Y0 = 2 * ty0 - 1;
  for (int i = threadIdx.x; i < kS1PH * kS1PW; i += blockDim.x) {
    const int py = i / kS1PW, px = i - py * kS1PW;
    const int Y = Y0 + py, X = X0 + px;
    float m = 0.f;
    if (Y >= 0 && Y < Sh && X >= 0 && X < Sh) {
      const float hv = hires_mask_value(lr, Sl, Y, X);
      m = (binarize ? (hv > 0.f ? 1.f : 0.f) : 1.f / (1.f + expf(-hv))) * scale + bias;
    }
    sh[py][px] = m;
  }
  float wr[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) wr[i] = __ldg(w + i);
  __syncthreads();
  const int lx = threadIdx.x % kS1TX, ly = threadIdx.x / kS1TX;
  const int ox = tx0 + lx, oy = ty0 + ly;
  if (ox >= So || oy >= So) return;
  float acc[4] = {cb[0], cb[1], cb[2], cb[3]};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float m = sh[2 * ly + ky][2 * lx + kx];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = fmaf(m, wr[c * 9 + ky * 3 + kx], acc[c]);
    }
  const float u = 0.25f * (acc[0] + acc[1] + acc[2] + acc[3]);
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) sq += (acc[c] - u) * (acc[c] - u);
  const float rstd = rsqrtf(0.25f * sq + 1e-6f);
  uint2 o;
  __nv_bfloat16* oe = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
  for (int c = 0; c < 4; ++c) oe[c] = __float2bfloat16(gelu_erf_c((acc[c] - u) * rstd * lnw[c] + lnb[c]));
  *reinterpret_cast<uint2*>(out + ((static_cast<long long>(b) * So + oy) * So + ox) * 4) = o;
}

// ---- direct conv3x3 s2 p1 (small Cin, Cout <= 16) + LN2d + GELU, channels-last bf16 ----------------
template <int CIN, int COUT>
__global__ void maskds_conv_kernel(const __nv_bfloat16* __restrict__ x, int B, int Hi, int Wi,
                                   const float* __restrict__ w, const float* __restrict__ cb,
                                   const float* __restrict__ lnw, const float* __restrict__ lnb,
                                   __nv_bfloat16* __restrict__ out) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  // weights re-laid out [tap][cin][cout]: the inner loop reads four output channels per 16-byte broadcast load (one
  // shared-memory load per FMA made this kernel LSU-bound: 576 loads per output pixel)
  __shared__ __align__(16) float sw[9 * CIN * COUT];
  for (int i = threadIdx.x; i < COUT * CIN * 9; i += blockDim.x) {
    const int o = i / (CIN * 9), c = (i / 9) % CIN, tap = i % 9;
    sw[(tap * CIN + c) * COUT + o] = w[i];
  }
  __syncthreads();
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * Ho * Wo) return;
  const int ox = static_cast<int>(i % Wo);
  const int oy = static_cast<int>((i / Wo) % Ho);
  const int b = static_cast<int>(i / (static_cast<long long>(Wo) * Ho));
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = cb[o];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= Hi) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= Wi) continue;
      const __nv_bfloat16* px = x + ((static_cast<long long>(b) * Hi + iy) * Wi + ix) * CIN;
      float xv[CIN];
      if (CIN == 4) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(px));
        const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        xv[0] = __low2float(p0);
        xv[1] = __high2float(p0);
        xv[2] = __low2float(p1);
        xv[3] = __high2float(p1);
      } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[c] = __bfloat162float(px[c]);
      }
      const float4* wt = reinterpret_cast<const float4*>(sw + (ky * 3 + kx) * CIN * COUT);
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int o4 = 0; o4 < COUT / 4; ++o4) {
          const float4 w4 = wt[c * (COUT / 4) + o4];
          acc[4 * o4] = fmaf(xv[c], w4.x, acc[4 * o4]);
          acc[4 * o4 + 1] = fmaf(xv[c], w4.y, acc[4 * o4 + 1]);
          acc[4 * o4 + 2] = fmaf(xv[c], w4.z, acc[4 * o4 + 2]);
          acc[4 * o4 + 3] = fmaf(xv[c], w4.w, acc[4 * o4 + 3]);
        }
    }
  }
  float u = 0.f;
#pragma unroll
  for (int o = 0; o < COUT; ++o) u += acc[o];
  u /= COUT;
  float s = 0.f;
#pragma unroll
  for (int o = 0; o < COUT; ++o) s += (acc[o] - u) * (acc[o] - u);
  const float rstd = rsqrtf(s / COUT + 1e-6f);
  __nv_bfloat16* po = out + i * COUT;
  if (COUT % 8 == 0) {
#pragma unroll
    for (int o8 = 0; o8 < COUT / 8; ++o8) {
      uint4 pk;
      uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int o = 8 * o8 + 2 * q;
        const __nv_bfloat162 v2 = __floats2bfloat162_rn(gelu_erf_c((acc[o] - u) * rstd * lnw[o] + lnb[o]),
                                                        gelu_erf_c((acc[o + 1] - u) * rstd * lnw[o + 1] + lnb[o + 1]));
        pw[q] = *reinterpret_cast<const uint32_t*>(&v2);
      }
      reinterpret_cast<uint4*>(po)[o8] = pk;
    }
  } else {
#pragma unroll
    for (int o = 0; o < COUT; ++o) po[o] = __float2bfloat16(gelu_erf_c((acc[o] - u) * rstd * lnw[o] + lnb[o]));
  }
}

}  // namespace ds2

extern "C" {

int ds2_im2col_patch(const void* frame_f16, void* out_bf16, int32_t S, int32_t Kpad, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(frame_f16 && out_bf16 && S > 0 && (S % 4) == 0 && Kpad >= 147 && (Kpad % 8) == 0, DS2_E_ARG,
              "ds2_im2col_patch: bad args");
  const long long n = static_cast<long long>(S / 4) * (S / 4) * (Kpad / 8);
  DS2_LAUNCH((im2col_patch_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), 
      reinterpret_cast<const __half*>(frame_f16), reinterpret_cast<__nv_bfloat16*>(out_bf16), S, Kpad);
  return post_launch("im2col_patch_kernel");
}

int ds2_im2col_k3s2(const void* x_bf16, void* out_bf16, int32_t B, int32_t Hi, int32_t Wi, int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x_bf16 && out_bf16 && B > 0 && (Hi % 2) == 0 && (Wi % 2) == 0 && (C % 8) == 0, DS2_E_ARG,
              "ds2_im2col_k3s2: bad args");
  const long long n = static_cast<long long>(B) * (Hi / 2) * (Wi / 2) * 9 * (C / 8);
  DS2_LAUNCH((im2col_k3s2_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), B, Hi, Wi, C);
  return post_launch("im2col_k3s2_kernel");
}

int ds2_dwconv7(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Hm, int32_t Wm,
                int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && w && y && B > 0 && Hm > 0 && Wm > 0 && C > 0, DS2_E_ARG, "ds2_dwconv7: bad args");
  // default: the shared-memory kernel; DS2_DWCONV_TILE = 16x4 | 16x2 | 8x4 | 8x8 selects a register-tiled variant (columns x
  // rows per thread) for measurements
  static const int cfg = [] {
    const char* e = getenv("DS2_DWCONV_TILE");
    if (!e) return 0;
    if (!strcmp(e, "16x2")) return 1;
    if (!strcmp(e, "8x4")) return 2;
    if (!strcmp(e, "8x8")) return 3;
    if (!strcmp(e, "16x4")) return 4;   // the register-tiled default of round 1
    return 0;
  }();
  if (cfg == 0 && (C % kDwC) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    static const cudaError_t attr = cudaFuncSetAttribute(dwconv7_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem);
    DS2_REQUIRE(attr == cudaSuccess, static_cast<int>(attr), "ds2_dwconv7: cudaFuncSetAttribute: %s", cudaGetErrorString(attr));
    dim3 grid(static_cast<unsigned>(B) * ((Hm + kDwT - 1) / kDwT) * ((Wm + kDwT - 1) / kDwT), C / kDwC);
    DS2_LAUNCH((dwconv7_smem_kernel), grid, 256, kDwSmem, as_stream(stream), x, w, bias, y, B, Hm, Wm, C);
    return post_launch("dwconv7_smem_kernel");
  }
  const int dwx = (cfg >= 2) ? 8 : 16;
  const int dwy = cfg == 1 ? 2 : (cfg == 3 ? 8 : 4);
  const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  dim3 grid(static_cast<unsigned>(B) * ((Hm + dwy - 1) / dwy) * ((Wm + dwx - 1) / dwx), (C + threads - 1) / threads);
  if (cfg == 1) DS2_LAUNCH((dwconv7_kernel<16, 2>), grid, threads, 0, as_stream(stream), x, w, bias, y, B, Hm, Wm, C);
  else if (cfg == 2) DS2_LAUNCH((dwconv7_kernel<8, 4>), grid, threads, 0, as_stream(stream), x, w, bias, y, B, Hm, Wm, C);
  else if (cfg == 3) DS2_LAUNCH((dwconv7_kernel<8, 8>), grid, threads, 0, as_stream(stream), x, w, bias, y, B, Hm, Wm, C);
  else DS2_LAUNCH((dwconv7_kernel<16, 4>), grid, threads, 0, as_stream(stream), x, w, bias, y, B, Hm, Wm, C);
  return post_launch("dwconv7_kernel");
}

int ds2_maskds_stage1(const float* lowres, int32_t B, int32_t Sl, int32_t binarize, float scale, float bias,
                      const float* w, const float* b, const float* ln_w, const float* ln_b, void* out_bf16,
                      void* stream) {
  using namespace ds2;
  DS2_REQUIRE(lowres && w && b && ln_w && ln_b && out_bf16 && B > 0 && Sl > 0, DS2_E_ARG,
              "ds2_maskds_stage1: bad args");
  const int So = Sl * 2;
  const long long ctas = static_cast<long long>(B) * ((So + kS1TX - 1) / kS1TX) * ((So + kS1TY - 1) / kS1TY);
  DS2_LAUNCH((maskds_stage1_kernel), static_cast<unsigned>(ctas), kS1TX * kS1TY, 0, as_stream(stream),
      lowres, B, Sl, binarize, scale, bias, w, b, ln_w, ln_b, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return post_launch("maskds_stage1_kernel");
}

int ds2_maskds_conv(const void* x_bf16, int32_t B, int32_t Hi, int32_t Wi, int32_t Cin, int32_t Cout,
                    const float* w, const float* b, const float* ln_w, const float* ln_b, void* out_bf16,
                    void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x_bf16 && w && b && ln_w && ln_b && out_bf16 && B > 0, DS2_E_ARG, "ds2_maskds_conv: bad args");
  DS2_REQUIRE(Cin == 4 && Cout == 16, DS2_E_ARG, "ds2_maskds_conv: only 4->16 is instantiated (got %d->%d)", Cin,
              Cout);
  const long long n = static_cast<long long>(B) * (Hi / 2) * (Wi / 2);
  DS2_LAUNCH((maskds_conv_kernel<4, 16>), static_cast<unsigned>((n + 127) / 128), 128, 0, as_stream(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), B, Hi, Wi, w, b, ln_w, ln_b,
      reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return post_launch("maskds_conv_kernel");
}

}  // extern "C"
