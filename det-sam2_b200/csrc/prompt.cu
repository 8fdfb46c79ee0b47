// prompt.cu — small latency-bound pieces around the mask decoder and the memory bank:
//   * prompt encoder (random-Fourier point embedding + label embeddings) fused with the assembly of
//     the decoder's token matrix,
//   * object-pointer bank tokens with the temporal positional encoding computed in-kernel,
//   * memory-encoder tail (no-object embedding + bf16 store).
#include <math.h>

#include "common.h"

namespace ds2 {

// tokens[b, t, :]:  t < n_out -> out_tokens[t];  t >= n_out -> sparse embedding of point t - n_out
// (point P is the padding point (0,0,label -1) appended when no box is given,
//  sam/prompt_encoder.py:73-95).  One CTA (128 threads) per (b, t); C = 256 = 2 * 128.
__global__ void __launch_bounds__(128) prompt_tokens_kernel(const float* __restrict__ coords,
                                                            const int* __restrict__ labels, int B, int P,
                                                            const float* __restrict__ gauss,      // [2,128]
                                                            const float* __restrict__ point_emb,  // [4,256]
                                                            const float* __restrict__ not_a_point,  // [256]
                                                            const float* __restrict__ out_tokens,   // [n_out,256]
                                                            int n_out, float image_size,
                                                            float* __restrict__ tokens) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int nt = n_out + P + 1;
  const int b = blockIdx.x / nt, t = blockIdx.x % nt;
  const int j = threadIdx.x;  // frequency index 0..127
  float* dst = tokens + (static_cast<long long>(b) * nt + t) * 256;
  if (t < n_out) {
    dst[j] = out_tokens[t * 256 + j];
    dst[j + 128] = out_tokens[t * 256 + 128 + j];
    return;
  }
  const int pidx = t - n_out;
  float x = 0.f, y = 0.f;
  int lab = -1;
  if (pidx < P) {
    x = coords[(static_cast<long long>(b) * P + pidx) * 2 + 0];
    y = coords[(static_cast<long long>(b) * P + pidx) * 2 + 1];
    lab = labels[static_cast<long long>(b) * P + pidx];
  }
  float s = 0.f, c = 0.f;
  if (lab != -1) {
    // (p + 0.5) / S -> [0,1] -> 2u - 1 -> @ G -> * 2 pi   (position_encoding.py:131-158)
    const float u = 2.f * ((x + 0.5f) / image_size) - 1.f;
    const float v = 2.f * ((y + 0.5f) / image_size) - 1.f;
    const float a = 6.283185307179586f * (u * gauss[j] + v * gauss[128 + j]);
    sincosf(a, &s, &c);
  }
  float e0 = s, e1 = c;
  if (lab == -1) {
    e0 = not_a_point[j];
    e1 = not_a_point[128 + j];
  } else if (lab >= 0 && lab < 4) {
    e0 += point_emb[lab * 256 + j];
    e1 += point_emb[lab * 256 + 128 + j];
  }
  dst[j] = e0;
  dst[j + 128] = e1;
}

// Object-pointer tokens of one stored frame (sam2_base.py:588-648): ptr f32 [B,256] -> 4 tokens of 64;
//   tpos = W[64,256] . sine1d(dist / t_diff_max, 256) + bias     (get_1d_sine_pe sam2_utils.py:69-79)
//   kin = bf16(ptr + tpos),  val = bf16(ptr).   grid = B, block = 256.
__global__ void __launch_bounds__(256) bank_ptr_pe_kernel(const float* __restrict__ ptr, float dist_norm,
                                                          const float* __restrict__ w,     // [64,256]
                                                          const float* __restrict__ bias,  // [64]
                                                          __nv_bfloat16* __restrict__ kin,
                                                          __nv_bfloat16* __restrict__ val, long long dst_bs,
                                                          int row0) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  __shared__ float pe[256];
  __shared__ float tp[64];
  const int t = threadIdx.x;
  {
    // pe_dim = 128; dim_t[i] = 10000^(2*(i/2)/128); [sin(pos/dim_t) | cos(pos/dim_t)]
    const int i = t & 127;
    const float dim_t = powf(10000.f, static_cast<float>(2 * (i / 2)) / 128.f);
    const float a = dist_norm / dim_t;
    pe[t] = (t < 128) ? sinf(a) : cosf(a);
  }
  __syncthreads();
  {
    // 4 threads per output channel
    const int c = t >> 2, part = t & 3;
    float acc = 0.f;
    for (int k = part * 64; k < part * 64 + 64; ++k) acc = fmaf(w[c * 256 + k], pe[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) tp[c] = acc + bias[c];
  }
  __syncthreads();
  const int b = blockIdx.x;
  const float pv = ptr[static_cast<long long>(b) * 256 + t];
  const long long dst = static_cast<long long>(b) * dst_bs + static_cast<long long>(row0 + t / 64) * 64 + (t % 64);
  kin[dst] = __float2bfloat16(pv + tp[t % 64]);
  val[dst] = __float2bfloat16(pv);
}

// maskmem = bf16(x + (1 - [score > 0]) * no_obj_embed)      (sam2_base.py:733-741, svp:1337)
__global__ void memenc_finish_kernel(const float* __restrict__ x, const float* __restrict__ score,
                                     const float* __restrict__ no_obj, __nv_bfloat16* __restrict__ out, int B,
                                     int T, int C) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * T * C;
  if (i >= n) return;
  const int c = static_cast<int>(i % C);
  const int b = static_cast<int>(i / (static_cast<long long>(T) * C));
  const float add = score[b] > 0.f ? 0.f : no_obj[c];
  out[i] = __float2bfloat16(x[i] + add);
}

// ---- dense mask prompt (sam2_base.py:399-448, prompt_encoder.py:97-100) -------------------------------
// (1) low = antialiased bilinear x1/4 of (mask * scale + bias)  — F.interpolate(..., antialias=True):
//     triangle filter of support 4 input pixels around centre 4(i + 0.5), taps clipped at the border and
//     renormalised (ATen UpSampleKernel _compute_indices_weights_aa), separable.
__global__ void __launch_bounds__(256) downsample4_aa_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int S,
                                                             float scale, float bias) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int So = S / 4;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * So * So) return;
  const int ox = static_cast<int>(i % So), oy = static_cast<int>((i / So) % So), b = static_cast<int>(i / (static_cast<long long>(So) * So));
  float wx[8], wy[8];
  int x0, y0, nx, ny;
  auto taps = [&](int o, float(&w)[8], int& lo, int& n) {
    const float c = 4.0f * (o + 0.5f);
    lo = max(static_cast<int>(c - 4.0f + 0.5f), 0);
    const int hi = min(static_cast<int>(c + 4.0f + 0.5f), S);
    n = hi - lo;
    float tot = 0.f;
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
      if (j < n) v = fmaxf(1.0f - fabsf((j + lo - c + 0.5f) * 0.25f), 0.f);
      w[j] = v;
      tot += v;
    }
    for (int j = 0; j < 8; ++j) w[j] /= tot;
  };
  taps(ox, wx, x0, nx);
  taps(oy, wy, y0, ny);
  const float* src = x + static_cast<long long>(b) * S * S;
  float acc = 0.f;
  for (int r = 0; r < ny; ++r) {
    float h = 0.f;
    for (int c = 0; c < nx; ++c) h = fmaf(wx[c], fmaf(src[static_cast<long long>(y0 + r) * S + x0 + c], scale, bias), h);
    acc = fmaf(wy[r], h, acc);
  }
  y[i] = acc;
}

// (2) mask_downsample (Conv2d 1->1, k4 s4) followed by the prompt encoder's mask_downscaling up to its last
//     1x1 conv: Conv2d(1->4, k2 s2) LN2d GELU Conv2d(4->16, k2 s2) LN2d GELU.  All strides equal the kernel sizes,
//     so one output token depends on one 16x16 patch of the mask: one thread per token, 16 bf16 features out
//     (the 16->256 1x1 conv is a ds2_gemm).
struct MaskEmbedParams {
  const float* mask;  // [B, S, S]
  int B, S;
  const float *wds, *bds;             // [16], [1]
  const float *w0, *b0, *g0, *be0;    // [4][4], [4], LN [4]
  const float *w3, *b3, *g3, *be3;    // [16][16] ((cin, ky, kx) per output), [16], LN [16]
  __nv_bfloat16* out;                 // [B * (S/16)^2, 16]
};
__device__ __forceinline__ float gelu_erf_p(float x) { return gelu_erf_fast(x); }

__global__ void __launch_bounds__(128) mask_prompt_embed_kernel(const MaskEmbedParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  // wds == nullptr: the input already IS the mask at the prompt encoder's input size (low-resolution logits of a
  // previous decode fed back as a dense prompt, sam2_base.py:306-329) — no k4 s4 stage, 4 x 4 values per output token
  const bool direct = (p.wds == nullptr);
  const int blk = direct ? 4 : 16;
  const int So = p.S / blk;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(p.B) * So * So) return;
  const int ox = static_cast<int>(i % So), oy = static_cast<int>((i / So) % So), b = static_cast<int>(i / (static_cast<long long>(So) * So));
  const float* src = p.mask + (static_cast<long long>(b) * p.S + oy * blk) * p.S + ox * blk;
  // k4 s4: 4 x 4 values of the 1/4-resolution mask
  float ds[4][4];
  for (int qy = 0; qy < 4; ++qy)
    for (int qx = 0; qx < 4; ++qx) {
      if (direct) {
        ds[qy][qx] = src[static_cast<long long>(qy) * p.S + qx];
        continue;
      }
      float a = p.bds[0];
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) a = fmaf(src[static_cast<long long>(qy * 4 + ky) * p.S + qx * 4 + kx], p.wds[ky * 4 + kx], a);
      ds[qy][qx] = a;
    }
  // k2 s2 1 -> 4, LN over the 4 channels, GELU: 2 x 2 positions
  float h0[2][2][4];
  for (int ry = 0; ry < 2; ++ry)
    for (int rx = 0; rx < 2; ++rx) {
      float v[4], u = 0.f;
      for (int c = 0; c < 4; ++c) {
        float a = p.b0[c];
        for (int ky = 0; ky < 2; ++ky)
          for (int kx = 0; kx < 2; ++kx) a = fmaf(ds[ry * 2 + ky][rx * 2 + kx], p.w0[c * 4 + ky * 2 + kx], a);
        v[c] = a;
        u += a;
      }
      u *= 0.25f;
      float s2 = 0.f;
      for (int c = 0; c < 4; ++c) s2 += (v[c] - u) * (v[c] - u);
      const float rstd = 1.0f / sqrtf(s2 * 0.25f + 1e-6f);
      for (int c = 0; c < 4; ++c) h0[ry][rx][c] = gelu_erf_p((v[c] - u) * rstd * p.g0[c] + p.be0[c]);
    }
  // k2 s2 4 -> 16, LN over the 16 channels, GELU
  float v[16], u = 0.f;
  for (int o = 0; o < 16; ++o) {
    float a = p.b3[o];
    for (int c = 0; c < 4; ++c)
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 2; ++kx) a = fmaf(h0[ky][kx][c], p.w3[o * 16 + c * 4 + ky * 2 + kx], a);
    v[o] = a;
    u += a;
  }
  u *= (1.0f / 16.0f);
  float s2 = 0.f;
  for (int o = 0; o < 16; ++o) s2 += (v[o] - u) * (v[o] - u);
  const float rstd = 1.0f / sqrtf(s2 * (1.0f / 16.0f) + 1e-6f);
  __nv_bfloat16* po = p.out + i * 16;
  for (int o = 0; o < 16; ++o) po[o] = __float2bfloat16(gelu_erf_p((v[o] - u) * rstd * p.g3[o] + p.be3[o]));
}

}  // namespace ds2

extern "C" {

int ds2_prompt_tokens(const float* coords, const int32_t* labels, int32_t B, int32_t P, const float* gauss,
                      const float* point_emb, const float* not_a_point, const float* out_tokens, int32_t n_out,
                      float image_size, float* tokens, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(gauss && point_emb && not_a_point && out_tokens && tokens && B > 0 && P >= 0 && n_out > 0, DS2_E_ARG,
              "ds2_prompt_tokens: bad args");
  DS2_REQUIRE(P == 0 || (coords && labels), DS2_E_ARG, "ds2_prompt_tokens: P > 0 needs coords and labels");
  const int nt = n_out + P + 1;
  DS2_LAUNCH((prompt_tokens_kernel), B * nt, 128, 0, as_stream(stream), coords, labels, B, P, gauss, point_emb, not_a_point,
                                                           out_tokens, n_out, image_size, tokens);
  return post_launch("prompt_tokens_kernel");
}

int ds2_bank_ptr_pe(const float* ptr, float dist_norm, const float* w, const float* bias, void* kin_bf16,
                    void* val_bf16, int32_t B, int64_t dst_bs, int32_t row0, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(ptr && w && bias && kin_bf16 && val_bf16 && B > 0, DS2_E_ARG, "ds2_bank_ptr_pe: bad args");
  DS2_LAUNCH((bank_ptr_pe_kernel), B, 256, 0, as_stream(stream), ptr, dist_norm, w, bias,
                                                     reinterpret_cast<__nv_bfloat16*>(kin_bf16),
                                                     reinterpret_cast<__nv_bfloat16*>(val_bf16), dst_bs, row0);
  return post_launch("bank_ptr_pe_kernel");
}

int ds2_memenc_finish(const float* x, const float* score, const float* no_obj_embed, void* out_bf16, int32_t B,
                      int32_t T, int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && score && no_obj_embed && out_bf16 && B > 0 && T > 0 && C > 0, DS2_E_ARG,
              "ds2_memenc_finish: bad args");
  const long long n = static_cast<long long>(B) * T * C;
  DS2_LAUNCH((memenc_finish_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), 
      x, score, no_obj_embed, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, T, C);
  return post_launch("memenc_finish_kernel");
}

}  // extern "C"

extern "C" int ds2_downsample4_aa(const float* x, float* y, int32_t B, int32_t S, float scale, float bias, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && y && B > 0 && S >= 4 && (S % 4) == 0, DS2_E_ARG, "ds2_downsample4_aa: bad args");
  const long long n = static_cast<long long>(B) * (S / 4) * (S / 4);
  DS2_LAUNCH((downsample4_aa_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), x, y, B, S, scale, bias);
  return post_launch("downsample4_aa_kernel");
}

extern "C" int ds2_mask_prompt_embed(const float* mask, int32_t B, int32_t S, const float* wds, const float* bds,
                                     const float* w0, const float* b0, const float* ln0_w, const float* ln0_b,
                                     const float* w3, const float* b3, const float* ln3_w, const float* ln3_b,
                                     void* out_bf16, void* stream) {
  using namespace ds2;
  const int blk = wds ? 16 : 4;
  DS2_REQUIRE(mask && (wds != nullptr) == (bds != nullptr) && w0 && b0 && ln0_w && ln0_b && w3 && b3 && ln3_w && ln3_b &&
                  out_bf16 && B > 0 && S >= blk && (S % blk) == 0,
              DS2_E_ARG, "ds2_mask_prompt_embed: bad args");
  MaskEmbedParams p;
  p.mask = mask;
  p.B = B;
  p.S = S;
  p.wds = wds;
  p.bds = bds;
  p.w0 = w0;
  p.b0 = b0;
  p.g0 = ln0_w;
  p.be0 = ln0_b;
  p.w3 = w3;
  p.b3 = b3;
  p.g3 = ln3_w;
  p.be3 = ln3_b;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  const long long n = static_cast<long long>(B) * (S / blk) * (S / blk);
  DS2_LAUNCH((mask_prompt_embed_kernel), static_cast<unsigned>((n + 127) / 128), 128, 0, as_stream(stream), p);
  return post_launch("mask_prompt_embed_kernel");
}
