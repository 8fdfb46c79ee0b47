// elementwise.cu — HBM-bound kernels: LayerNorm (warp-shuffle reduction), axpby, casts, pooling,
// FPN top-down add, memory-bank gather.  Vectorised 16-byte accesses, one pass over the data.
#include <math.h>

#include "common.h"

namespace ds2 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gelu_erf_e(float x) { return gelu_erf_fast(x); }

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 32*MAXV), two-pass mean / variance.
// ---------------------------------------------------------------------------------------------
struct LnParams {
  const float* x;
  const __nv_bfloat16* xb;
  long long ldx;
  int rows, C;
  const float* w;
  const float* b;
  float eps;
  int act;
  float* of;
  __nv_bfloat16* ob;
  long long ldo;
  const float* pos;
  int pos_row_mod;
  __nv_bfloat16* ob2;
};

template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    float t = 0.f;
    if (c < p.C) t = p.x ? p.x[row * p.ldx + c] : __bfloat162float(p.xb[row * p.ldx + c]);
    v[i] = t;
    s += t;
  }
  const float mean = warp_sum(s) / p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float d = (c < p.C) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / p.C + p.eps);
  const long long prow = p.pos ? (p.pos_row_mod > 0 ? row % p.pos_row_mod : row) : 0;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < p.C) {
      float y = (v[i] - mean) * rstd * p.w[c] + p.b[c];
      if (p.act == 2) y = gelu_erf_e(y);
      if (p.of) p.of[row * p.ldo + c] = y;
      if (p.ob) p.ob[row * p.ldo + c] = __float2bfloat16(y);
      if (p.ob2) p.ob2[row * p.ldo + c] = __float2bfloat16(y + p.pos[prow * p.C + c]);
    }
  }
}

// Vectorised variant (f32 input, C % 4 == 0, 16-byte aligned rows): each lane owns float4 number
// lane + 32*i of the row, so every load / store instruction of a warp covers 512 contiguous bytes.
template <int MAXV4>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const LnParams p) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const int C4 = p.C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<long long>(row) * p.ldx);
  float4 v[MAXV4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV4; ++i) {
    const int c4 = lane + i * 32;
    v[i] = (c4 < C4) ? xr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV4; ++i) {
    if (lane + i * 32 < C4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / p.C + p.eps);
  const long long prow = p.pos ? (p.pos_row_mod > 0 ? row % p.pos_row_mod : row) : 0;
  const float4* w4 = reinterpret_cast<const float4*>(p.w);
  const float4* b4 = reinterpret_cast<const float4*>(p.b);
#pragma unroll
  for (int i = 0; i < MAXV4; ++i) {
    const int c4 = lane + i * 32;
    if (c4 < C4) {
      const float4 w = __ldg(w4 + c4), b = __ldg(b4 + c4);
      float4 y;
      y.x = (v[i].x - mean) * rstd * w.x + b.x;
      y.y = (v[i].y - mean) * rstd * w.y + b.y;
      y.z = (v[i].z - mean) * rstd * w.z + b.z;
      y.w = (v[i].w - mean) * rstd * w.w + b.w;
      if (p.act == 2) {
        y.x = gelu_erf_e(y.x);
        y.y = gelu_erf_e(y.y);
        y.z = gelu_erf_e(y.z);
        y.w = gelu_erf_e(y.w);
      }
      const long long o = static_cast<long long>(row) * p.ldo + c4 * 4;
      if (p.of) *reinterpret_cast<float4*>(p.of + o) = y;
      if (p.ob) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&lo);
        t.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.ob + o) = t;
      }
      if (p.ob2) {
        const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + prow * p.C) + c4);
        __nv_bfloat162 lo = __floats2bfloat162_rn(y.x + ps.x, y.y + ps.y), hi = __floats2bfloat162_rn(y.z + ps.z, y.w + ps.w);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&lo);
        t.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.ob2 + o) = t;
      }
    }
  }
}

__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4,
                             int C4, int b_row_mod, float alpha, float beta, float* __restrict__ of,
                             __nv_bfloat16* __restrict__ ob) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long row = i / C4;
  const int c4 = static_cast<int>(i % C4);
  float4 x = reinterpret_cast<const float4*>(a)[i];
  float4 r = make_float4(x.x * alpha, x.y * alpha, x.z * alpha, x.w * alpha);
  if (b) {
    const long long br = b_row_mod > 0 ? row % b_row_mod : row;
    const float4 y = reinterpret_cast<const float4*>(b)[br * C4 + c4];
    r.x += y.x * beta;
    r.y += y.y * beta;
    r.z += y.z * beta;
    r.w += y.w * beta;
  }
  if (of) reinterpret_cast<float4*>(of)[i] = r;
  if (ob) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&lo);
    t.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(ob)[i] = t;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     long long n) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&lo);
    t.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(y + i) = t;
  } else {
    for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                     long long n) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}

__global__ void maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int Hm,
                                  int Wm, int C4) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Ho = Hm / 2, Wo = Wm / 2;
  const long long n = static_cast<long long>(B) * Ho * Wo * C4;
  if (i >= n) return;
  const int c = static_cast<int>(i % C4);
  long long r = i / C4;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const float4* src = reinterpret_cast<const float4*>(x);
  auto at = [&](int yy, int xx) { return src[((static_cast<long long>(b) * Hm + yy) * Wm + xx) * C4 + c]; };
  const float4 a = at(2 * oy, 2 * ox), bq = at(2 * oy, 2 * ox + 1), cq = at(2 * oy + 1, 2 * ox),
               d = at(2 * oy + 1, 2 * ox + 1);
  float4 m;
  m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x));
  m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
  m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z));
  m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
  reinterpret_cast<float4*>(y)[i] = m;
}

__global__ void upsample2x_add_kernel(const float* __restrict__ top, const float* __restrict__ lat,
                                      float* __restrict__ y, int B, int Hm, int Wm, int C4) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Ho = Hm * 2, Wo = Wm * 2;
  const long long n = static_cast<long long>(B) * Ho * Wo * C4;
  if (i >= n) return;
  const int c = static_cast<int>(i % C4);
  long long r = i / C4;
  const int ox = static_cast<int>(r % Wo);
  r /= Wo;
  const int oy = static_cast<int>(r % Ho);
  const int b = static_cast<int>(r / Ho);
  const float4 t = reinterpret_cast<const float4*>(top)[((static_cast<long long>(b) * Hm + oy / 2) * Wm + ox / 2) * C4 + c];
  const float4 l = reinterpret_cast<const float4*>(lat)[i];
  reinterpret_cast<float4*>(y)[i] = make_float4(t.x + l.x, t.y + l.y, t.z + l.z, t.w + l.w);
}

// memory-bank gather: 8 channels (16 bytes of bf16) per thread
__global__ void bank_gather_kernel(const __nv_bfloat16* __restrict__ mem, const float* __restrict__ pos,
                                   const float* __restrict__ tpos, __nv_bfloat16* __restrict__ kin,
                                   __nv_bfloat16* __restrict__ val, int B, int T, int C, long long dst_bs,
                                   int row0) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int C8 = C / 8;
  const long long n = static_cast<long long>(B) * T * C8;
  if (i >= n) return;
  const int c8 = static_cast<int>(i % C8);
  long long r = i / C8;
  const int t = static_cast<int>(r % T);
  const int b = static_cast<int>(r / T);
  const uint4 m = *reinterpret_cast<const uint4*>(mem + (static_cast<long long>(b) * T + t) * C + c8 * 8);
  const __nv_bfloat16* me = reinterpret_cast<const __nv_bfloat16*>(&m);
  const float* pp = pos + static_cast<long long>(t) * C + c8 * 8;
  const float* tp = tpos + c8 * 8;
  uint4 ko;
  __nv_bfloat16* ke = reinterpret_cast<__nv_bfloat16*>(&ko);
#pragma unroll
  for (int e = 0; e < 8; ++e) ke[e] = __float2bfloat16(__bfloat162float(me[e]) + (pp[e] + tp[e]));
  const long long dst = static_cast<long long>(b) * dst_bs + static_cast<long long>(row0 + t) * C + c8 * 8;
  *reinterpret_cast<uint4*>(kin + dst) = ko;
  *reinterpret_cast<uint4*>(val + dst) = m;
}

__global__ void bank_ptr_kernel(const float* __restrict__ ptr, const float* __restrict__ tpos,
                                __nv_bfloat16* __restrict__ kin, __nv_bfloat16* __restrict__ val, int B,
                                long long dst_bs, int row0) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over B*256
  if (i >= B * 256) return;
  const int b = i / 256, c = i % 256;
  const float pv = ptr[i];
  const long long dst = static_cast<long long>(b) * dst_bs + static_cast<long long>(row0 + c / 64) * 64 + (c % 64);
  kin[dst] = __float2bfloat16(pv + tpos[c % 64]);
  val[dst] = __float2bfloat16(pv);
}

}  // namespace ds2

extern "C" {

int ds2_layernorm(const ds2_ln_args* a, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a && (a->x || a->x_bf16) && a->w && a->b, DS2_E_ARG, "ds2_layernorm: null pointer");
  DS2_REQUIRE(a->rows > 0 && a->C > 0 && a->C <= 32 * 40, DS2_E_ARG, "ds2_layernorm: C=%d unsupported", a->C);
  DS2_REQUIRE(a->out_f32 || a->out_bf16 || a->out2_bf16, DS2_E_ARG, "ds2_layernorm: no output");
  DS2_REQUIRE(!a->out2_bf16 || a->pos, DS2_E_ARG, "ds2_layernorm: out2 needs pos");
  LnParams p;
  p.x = a->x;
  p.xb = reinterpret_cast<const __nv_bfloat16*>(a->x_bf16);
  p.ldx = a->ldx;
  p.rows = a->rows;
  p.C = a->C;
  p.w = a->w;
  p.b = a->b;
  p.eps = a->eps;
  p.act = a->act;
  p.of = a->out_f32;
  p.ob = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.ldo = a->ldo;
  p.pos = a->pos;
  p.pos_row_mod = a->pos_row_mod;
  p.ob2 = reinterpret_cast<__nv_bfloat16*>(a->out2_bf16);
  const int grid = (a->rows + 7) / 8;
  cudaStream_t st = as_stream(stream);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a->x && (a->C % 4) == 0 && (a->ldx % 4) == 0 && (a->ldo % 4) == 0 && al16(a->x) && al16(a->w) && al16(a->b) &&
      al16(a->out_f32) && al16(a->out_bf16) && al16(a->out2_bf16) && al16(a->pos) && (!a->out2_bf16 || (a->C % 4) == 0)) {
    if (a->C <= 128) DS2_LAUNCH((layernorm_vec_kernel<1>), grid, 256, 0, st, p);
    else if (a->C <= 256) DS2_LAUNCH((layernorm_vec_kernel<2>), grid, 256, 0, st, p);
    else if (a->C <= 640) DS2_LAUNCH((layernorm_vec_kernel<5>), grid, 256, 0, st, p);
    else DS2_LAUNCH((layernorm_vec_kernel<10>), grid, 256, 0, st, p);
    return post_launch("layernorm_vec_kernel");
  }
  if (a->C <= 64) DS2_LAUNCH((layernorm_kernel<2>), grid, 256, 0, st, p);
  else if (a->C <= 256) DS2_LAUNCH((layernorm_kernel<8>), grid, 256, 0, st, p);
  else if (a->C <= 576) DS2_LAUNCH((layernorm_kernel<18>), grid, 256, 0, st, p);
  else DS2_LAUNCH((layernorm_kernel<40>), grid, 256, 0, st, p);
  return post_launch("layernorm_kernel");
}

int ds2_axpby(const float* a, const float* b, int64_t rows, int32_t C, int32_t b_row_mod, float alpha,
              float beta, float* out_f32, void* out_bf16, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(a && rows > 0 && C > 0 && (C % 4) == 0, DS2_E_ARG, "ds2_axpby: bad args (C %% 4 != 0?)");
  DS2_REQUIRE(out_f32 || out_bf16, DS2_E_ARG, "ds2_axpby: no output");
  const long long n4 = rows * (C / 4);
  DS2_LAUNCH((axpby_kernel), static_cast<unsigned>((n4 + 255) / 256), 256, 0, as_stream(stream), 
      a, b, n4, C / 4, b_row_mod, alpha, beta, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return post_launch("axpby_kernel");
}

int ds2_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && y && n > 0, DS2_E_ARG, "ds2_cast_f32_bf16: bad args");
  const long long t = (n + 3) / 4;
  DS2_LAUNCH((cast_f32_bf16_kernel), static_cast<unsigned>((t + 255) / 256), 256, 0, as_stream(stream), 
      x, reinterpret_cast<__nv_bfloat16*>(y), n);
  return post_launch("cast_f32_bf16_kernel");
}

int ds2_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && y && n > 0, DS2_E_ARG, "ds2_cast_bf16_f32: bad args");
  DS2_LAUNCH((cast_bf16_f32_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), y, n);
  return post_launch("cast_bf16_f32_kernel");
}

int ds2_maxpool2x2(const float* x, float* y, int32_t B, int32_t Hm, int32_t Wm, int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && y && B > 0 && (Hm % 2) == 0 && (Wm % 2) == 0 && (C % 4) == 0, DS2_E_ARG,
              "ds2_maxpool2x2: bad args");
  const long long n = static_cast<long long>(B) * (Hm / 2) * (Wm / 2) * (C / 4);
  DS2_LAUNCH((maxpool2x2_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), x, y, B, Hm, Wm, C / 4);
  return post_launch("maxpool2x2_kernel");
}

int ds2_upsample2x_add(const float* top, const float* lat, float* y, int32_t B, int32_t Hm, int32_t Wm,
                       int32_t C, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(top && lat && y && B > 0 && (C % 4) == 0, DS2_E_ARG, "ds2_upsample2x_add: bad args");
  const long long n = static_cast<long long>(B) * Hm * 2 * Wm * 2 * (C / 4);
  DS2_LAUNCH((upsample2x_add_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), top, lat, y, B, Hm,
                                                                                            Wm, C / 4);
  return post_launch("upsample2x_add_kernel");
}

int ds2_bank_gather(const void* mem_bf16, const float* pos, const float* tpos, void* kin_bf16, void* val_bf16,
                    int32_t B, int32_t T, int32_t C, int64_t dst_bs, int32_t row0, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(mem_bf16 && pos && tpos && kin_bf16 && val_bf16 && B > 0 && T > 0 && (C % 8) == 0, DS2_E_ARG,
              "ds2_bank_gather: bad args");
  const long long n = static_cast<long long>(B) * T * (C / 8);
  DS2_LAUNCH((bank_gather_kernel), static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream), 
      reinterpret_cast<const __nv_bfloat16*>(mem_bf16), pos, tpos, reinterpret_cast<__nv_bfloat16*>(kin_bf16),
      reinterpret_cast<__nv_bfloat16*>(val_bf16), B, T, C, dst_bs, row0);
  return post_launch("bank_gather_kernel");
}

int ds2_bank_ptr(const float* ptr, const float* tpos, void* kin_bf16, void* val_bf16, int32_t B, int64_t dst_bs,
                 int32_t row0, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(ptr && tpos && kin_bf16 && val_bf16 && B > 0, DS2_E_ARG, "ds2_bank_ptr: bad args");
  DS2_LAUNCH((bank_ptr_kernel), (B * 256 + 255) / 256, 256, 0, as_stream(stream), 
      ptr, tpos, reinterpret_cast<__nv_bfloat16*>(kin_bf16), reinterpret_cast<__nv_bfloat16*>(val_bf16), B,
      dst_bs, row0);
  return post_launch("bank_ptr_kernel");
}

}  // extern "C"
