// common.h — host-side helpers shared by the kernel translation units of libdetsam2.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/detsam2.h"

namespace ds2 {

extern std::atomic<int64_t> g_launches;
void set_error(const char* fmt, ...);

inline int post_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return DS2_OK;
}

#define DS2_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::ds2::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// 2-D / 3-D bf16 tensor maps with 128-byte swizzle; inner box is always 64 elements (128 bytes).
// dims/strides innermost first; strides in BYTES for dims 1.. (dim 0 is contiguous).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// General form: elt_bytes 2 (bf16) or 4 (f32); swizzle_bytes 64 or 128 (inner box bytes <= swizzle).
int make_tmap(CUtensorMap* out, const void* base, int elt_bytes, int swizzle_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

int sm_count();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and
// executes pdl_sync() before it touches global memory: griddepcontrol.wait blocks until the preceding
// grid has completed and flushed, griddepcontrol.launch_dependents lets the NEXT grid's CTAs be
// scheduled (they run their prologue and then block in their own wait).  The launch latency and the
// prologue (barrier init, TMEM allocation, tensor-map prefetch) of kernel i+1 thereby overlap the tail
// of kernel i — the per-frame path is ~520 short dependent kernels, so the serialised launch gaps were
// a measurable part of the frame.  Inside a stream capture the attribute becomes a programmatic graph
// edge.  DS2_PDL=0 in the environment turns the attribute off (the device-side calls are no-ops then).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_launch();
}

// erf-GELU on the MUFU fast paths: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7) with rcp.approx / ex2.approx —
// ~14 issue slots per element instead of ~50 for libdevice erff.  Total error ~5e-7 absolute (the exact-erf form of
// nn.GELU(), not the tanh approximation).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), hx, hx);
}

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DS2_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::ds2::launch_kernel(kernel, dim3(grid), dim3(block), static_cast<size_t>(smem), stream, __VA_ARGS__)
#endif

}  // namespace ds2
