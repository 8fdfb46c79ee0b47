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

}  // namespace ds2
