// runtime.cu — library state: error string, launch counter, TMA descriptor encoding.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <string>

#include <stdlib.h>

#include "common.h"

namespace ds2 {

std::atomic<int64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

struct TmapKey {
  uint64_t v[16];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box);
}

int make_tmap(CUtensorMap* out, const void* base, int elt_bytes, int swizzle_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  DS2_REQUIRE(elt_bytes == 2 || elt_bytes == 4, DS2_E_ARG, "tensor map element size %d unsupported", elt_bytes);
  DS2_REQUIRE(swizzle_bytes == 64 || swizzle_bytes == 128, DS2_E_ARG, "tensor map swizzle %d unsupported",
              swizzle_bytes);
  DS2_REQUIRE(enc != nullptr, DS2_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  DS2_REQUIRE(rank >= 2 && rank <= 4, DS2_E_ARG, "tensor map rank %d unsupported", rank);
  DS2_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, DS2_E_ALIGN,
              "TMA base pointer %p not 16-byte aligned", base);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = static_cast<uint64_t>(rank) | (static_cast<uint64_t>(elt_bytes) << 8) |
             (static_cast<uint64_t>(swizzle_bytes) << 16);
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = box[i];
    if (i > 0) {
      key.v[10 + i] = strides_bytes[i - 1];
      DS2_REQUIRE((strides_bytes[i - 1] & 15) == 0, DS2_E_ALIGN,
                  "TMA stride %llu bytes not a multiple of 16",
                  static_cast<unsigned long long>(strides_bytes[i - 1]));
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      return DS2_OK;
    }
  }
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMap m;
  CUresult r = enc(&m, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DS2_REQUIRE(r == CUDA_SUCCESS, DS2_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d",
              static_cast<int>(r));
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() > 8192) g_tmaps.clear();
    g_tmaps.emplace(key, m);
  }
  *out = m;
  return DS2_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DS2_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    n = p.multiProcessorCount;
  }
  return n;
}

}  // namespace ds2

extern "C" {
int ds2_version(void) { return 100; }
int64_t ds2_launch_count(void) { return ds2::g_launches.load(); }
const char* ds2_last_error(void) { return ds2::g_err; }
int ds2_device_sm_count(void) { return ds2::sm_count(); }
}
