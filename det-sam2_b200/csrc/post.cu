// post.cu — post-processing of mask logits: 8-connected component labelling (lock-free union-find
// with atomicMin), hole filling, bilinear resize to video resolution, threshold + bit-pack.
// Integer / byte work, HBM- and latency-bound; grids cover all objects of a frame in one launch
// (the reference launches 6 kernels per object, csrc/connected_components.cu:245-276).
#include <limits.h>

#include "common.h"

namespace ds2 {

__device__ __forceinline__ int uf_find(const int* parent, int n) {
  int p = parent[n];
  while (p != n) {
    n = p;
    p = parent[n];
  }
  return n;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    // a < b : hang b under a if b is still a root
    const int old = atomicMin(parent + b, a);
    if (old == b) return;
    b = old;
  }
}

// fg(p): foreground predicate source is either a uint8 mask (!=0) or f32 scores (<= 0)
__device__ __forceinline__ bool is_fg(const uint8_t* m8, const float* sc, long long i) {
  return m8 ? (m8[i] != 0) : (sc[i] <= 0.f);
}

// Initial parent = left end of the pixel's horizontal run *inside its warp's 32-pixel segment*
// (ballot + clz), so the long horizontal chains of big components never go through union-find.
__global__ void cc_init_kernel(const uint8_t* __restrict__ m8, const float* __restrict__ sc, int* __restrict__ parent,
                               int* __restrict__ area, int* __restrict__ minblk, long long total, int HW, int W) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = i < total;
  const bool fg = valid && is_fg(m8, sc, i);
  const int p = valid ? static_cast<int>(i % HW) : 0;
  const int x = p % W;
  const unsigned fgmask = __ballot_sync(0xffffffffu, fg);
  // linked to the left neighbour: both foreground, same row, same warp segment
  const bool link = fg && lane > 0 && x > 0 && ((fgmask >> (lane - 1)) & 1u);
  const unsigned startmask = ~__ballot_sync(0xffffffffu, link);  // bit set = lane starts a run
  if (!valid) return;
  const unsigned below = startmask & (0xffffffffu >> (31 - lane));  // starts at lanes <= lane (lane 0 always set)
  const int start_lane = 31 - __clz(below);
  parent[i] = fg ? p - (lane - start_lane) : -1;
  area[i] = 0;
  if (minblk) minblk[i] = INT_MAX;
}

// Unions only where a link is not already implied by a neighbour's link (8-connectivity):
//   * lane-0 pixels re-join runs split at warp-segment boundaries;
//   * a pixel whose left neighbour is foreground inherits that neighbour's links to the row above,
//     except the up-right one when the pixel directly above is background;
//   * otherwise: up if foreground, else up-left and up-right individually.
__global__ void cc_merge_kernel(int* __restrict__ parent, long long total, int H, int W) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int HW = H * W;
  const int p = static_cast<int>(i % HW);
  int* par = parent + (i - p);
  if (par[p] < 0) return;
  const int y = p / W, x = p % W;
  const bool left = x > 0 && par[p - 1] >= 0;
  if (left && (threadIdx.x & 31) == 0) uf_union(par, p, p - 1);
  if (y == 0) return;
  const bool up = par[p - W] >= 0;
  const bool upr = x + 1 < W && par[p - W + 1] >= 0;
  if (left) {
    if (upr && !up) uf_union(par, p, p - W + 1);
    return;
  }
  if (up) {
    uf_union(par, p, p - W);
    return;
  }
  if (x > 0 && par[p - W - 1] >= 0) uf_union(par, p, p - W - 1);
  if (upr) uf_union(par, p, p - W + 1);
}

__global__ void cc_count_kernel(const int* __restrict__ parent, int* __restrict__ area, int* __restrict__ minblk,
                                long long total, int H, int W) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int HW = H * W;
  const bool valid = i < total;
  const int p = valid ? static_cast<int>(i % HW) : 0;
  const long long base = i - p;
  const bool fg = valid && parent[i] >= 0;
  const unsigned active = __ballot_sync(0xffffffffu, fg);
  if (!fg) return;
  const int root = uf_find(parent + base, p);
  // warp-aggregated atomics: lanes sharing (image, root) elect one leader — a big component would
  // otherwise serialise tens of thousands of atomics on a single address
  const long long key = base + root;
  const unsigned peers = __match_any_sync(active, key);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(area + key, __popc(peers));
  if (minblk) {
    const int y = p / W, x = p % W;
    atomicMin(minblk + key, (y & ~1) * W + (x & ~1));
  }
}

__global__ void cc_emit_kernel(const int* __restrict__ parent, const int* __restrict__ minblk,
                               int* __restrict__ labels, int* __restrict__ counts, long long total, int HW) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = static_cast<int>(i % HW);
  const long long base = i - p;
  if (parent[i] < 0) {
    labels[i] = 0;
    counts[i] = 0;
    return;
  }
  const int root = uf_find(parent + base, p);
  labels[i] = minblk[base + root] + 1;
  counts[i] = counts[base + root];  // roots keep their own value; non-roots are never read as roots
}

__global__ void fill_holes_kernel(float* __restrict__ scores, const int* __restrict__ parent,
                                  const int* __restrict__ area, long long total, int HW, int max_area) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (parent[i] < 0) return;
  const int p = static_cast<int>(i % HW);
  const long long base = i - p;
  const int root = uf_find(parent + base, p);
  if (area[base + root] <= max_area) scores[i] = 0.1f;
}

// PyTorch upsample_bilinear2d, align_corners=False, antialias=False
// blockIdx.y = output row, blockIdx.z = image; one thread = 4 consecutive output pixels of that row (one 16-byte store
// when the row pitch allows).  The first version spent its time in three 64-bit divisions per pixel and scalar
// stores: 100 us for 16 x 1024^2, against 67 MB / HBM rate ~ 12 us.
template <bool kVec>
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ x, float* __restrict__ y, int Hi, int Wi,
                                                              int Ho, int Wo, float sh, float sw) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (ox0 >= Wo) return;
  const int oy = blockIdx.y;
  const long long n = blockIdx.z;
  const float sy = fmaxf(sh * (oy + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Hi - 1);
  const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0);
  const float ly = fminf(fmaxf(sy - y0, 0.f), 1.f);
  const float hy = 1.f - ly;
  const float* r0 = x + (n * Hi + y0) * Wi;
  const float* r1 = x + (n * Hi + y1) * Wi;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int ox = min(ox0 + e, Wo - 1);
    const float sx = fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
    const int x0 = min(static_cast<int>(sx), Wi - 1);
    const int x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float lx = fminf(fmaxf(sx - x0, 0.f), 1.f);
    const float hx = 1.f - lx;
    v[e] = hy * (hx * __ldg(r0 + x0) + lx * __ldg(r0 + x1)) + ly * (hx * __ldg(r1 + x0) + lx * __ldg(r1 + x1));
  }
  float* dst = y + (n * Ho + oy) * Wo + ox0;
  if (kVec) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));   // streaming: nobody on the device re-reads it
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (ox0 + e < Wo) dst[e] = v[e];
  }
}

__global__ void threshold_pack_kernel(const float* __restrict__ x, uint8_t* __restrict__ bits, long long n) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const long long byte = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long i0 = byte * 8;
  if (i0 >= n) return;
  unsigned v = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (i0 + e < n && x[i0 + e] > 0.f) v |= 1u << e;
  bits[byte] = static_cast<uint8_t>(v);
}

// ---- thresholded masks -> packed bits + per-object area and coordinate sums (integer path) ----------
// One warp per row chunk of 256 pixels: lane l owns pixels [8l, 8l+8) -> one output byte (bit e = pixel 8l+e, the
// numpy.packbits(bitorder="little") convention of ds2_threshold_pack); popcounts and the coordinate sums of the
// set pixels are reduced with shuffles and added to stats[obj] = {area, sum_x, sum_y} with 64-bit integer atomics
// (exact and order-independent, so the result is bit-reproducible).
__global__ void __launch_bounds__(256) mask_pack_stats_kernel(const float* __restrict__ x, uint8_t* __restrict__ bits,
                                                              unsigned long long* __restrict__ stats, int N, int H, int W) {
  pdl_sync();  // griddepcontrol.wait + launch_dependents (programmatic dependent launch)
  const int chunks = (W + 255) / 256;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(N) * H * chunks;
  if (wid >= total) return;
  const int ch = static_cast<int>(wid % chunks);
  const int y = static_cast<int>((wid / chunks) % H);
  const int n = static_cast<int>(wid / (static_cast<long long>(chunks) * H));
  const int x0 = ch * 256 + lane * 8;
  const float* row = x + (static_cast<long long>(n) * H + y) * W;
  unsigned v = 0;
  if (x0 + 7 < W && ((reinterpret_cast<uintptr_t>(row + x0) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row + x0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row + x0) + 1);
    v = (a.x > 0.f) | ((a.y > 0.f) << 1) | ((a.z > 0.f) << 2) | ((a.w > 0.f) << 3) | ((b.x > 0.f) << 4) |
        ((b.y > 0.f) << 5) | ((b.z > 0.f) << 6) | ((b.w > 0.f) << 7);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (x0 + e < W && row[x0 + e] > 0.f) v |= 1u << e;
  }
  if (bits && x0 < W) {
    // rows are packed independently when W is not a multiple of 8 (W/8 rounded up bytes per row)
    bits[(static_cast<long long>(n) * H + y) * ((W + 7) / 8) + (x0 >> 3)] = static_cast<uint8_t>(v);
  }
  unsigned cnt = __popc(v);
  unsigned sx = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) sx += ((v >> e) & 1u) * static_cast<unsigned>(x0 + e);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
  }
  if (lane == 0 && cnt && stats) {
    atomicAdd(stats + 3 * n, static_cast<unsigned long long>(cnt));
    atomicAdd(stats + 3 * n + 1, static_cast<unsigned long long>(sx));
    atomicAdd(stats + 3 * n + 2, static_cast<unsigned long long>(cnt) * static_cast<unsigned long long>(y));
  }
}

static inline unsigned nblocks(long long n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

}  // namespace ds2

extern "C" {

int ds2_connected_components(const uint8_t* mask, int32_t* labels, int32_t* counts, int32_t N, int32_t H, int32_t W,
                             void* stream) {
  using namespace ds2;
  DS2_REQUIRE(mask && labels && counts && N > 0 && H > 0 && W > 0, DS2_E_ARG, "ds2_connected_components: bad args");
  // the reference asserts even H and W (csrc/connected_components.cu:229-232)
  DS2_REQUIRE((H % 2) == 0 && (W % 2) == 0, DS2_E_ARG, "ds2_connected_components: height and width must be even");
  cudaStream_t st = as_stream(stream);
  const long long total = static_cast<long long>(N) * H * W;
  int* tmp = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&tmp), static_cast<size_t>(total) * 2 * sizeof(int), st);
  DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_connected_components: cudaMallocAsync: %s",
              cudaGetErrorString(e));
  int* parent = tmp;
  int* minblk = tmp + total;
  const int HW = H * W;
  DS2_LAUNCH((cc_init_kernel), nblocks(total, 256), 256, 0, st, mask, nullptr, parent, counts, minblk, total, HW, W);
  int rc = post_launch("cc_init_kernel");
  if (!rc) {
    DS2_LAUNCH((cc_merge_kernel), nblocks(total, 256), 256, 0, st, parent, total, H, W);
    rc = post_launch("cc_merge_kernel");
  }
  if (!rc) {
    DS2_LAUNCH((cc_count_kernel), nblocks(total, 256), 256, 0, st, parent, counts, minblk, total, H, W);
    rc = post_launch("cc_count_kernel");
  }
  if (!rc) {
    DS2_LAUNCH((cc_emit_kernel), nblocks(total, 256), 256, 0, st, parent, minblk, labels, counts, total, HW);
    rc = post_launch("cc_emit_kernel");
  }
  cudaFreeAsync(tmp, st);
  return rc;
}

int ds2_fill_holes(float* scores, int32_t* labels_ws, int32_t* counts_ws, int32_t N, int32_t H, int32_t W,
                   int32_t max_area, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(scores && labels_ws && counts_ws && N > 0 && H > 0 && W > 0 && max_area > 0, DS2_E_ARG,
              "ds2_fill_holes: bad args");
  cudaStream_t st = as_stream(stream);
  const long long total = static_cast<long long>(N) * H * W;
  const int HW = H * W;
  DS2_LAUNCH((cc_init_kernel), nblocks(total, 256), 256, 0, st, nullptr, scores, labels_ws, counts_ws, nullptr, total, HW, W);
  int rc = post_launch("cc_init_kernel");
  if (rc) return rc;
  DS2_LAUNCH((cc_merge_kernel), nblocks(total, 256), 256, 0, st, labels_ws, total, H, W);
  rc = post_launch("cc_merge_kernel");
  if (rc) return rc;
  DS2_LAUNCH((cc_count_kernel), nblocks(total, 256), 256, 0, st, labels_ws, counts_ws, nullptr, total, H, W);
  rc = post_launch("cc_count_kernel");
  if (rc) return rc;
  DS2_LAUNCH((fill_holes_kernel), nblocks(total, 256), 256, 0, st, scores, labels_ws, counts_ws, total, HW, max_area);
  return post_launch("fill_holes_kernel");
}

int ds2_resize_bilinear(const float* x, float* y, int32_t N, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                        void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && y && N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, DS2_E_ARG, "ds2_resize_bilinear: bad args");
  const long long total = static_cast<long long>(N) * Ho * Wo;
  DS2_REQUIRE(Ho <= 65535 && N <= 65535, DS2_E_ARG, "ds2_resize_bilinear: Ho %d / N %d above the grid limit", Ho, N);
  const int tx = Wo >= 1024 ? 256 : (Wo >= 256 ? 64 : 32);
  const dim3 grid((Wo + 4 * tx - 1) / (4 * tx), Ho, N);
  const float sh = static_cast<float>(Hi) / Ho, sw = static_cast<float>(Wi) / Wo;
  if ((Wo & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    DS2_LAUNCH((resize_bilinear_kernel<true>), grid, tx, 0, as_stream(stream), x, y, Hi, Wi, Ho, Wo, sh, sw);
  } else {
    DS2_LAUNCH((resize_bilinear_kernel<false>), grid, tx, 0, as_stream(stream), x, y, Hi, Wi, Ho, Wo, sh, sw);
  }
  return post_launch("resize_bilinear_kernel");
}

int ds2_mask_pack_stats(const float* x, uint8_t* bits, uint64_t* stats, int32_t N, int32_t H, int32_t W, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && (bits || stats) && N > 0 && H > 0 && W > 0, DS2_E_ARG, "ds2_mask_pack_stats: bad args");
  cudaStream_t st = as_stream(stream);
  if (stats) {
    cudaError_t e = cudaMemsetAsync(stats, 0, static_cast<size_t>(N) * 3 * sizeof(uint64_t), st);
    DS2_REQUIRE(e == cudaSuccess, static_cast<int>(e), "ds2_mask_pack_stats: memset: %s", cudaGetErrorString(e));
  }
  const long long warps = static_cast<long long>(N) * H * ((W + 255) / 256);
  DS2_LAUNCH((mask_pack_stats_kernel), nblocks(warps * 32, 256), 256, 0, st, x, bits,
             reinterpret_cast<unsigned long long*>(stats), N, H, W);
  return post_launch("mask_pack_stats_kernel");
}

int ds2_threshold_pack(const float* x, uint8_t* bits, int64_t n, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(x && bits && n > 0, DS2_E_ARG, "ds2_threshold_pack: bad args");
  DS2_LAUNCH((threshold_pack_kernel), nblocks((n + 7) / 8, 256), 256, 0, as_stream(stream), x, bits, n);
  return post_launch("threshold_pack_kernel");
}

}  // extern "C"
