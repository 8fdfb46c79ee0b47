// ingest.cu — frame ingest on the device.
//
// Replaces, per frame, the host work of /root/reference/sam2/utils/misc.py:336-359:
//   cv2.resize(frame_rgb, (S, S)) / 255.0 -> fp16 -> `-= mean` -> `/= std`
// with one launch over uint8 RGB frames already in HBM.  Byte/integer work, HBM-bound: 3 B read per source
// pixel touched, 6 B written per destination pixel.
//
//  * The resize is OpenCV's 8-bit bilinear (imgproc/src/resize.cpp, opencv-python==4.10.0.84 in the reference's
//    requirements.txt:76), restated in oracle/resize_oracle.py and reproduced here bit for bit: float32 source
//    coordinates from a double product, 11-bit fixed-point weights (round-half-even), int32 horizontal pass,
//    `(((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2` vertical pass, the exact 2x decimation
//    redirected to a 2x2 box mean, equal sizes copied.  Every floating-point step uses the explicitly rounded
//    intrinsics (__dmul_rn, __fsub_rn, ...) so that the compiler cannot contract them into FMAs.
//  * /255, the fp16 store and the two in-place fp16 normalisation steps have only 3 x 256 distinct inputs: the host
//    builds that table once with the reference's own tensor operations (frames.py:_normalize_lut) and the kernel
//    gathers through it from shared memory — identical to the arithmetic path by construction.
//
// Layout: src uint8 [N][Hv][Wv][3] (row pitch / frame stride in bytes), dst fp16 [N][3][S][S] planar (the
// reference's `images` tensor).  One CTA = 128 x 8 destination pixels; a thread owns two adjacent columns so a
// warp writes 128 contiguous bytes per plane row.
#include "common.h"

namespace ds2 {

constexpr int kIngTileX = 128;
constexpr int kIngTileY = 8;
constexpr int kIngThreads = 256;

__device__ __forceinline__ int ing_sat16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// source index and the two 11-bit weights of destination index d (resize.cpp: resizeGeneric_ table set-up)
__device__ __forceinline__ void ing_axis(int d, int src, int dst, bool zero_outside, int& s, int& w0, int& w1) {
  const double scale = __ddiv_rn(static_cast<double>(src), static_cast<double>(dst));
  float f = __double2float_rn(__dadd_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), -0.5));
  int si = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(si));
  if (zero_outside) {
    if (si < 0) {
      f = 0.f;
      si = 0;
    }
    if (si >= src - 1) {
      f = 0.f;
      si = src - 1;
    }
  }
  s = si;
  w0 = ing_sat16(__float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f)));
  w1 = ing_sat16(__float2int_rn(__fmul_rn(f, 2048.0f)));
}

__global__ void __launch_bounds__(kIngThreads)
ingest_u8_kernel(const uint8_t* __restrict__ src, int Hv, int Wv, long long pitch, long long frame_stride,
                 const uint16_t* __restrict__ lut, uint16_t* __restrict__ dst, int S) {
  __shared__ uint16_t s_lut[3 * 256];
  __shared__ int s_x[kIngTileX], s_a0[kIngTileX], s_a1[kIngTileX];
  __shared__ int s_r0[kIngTileY], s_r1[kIngTileY], s_b0[kIngTileY], s_b1[kIngTileY];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kIngTileX, y0 = blockIdx.y * kIngTileY;
  const bool same = (Hv == S && Wv == S);
  const bool area2 = (Hv == 2 * S && Wv == 2 * S);
  // coefficient tables of this tile: pure functions of the sizes, computed before waiting for the producer
  if (!same && !area2) {
    if (tid < kIngTileX) {
      int s, a0, a1;
      ing_axis(min(x0 + tid, S - 1), Wv, S, true, s, a0, a1);
      s_x[tid] = s;
      s_a0[tid] = a0;
      s_a1[tid] = a1;
    } else if (tid < kIngTileX + kIngTileY) {
      const int j = tid - kIngTileX;
      int s, b0, b1;
      ing_axis(min(y0 + j, S - 1), Hv, S, false, s, b0, b1);
      s_r0[j] = min(max(s, 0), Hv - 1);
      s_r1[j] = min(max(s + 1, 0), Hv - 1);
      s_b0[j] = b0;
      s_b1[j] = b1;
    }
  }
  pdl_sync();
  for (int i = tid; i < 3 * 256; i += kIngThreads) s_lut[i] = __ldg(lut + i);
  __syncthreads();

  const uint8_t* frame = src + static_cast<long long>(blockIdx.z) * frame_stride;
  uint16_t* out = dst + static_cast<long long>(blockIdx.z) * 3 * S * S;
  const int lx = (tid & 63) * 2;
  const bool vec_ok = (S & 1) == 0;  // element index of (c, y, x) is even for even x
  for (int ly = tid >> 6; ly < kIngTileY; ly += kIngThreads / 64) {
    const int y = y0 + ly;
    if (y >= S) break;
    uint16_t px[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int x = x0 + lx + i;
      if (x >= S) {
        px[i][0] = px[i][1] = px[i][2] = 0;
        continue;
      }
      int v[3];
      if (same) {
        const uint8_t* p = frame + y * pitch + 3ll * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(p + c);
      } else if (area2) {
        const uint8_t* p = frame + (2ll * y) * pitch + 6ll * x;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[c] = (static_cast<int>(__ldg(p + c)) + __ldg(p + 3 + c) + __ldg(p + pitch + c) + __ldg(p + pitch + 3 + c) + 2) >> 2;
      } else {
        const int sx = s_x[lx + i], a0 = s_a0[lx + i], a1 = s_a1[lx + i];
        const int sx1 = min(sx + 1, Wv - 1);  // weight a1 is 0 wherever the clamp acts
        const uint8_t* r0 = frame + s_r0[ly] * pitch;
        const uint8_t* r1 = frame + s_r1[ly] * pitch;
        const int b0 = s_b0[ly], b1 = s_b1[ly];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = static_cast<int>(__ldg(r0 + 3ll * sx + c)) * a0 + static_cast<int>(__ldg(r0 + 3ll * sx1 + c)) * a1;
          const int h1 = static_cast<int>(__ldg(r1 + 3ll * sx + c)) * a0 + static_cast<int>(__ldg(r1 + 3ll * sx1 + c)) * a1;
          const int t = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
          v[c] = min(max(t, 0), 255);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) px[i][c] = s_lut[c * 256 + v[c]];
    }
    const int x = x0 + lx;
    if (x >= S) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint16_t* o = out + (static_cast<long long>(c) * S + y) * S + x;
      if (vec_ok && x + 1 < S) {
        *reinterpret_cast<uint32_t*>(o) = static_cast<uint32_t>(px[0][c]) | (static_cast<uint32_t>(px[1][c]) << 16);
      } else {
        o[0] = px[0][c];
        if (x + 1 < S) o[1] = px[1][c];
      }
    }
  }
}


// ---- detector pre-processing on the device (second half) --------------------------------------
// Det-SAM2 hands every detect_interval-th frame to YOLOv8 (det_sam2_RT.py:201-265); ultralytics (third party, pinned
// 8.2.82, requirements.txt:134, not vendored) then letterboxes it on the HOST — cv2.resize(INTER_LINEAR) to the size that
// fits imgsz at the frame's aspect ratio, cv2.copyMakeBorder with grey 114 up to the next multiple of the model stride —
// stacks, flips BGR->RGB, uploads, casts and divides by 255 (data/augment.py LetterBox.__call__, engine/predictor.py
// preprocess).  The uint8 RGB frame is already in HBM for the SAM 2 ingest, so the same byte-exact bilinear arithmetic
// produces the detector's input tensor there: planar RGB [N][3][Hd][Wd], fp32 or fp16, values v / 255 through a
// 256-entry table built by the host with torch's own arithmetic.  No second upload, no host resize.
template <typename T>
__global__ void __launch_bounds__(kIngThreads)
letterbox_u8_kernel(const uint8_t* __restrict__ src, int Hv, int Wv, long long pitch, long long frame_stride,
                    const T* __restrict__ lut, T* __restrict__ dst, int Hd, int Wd, int new_h, int new_w, int top,
                    int left, int pad) {
  __shared__ T s_lut[256];
  __shared__ int s_x[kIngTileX], s_a0[kIngTileX], s_a1[kIngTileX];
  __shared__ int s_r0[kIngTileY], s_r1[kIngTileY], s_b0[kIngTileY], s_b1[kIngTileY];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kIngTileX, y0 = blockIdx.y * kIngTileY;
  const bool same = (Hv == new_h && Wv == new_w);          // LetterBox skips the resize then
  const bool area2 = (Hv == 2 * new_h && Wv == 2 * new_w);  // cv2 turns the exact 2x INTER_LINEAR case into a box mean
  if (!same && !area2) {
    if (tid < kIngTileX) {
      int s, a0, a1;
      ing_axis(min(max(x0 + tid - left, 0), new_w - 1), Wv, new_w, true, s, a0, a1);
      s_x[tid] = s;
      s_a0[tid] = a0;
      s_a1[tid] = a1;
    } else if (tid < kIngTileX + kIngTileY) {
      const int j = tid - kIngTileX;
      int s, b0, b1;
      ing_axis(min(max(y0 + j - top, 0), new_h - 1), Hv, new_h, false, s, b0, b1);
      s_r0[j] = min(max(s, 0), Hv - 1);
      s_r1[j] = min(max(s + 1, 0), Hv - 1);
      s_b0[j] = b0;
      s_b1[j] = b1;
    }
  }
  pdl_sync();
  for (int i = tid; i < 256; i += kIngThreads) s_lut[i] = lut[i];
  __syncthreads();
  const uint8_t* frame = src + static_cast<long long>(blockIdx.z) * frame_stride;
  T* out = dst + static_cast<long long>(blockIdx.z) * 3 * Hd * Wd;
  const int lx = tid & (kIngTileX - 1);
  for (int ly = tid / kIngTileX; ly < kIngTileY; ly += kIngThreads / kIngTileX) {
    const int y = y0 + ly, x = x0 + lx;
    if (y >= Hd || x >= Wd) continue;
    int v[3] = {pad, pad, pad};
    const int ry = y - top, rx = x - left;
    if (ry >= 0 && ry < new_h && rx >= 0 && rx < new_w) {
      if (same) {
        const uint8_t* p = frame + ry * pitch + 3ll * rx;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(p + c);
      } else if (area2) {
        const uint8_t* p = frame + (2ll * ry) * pitch + 6ll * rx;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[c] = (static_cast<int>(__ldg(p + c)) + __ldg(p + 3 + c) + __ldg(p + pitch + c) + __ldg(p + pitch + 3 + c) + 2) >> 2;
      } else {
        const int sx = s_x[lx], a0 = s_a0[lx], a1 = s_a1[lx];
        const int sx1 = min(sx + 1, Wv - 1);
        const uint8_t* r0 = frame + s_r0[ly] * pitch;
        const uint8_t* r1 = frame + s_r1[ly] * pitch;
        const int b0 = s_b0[ly], b1 = s_b1[ly];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = static_cast<int>(__ldg(r0 + 3ll * sx + c)) * a0 + static_cast<int>(__ldg(r0 + 3ll * sx1 + c)) * a1;
          const int h1 = static_cast<int>(__ldg(r1 + 3ll * sx + c)) * a0 + static_cast<int>(__ldg(r1 + 3ll * sx1 + c)) * a1;
          const int t = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
          v[c] = min(max(t, 0), 255);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(static_cast<long long>(c) * Hd + y) * Wd + x] = s_lut[v[c]];
  }
}

}  // namespace ds2

extern "C" int ds2_ingest_frames(const uint8_t* src_u8, int32_t N, int32_t Hv, int32_t Wv, int64_t pitch_bytes,
                                 int64_t frame_stride_bytes, const uint16_t* lut_3x256, void* dst_f16, int32_t S,
                                 void* stream) {
  using namespace ds2;
  DS2_REQUIRE(src_u8 && lut_3x256 && dst_f16, DS2_E_ARG, "ds2_ingest_frames: null pointer");
  DS2_REQUIRE(N > 0 && Hv > 0 && Wv > 0 && S > 0, DS2_E_ARG, "ds2_ingest_frames: bad shape N=%d Hv=%d Wv=%d S=%d", N,
              Hv, Wv, S);
  DS2_REQUIRE(pitch_bytes >= 3ll * Wv && frame_stride_bytes >= pitch_bytes * Hv, DS2_E_ARG,
              "ds2_ingest_frames: pitch %lld / frame stride %lld too small for %d x %d RGB",
              static_cast<long long>(pitch_bytes), static_cast<long long>(frame_stride_bytes), Hv, Wv);
  DS2_REQUIRE((reinterpret_cast<uintptr_t>(dst_f16) & 3) == 0, DS2_E_ALIGN, "ds2_ingest_frames: dst must be 4-byte aligned");
  DS2_REQUIRE(N <= 65535, DS2_E_ARG, "ds2_ingest_frames: at most 65535 frames per call");
  dim3 grid((S + kIngTileX - 1) / kIngTileX, (S + kIngTileY - 1) / kIngTileY, N);
  DS2_LAUNCH((ingest_u8_kernel), grid, kIngThreads, 0, as_stream(stream), src_u8, Hv, Wv,
             static_cast<long long>(pitch_bytes), static_cast<long long>(frame_stride_bytes), lut_3x256,
             reinterpret_cast<uint16_t*>(dst_f16), S);
  return post_launch("ingest_u8_kernel");
}

extern "C" int ds2_letterbox_frames(const uint8_t* src_u8, int32_t N, int32_t Hv, int32_t Wv, int64_t pitch_bytes,
                                    int64_t frame_stride_bytes, const void* lut_256, void* dst, int32_t dst_is_f32,
                                    int32_t Hd, int32_t Wd, int32_t new_h, int32_t new_w, int32_t top, int32_t left,
                                    int32_t pad_value, void* stream) {
  using namespace ds2;
  DS2_REQUIRE(src_u8 && lut_256 && dst, DS2_E_ARG, "ds2_letterbox_frames: null pointer");
  DS2_REQUIRE(N > 0 && N <= 65535 && Hv > 0 && Wv > 0 && Hd > 0 && Wd > 0 && new_h > 0 && new_w > 0, DS2_E_ARG,
              "ds2_letterbox_frames: bad shape");
  DS2_REQUIRE(top >= 0 && left >= 0 && top + new_h <= Hd && left + new_w <= Wd && pad_value >= 0 && pad_value <= 255,
              DS2_E_ARG, "ds2_letterbox_frames: the resized frame (%d x %d at %d, %d) does not fit the %d x %d output", new_w,
              new_h, left, top, Wd, Hd);
  DS2_REQUIRE(pitch_bytes >= 3ll * Wv && frame_stride_bytes >= pitch_bytes * Hv, DS2_E_ARG,
              "ds2_letterbox_frames: pitch / frame stride too small");
  dim3 grid((Wd + kIngTileX - 1) / kIngTileX, (Hd + kIngTileY - 1) / kIngTileY, N);
  if (dst_is_f32) {
    DS2_LAUNCH((letterbox_u8_kernel<float>), grid, kIngThreads, 0, as_stream(stream), src_u8, Hv, Wv,
               static_cast<long long>(pitch_bytes), static_cast<long long>(frame_stride_bytes),
               reinterpret_cast<const float*>(lut_256), reinterpret_cast<float*>(dst), Hd, Wd, new_h, new_w, top, left, pad_value);
  } else {
    DS2_LAUNCH((letterbox_u8_kernel<uint16_t>), grid, kIngThreads, 0, as_stream(stream), src_u8, Hv, Wv,
               static_cast<long long>(pitch_bytes), static_cast<long long>(frame_stride_bytes),
               reinterpret_cast<const uint16_t*>(lut_256), reinterpret_cast<uint16_t*>(dst), Hd, Wd, new_h, new_w, top, left,
               pad_value);
  }
  return post_launch("letterbox_u8_kernel");
}
