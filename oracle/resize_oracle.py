"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the uint8 bilinear resize behind frame ingest.

The reference resizes every frame on the host with ``cv2.resize(frame_rgb, (image_size, image_size))``
(/root/reference/sam2/utils/misc.py:338,345; default interpolation INTER_LINEAR).  The arithmetic lives in a
third-party dependency that is not vendored: opencv-python==4.10.0.84 (requirements.txt:76; this image has cv2 4.13).
Its published algorithm for 8-bit images (modules/imgproc/src/resize.cpp: ``resizeGeneric_`` with
``HResizeLinear`` / ``VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>``) is restated here:

  * source coordinate of destination index d: ``f = float((d + 0.5) * scale - 0.5)`` with ``scale = src/dst`` in
    double; ``s = floor(f)``, ``f -= s`` (float32);
  * columns: ``s < 0`` -> ``s = 0, f = 0``; ``s >= src_w - 1`` -> ``s = src_w - 1, f = 0`` (the right neighbour is then
    never used); rows are only clamped into ``[0, src_h - 1]`` — their weights are kept;
  * weights are 11-bit fixed point: ``w1 = rint(f * 2048)``, ``w0 = rint((1 - f) * 2048)`` (float32 product,
    round-half-even, saturated to int16);
  * horizontal pass in int32: ``h = S[s] * w0 + S[s + 1] * w1``;
  * vertical pass: ``dst = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2``;
  * an exact 2x2 decimation (src = 2 * dst in both axes) is redirected to INTER_AREA: ``(a + b + c + d + 2) >> 2``;
  * equal sizes return a copy.

Pinning: ``tests/test_ingest.py`` checks this restatement bit for bit against cv2 itself (the reference's own
dependency, present in this image and on the GPU box) over a sweep of frame sizes, including 1080p, 720p, odd sizes,
up-scaling and the 2x fast path.  Only tests may import this module.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _axis_tables(src, dst, zero_weight_outside):
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if zero_weight_outside:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
    w0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int32)
    w1 = np.clip(np.rint(f * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int32)
    return s, w0, w1


def resize_u8_bilinear(img, dst_h, dst_w):
    """img uint8 [H, W, C] -> uint8 [dst_h, dst_w, C], equal to cv2.resize(img, (dst_w, dst_h))."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    if (H, W) == (dst_h, dst_w):
        return img.copy()
    if H == 2 * dst_h and W == 2 * dst_w:
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _axis_tables(W, dst_w, True)
    sy, b0, b1 = _axis_tables(H, dst_h, False)
    src = img.astype(np.int32)
    sx1 = np.minimum(sx + 1, W - 1)                    # weight is 0 wherever this clamp acts
    h = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]       # [H, dst_w, C] int32
    r0 = np.clip(sy, 0, H - 1)
    r1 = np.clip(sy + 1, 0, H - 1)
    t0 = (b0[:, None, None] * (h[r0] >> 4)) >> 16
    t1 = (b1[:, None, None] * (h[r1] >> 4)) >> 16
    return np.clip((t0 + t1 + 2) >> 2, 0, 255).astype(np.uint8)
