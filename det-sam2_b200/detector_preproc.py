"""Detector (YOLOv8) pre-processing on the device — SURVEY.md 8f rank 1, second half.

Det-SAM2 feeds every ``detect_interval``-th frame to ultralytics (det_sam2_inference/det_sam2_RT.py:201-245), which
letterboxes it on the host (``LetterBox``: cv2.resize + grey border to a stride multiple), uploads it and divides by
255.  The same uint8 RGB frames are uploaded for the SAM 2 ingest anyway; ``DeviceLetterbox`` builds the detector's
input tensor from that copy with ``ds2_letterbox_frames`` (byte-exact with cv2) and ``unletterbox`` maps the detector's
boxes back to frame pixels (``ops.scale_boxes``).  The detector itself stays untouched: ultralytics accepts a ready
``[N, 3, H, W]`` float tensor in [0, 1] and then skips its own pre-processing.
"""
import numpy as np
import torch

from . import ops


def letterbox_params(shape_hw, new_shape=(640, 640), auto=True, scaleup=True, stride=32):
    """ultralytics LetterBox geometry for an (h, w) frame -> (new_h, new_w, top, left, out_h, out_w, ratio)."""
    h, w = int(shape_hw[0]), int(shape_hw[1])
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:                                   # minimum rectangle: pad only up to the next stride multiple
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_h, new_w, top, left, new_h + top + bottom, new_w + left + right, r


class DeviceLetterbox:
    """uint8 RGB frames in HBM -> the detector's input tensor, and boxes back to frame pixels."""

    def __init__(self, imgsz=640, half=False, auto=True, stride=32, pad_value=114):
        self.imgsz, self.half, self.auto, self.stride, self.pad_value = imgsz, bool(half), auto, stride, pad_value
        self._lut = {}

    def _table(self, device):
        key = (str(device), self.half)
        if key not in self._lut:
            t = torch.arange(256, dtype=torch.uint8)
            t = t.half() if self.half else t.float()      # BasePredictor.preprocess: cast, then `/= 255`
            t /= 255
            self._lut[key] = t.to(device)
        return self._lut[key]

    def __call__(self, frames_u8):
        """frames_u8: CUDA uint8 [N, H, W, 3] (or [H, W, 3]) RGB -> [N, 3, Hd, Wd] fp32 / fp16 in [0, 1], RGB."""
        if frames_u8.dim() == 3:
            frames_u8 = frames_u8.unsqueeze(0)
        if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8:
            raise ops.capi.Ds2Error("DeviceLetterbox needs uint8 frames on a CUDA device (there is no host path here)")
        N, H, W, _ = frames_u8.shape
        new_h, new_w, top, left, out_h, out_w, _ = letterbox_params((H, W), self.imgsz, self.auto, True, self.stride)
        out = torch.empty((N, 3, out_h, out_w), dtype=torch.float16 if self.half else torch.float32, device=frames_u8.device)
        return ops.letterbox_frames(frames_u8, self._table(frames_u8.device), out, new_h, new_w, top, left, self.pad_value)

    def unletterbox(self, boxes_xyxy, shape_hw):
        """Boxes in the pixels of the letterboxed tensor -> original frame pixels, clipped (ops.scale_boxes)."""
        _, _, top, left, _, _, r = letterbox_params(shape_hw, self.imgsz, self.auto, True, self.stride)
        b = np.array(boxes_xyxy, dtype=np.float32, copy=True).reshape(-1, 4)
        b[:, [0, 2]] = ((b[:, [0, 2]] - left) / r).clip(0, shape_hw[1])
        b[:, [1, 3]] = ((b[:, [1, 3]] - top) / r).clip(0, shape_hw[0])
        return b
