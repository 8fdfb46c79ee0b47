"""TEST INFRASTRUCTURE — CPU restatement of the detector pre-processing Det-SAM2 inherits from ultralytics when it calls
``self.detect_model(frames_bgr, ...)`` (det_sam2_inference/det_sam2_RT.py:229-230).

The algorithm lives in a THIRD-PARTY dependency that is absent from /root/reference and from this image: ultralytics,
pinned 8.2.82 (requirements.txt:134).  Restated from its published source:
  * ``ultralytics/data/augment.py`` ``LetterBox.__call__`` (new_shape = imgsz, auto = True for a PyTorch model fed
    same-shaped images, scaleFill False, scaleup True, center True, stride 32): r = min(h_new / h, w_new / w);
    new_unpad = (round(w r), round(h r)); (dw, dh) = remaining border, reduced mod stride when auto, halved;
    cv2.resize(..., INTER_LINEAR) unless the size is unchanged; cv2.copyMakeBorder with round(d -+ 0.1) and (114,114,114);
  * ``ultralytics/engine/predictor.py`` ``BasePredictor.preprocess``: stack, BGR -> RGB, HWC -> CHW, to device,
    ``.half()`` or ``.float()``, ``/= 255``.
PARITY UNPINNED against ultralytics itself (not installable here); the arithmetic it delegates to — OpenCV's 8-bit
bilinear resize and border copy — is executed by cv2 itself below, and ds2_letterbox_frames is compared bit for bit.
"""
import numpy as np
import torch


def letterbox_params(shape_hw, new_shape=(640, 640), auto=True, scaleup=True, stride=32):
    """-> (new_h, new_w, top, left, out_h, out_w, ratio) of LetterBox for an (h, w) frame."""
    h, w = shape_hw
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_h, new_w, top, left, new_h + top + bottom, new_w + left + right, r


def letterbox_bgr(img_bgr, new_shape=(640, 640), auto=True, scaleup=True, stride=32):
    """LetterBox.__call__ on one BGR uint8 frame (the form YOLO is fed by Det-SAM2)."""
    import cv2
    new_h, new_w, top, left, out_h, out_w, _ = letterbox_params(img_bgr.shape[:2], new_shape, auto, scaleup, stride)
    if img_bgr.shape[:2] != (new_h, new_w):
        img_bgr = cv2.resize(img_bgr, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img_bgr, top, out_h - new_h - top, left, out_w - new_w - left, cv2.BORDER_CONSTANT,
                              value=(114, 114, 114))


def preprocess(frames_rgb, new_shape=(640, 640), half=False, auto=True, stride=32):
    """What the detector's network sees for a list of same-sized RGB uint8 frames (Det-SAM2 converts them to BGR first,
    det_sam2_RT.py:221; the predictor flips them back): float tensor [N, 3, H, W] in [0, 1], RGB."""
    im = np.stack([letterbox_bgr(np.ascontiguousarray(f[..., ::-1]), new_shape, auto=auto, stride=stride) for f in frames_rgb])
    im = np.ascontiguousarray(im[..., ::-1].transpose((0, 3, 1, 2)))
    t = torch.from_numpy(im)
    t = t.half() if half else t.float()
    t /= 255
    return t


def unletterbox_boxes(boxes_xyxy, shape_hw, new_shape=(640, 640), auto=True, stride=32):
    """ops.scale_boxes: boxes in letterboxed-tensor pixels -> original frame pixels (clipped)."""
    new_h, new_w, top, left, _, _, r = letterbox_params(shape_hw, new_shape, auto, True, stride)
    b = np.array(boxes_xyxy, dtype=np.float32, copy=True).reshape(-1, 4)
    b[:, [0, 2]] = (b[:, [0, 2]] - left) / r
    b[:, [1, 3]] = (b[:, [1, 3]] - top) / r
    b[:, [0, 2]] = b[:, [0, 2]].clip(0, shape_hw[1])
    b[:, [1, 3]] = b[:, [1, 3]].clip(0, shape_hw[0])
    return b
