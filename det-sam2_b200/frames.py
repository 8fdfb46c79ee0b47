"""Frame ingest for the predictor: same input kinds and the same rounding points as the reference's
``load_video_frames`` (/root/reference/sam2/utils/misc.py:236-363): resize to S x S on the host,
/255 in float64, store into an fp16 tensor, then in-place ``-= mean`` and ``/= std`` (each result
rounded to fp16).  Frames stay on the host by default (``offload_video_to_cpu``); when a CUDA engine
is in use the host tensor is pinned so the per-step upload (sam2_video_predictor.py:1184-1186) is an
asynchronous DMA.
"""
import os

import numpy as np
import torch

IMG_MEAN = (0.485, 0.456, 0.406)
IMG_STD = (0.229, 0.224, 0.225)


def _resize_u8(frame_rgb, size):
    import cv2
    return cv2.resize(frame_rgb, (size, size))


def _load_path(path, size):
    """misc.py:93-103: PIL decode, RGB, PIL bilinear-free default resize."""
    from PIL import Image
    pil = Image.open(path)
    a = np.array(pil.convert("RGB").resize((size, size)))
    if a.dtype != np.uint8:
        raise RuntimeError(f"unknown image dtype {a.dtype} in {path}")
    w, h = pil.size
    return a, h, w


def _is_path(p):
    return isinstance(p, (str, bytes, os.PathLike))


def load_video_frames(video_path, image_size, offload_video_to_cpu=True, compute_device=None,
                      img_mean=IMG_MEAN, img_std=IMG_STD, async_loading_frames=False, pin=False):
    """Returns (images fp16 [N,3,S,S], video_height, video_width).

    ``video_path``: a JPEG directory ("<index>.jpg"), a list of image paths, one image path, one RGB
    uint8 ndarray [H,W,3] or a list of them (misc.py:254-303).  ``async_loading_frames`` is accepted
    for signature compatibility; loading is synchronous (the reference's async loader yields the same
    tensor values).
    """
    arrays = None
    paths = None
    if _is_path(video_path) and os.path.isdir(video_path):
        names = [p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG")]
        names.sort(key=lambda p: int(os.path.splitext(p)[0]))
        if not names:
            raise RuntimeError(f"no images found in {video_path}")
        paths = [os.path.join(video_path, n) for n in names]
    elif isinstance(video_path, list) and len(video_path) > 0 and all(_is_path(p) and os.path.isfile(p) for p in video_path):
        paths = list(video_path)
    elif isinstance(video_path, np.ndarray):
        arrays = [video_path]
    elif isinstance(video_path, list) and len(video_path) > 0 and all(isinstance(p, np.ndarray) for p in video_path):
        arrays = video_path
    elif _is_path(video_path) and os.path.isfile(video_path):
        paths = [video_path]
    else:
        raise NotImplementedError(
            "unsupported frame source: pass a JPEG folder, a list of image paths, one image path, "
            "an RGB uint8 ndarray or a list of them")
    n = len(arrays) if arrays is not None else len(paths)
    images = torch.zeros(n, 3, image_size, image_size, dtype=torch.float16)
    if arrays is not None:
        for i, fr in enumerate(arrays):
            a = _resize_u8(fr, image_size) / 255.0
            images[i] = torch.from_numpy(a).permute(2, 0, 1)
        vh, vw = arrays[0].shape[:2]
    else:
        vh = vw = None
        for i, p in enumerate(paths):
            a, vh, vw = _load_path(p, image_size)
            images[i] = torch.from_numpy(a / 255.0).permute(2, 0, 1)
    mean = torch.tensor(img_mean, dtype=torch.float32)[:, None, None]
    std = torch.tensor(img_std, dtype=torch.float32)[:, None, None]
    if not offload_video_to_cpu and compute_device is not None:
        images = images.to(compute_device)
        mean, std = mean.to(compute_device), std.to(compute_device)
    images -= mean
    images /= std
    if pin and images.device.type == "cpu" and torch.cuda.is_available():
        images = images.pin_memory()
    return images, vh, vw
