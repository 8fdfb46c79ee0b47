"""Frame ingest for the predictor: same input kinds and the same rounding points as the reference's
``load_video_frames`` (/root/reference/sam2/utils/misc.py:236-363): resize to S x S on the host,
/255 in float64, store into an fp16 tensor, then in-place ``-= mean`` and ``/= std`` (each result
rounded to fp16).  Frames stay on the host by default (``offload_video_to_cpu``); when a CUDA engine
is in use the host tensor is pinned so the per-step upload (sam2_video_predictor.py:1184-1186) is an
asynchronous DMA.

With a CUDA compute device, ndarray frames are ingested ON THE DEVICE (``ds2_ingest_frames``, csrc/ingest.cu): the raw
uint8 frames are staged through pinned memory, resized with OpenCV's exact 8-bit bilinear arithmetic and normalised
through the same 3 x 256 table in one kernel, and the fp16 result is either kept in HBM
(``offload_video_to_cpu=False``) or copied back into the pinned host tensor — bit-identical to the host path
(tests/test_ingest.py) at a fraction of its time.  ``DS2_HOST_INGEST=1`` forces the host path.
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
                      img_mean=IMG_MEAN, img_std=IMG_STD, async_loading_frames=False, pin=False, device_ingest=None):
    """Returns (images fp16 [N,3,S,S], video_height, video_width).

    ``video_path``: a JPEG directory ("<index>.jpg"), a list of image paths, one image path, one RGB
    uint8 ndarray [H,W,3] or a list of them (misc.py:254-303).  ``async_loading_frames`` is accepted
    for signature compatibility; loading is synchronous (the reference's async loader yields the same
    tensor values).  ``device_ingest``: None = on the device whenever ``compute_device`` is CUDA and the frames are
    ndarrays (unless DS2_HOST_INGEST=1), False = host, True = device (raises without CUDA).
    """
    arrays = None
    paths = None
    if isinstance(video_path, torch.Tensor):
        # addition: uint8 RGB frames [N, H, W, 3] ALREADY in HBM (VideoProcessor uploads a chunk once and shares it
        # between the detector pre-processing and this ingest): straight into ds2_ingest_frames, no host staging
        t = video_path
        if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 3):
            raise NotImplementedError("tensor frame sources must be CUDA uint8 [N, H, W, 3] RGB")
        from . import ops
        device = t.device
        lut = _normalize_lut(tuple(img_mean), tuple(img_std))
        key = (str(device), lut.tobytes())
        dev_lut = _DEV_LUTS.get(key)
        if dev_lut is None:
            dev_lut = _DEV_LUTS[key] = torch.from_numpy(lut.view(np.int16).copy()).to(device)
        dst = torch.empty(t.shape[0], 3, image_size, image_size, dtype=torch.float16, device=device)
        ops.ingest_frames(t.contiguous(), dev_lut, dst)
        if offload_video_to_cpu:
            images = torch.empty(dst.shape, dtype=torch.float16, pin_memory=True)
            images.copy_(dst, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            return images, int(t.shape[1]), int(t.shape[2])
        return dst, int(t.shape[1]), int(t.shape[2])
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
    to_device = not offload_video_to_cpu and compute_device is not None
    pinned = (pin or to_device) and torch.cuda.is_available()
    on_cuda = compute_device is not None and torch.device(compute_device).type == "cuda"
    if device_ingest is None:
        device_ingest = on_cuda and arrays is not None and os.environ.get("DS2_HOST_INGEST", "0") != "1"
    if device_ingest:
        if arrays is None or not on_cuda:
            raise RuntimeError("device ingest needs ndarray frames and a CUDA compute device")
        lut = _normalize_lut(tuple(img_mean), tuple(img_std))
        images = _ingest_on_device(arrays, image_size, lut, torch.device(compute_device), keep_on_device=to_device,
                                   pin=pinned)
        vh, vw = arrays[0].shape[:2]
        return images, vh, vw
    images = torch.empty(n, 3, image_size, image_size, dtype=torch.float16, pin_memory=pinned)
    # Every pixel goes through the same three roundings as in the reference (u8/255 in float64 -> fp16, `-= mean`
    # and `/= std` each rounded to fp16), but there are only 3 x 256 distinct inputs: the per-channel table is
    # built ONCE with exactly the reference's tensor operations and the frame is a gather through it — bit-identical
    # to the arithmetic path (tests/test_frames.py) at ~1/8 of its host time (float64 division and fp16 CPU
    # arithmetic over 3 M values per frame were a third of the wall time of the streaming driver).
    lut = _normalize_lut(tuple(img_mean), tuple(img_std))
    out16 = images.numpy().view(np.uint16)
    if arrays is not None:
        for i, fr in enumerate(arrays):
            _gather_frame(_resize_u8(fr, image_size), lut, out16[i])
        vh, vw = arrays[0].shape[:2]
    else:
        vh = vw = None
        for i, p in enumerate(paths):
            a, vh, vw = _load_path(p, image_size)
            _gather_frame(a, lut, out16[i])
    if to_device:
        images = images.to(compute_device, non_blocking=True)
        torch.cuda.current_stream(compute_device).synchronize()
    return images, vh, vw


_STAGE = {}
_DEV_LUTS = {}
_INGEST_BATCH = 16     # frames per staging round: bounds the pinned staging buffer (16 x 1080p = 100 MB)


def _check_frame(fr):
    if not isinstance(fr, np.ndarray) or fr.dtype != np.uint8 or fr.ndim != 3 or fr.shape[2] != 3:
        raise RuntimeError(f"expected an RGB uint8 frame [H,W,3], got "
                           f"{getattr(fr, 'dtype', type(fr))} {getattr(fr, 'shape', '')}")


def _ingest_on_device(arrays, image_size, lut, device, keep_on_device, pin):
    """uint8 RGB ndarrays -> fp16 [N,3,S,S] through ds2_ingest_frames.  Frames go through a re-used pinned staging
    buffer in rounds of ``_INGEST_BATCH``; consecutive frames of one size share a launch.  The result stays on the
    device (``keep_on_device``) or is copied back into a (pinned) host tensor, one stream synchronisation in total."""
    from . import ops
    n = len(arrays)
    for fr in arrays:
        _check_frame(fr)
    stream = torch.cuda.current_stream(device)
    key = (str(device), lut.tobytes())
    dev_lut = _DEV_LUTS.get(key)
    if dev_lut is None:     # fp16 bit patterns travel as int16 (torch has no arithmetic-free uint16 path on every op)
        dev_lut = _DEV_LUTS[key] = torch.from_numpy(lut.view(np.int16).copy()).to(device)
    with torch.cuda.device(device):
        dst = torch.empty(n, 3, image_size, image_size, dtype=torch.float16, device=device)
        i = 0
        rounds = []
        while i < n:
            shape = arrays[i].shape
            j = i
            while j < n and j - i < _INGEST_BATCH and arrays[j].shape == shape:
                j += 1
            nbytes = (j - i) * shape[0] * shape[1] * 3
            # two staging buffers alternate so that filling round r+1 overlaps the DMA of round r
            slot = len(rounds) & 1
            stage = _STAGE.get(slot)
            if stage is None or stage.numel() < nbytes:
                stage = _STAGE[slot] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
            if len(rounds) >= 2:
                rounds[-2].synchronize()       # the DMA that last read this slot
            view = stage[:nbytes].view(j - i, shape[0], shape[1], 3)
            host = view.numpy()
            for k in range(i, j):
                np.copyto(host[k - i], arrays[k])
            dev_u8 = view.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            rounds.append(ev)
            ops.ingest_frames(dev_u8, dev_lut, dst[i:j])
            i = j
        if keep_on_device:
            stream.synchronize()
            return dst
        # always pinned: a pageable destination would turn the copy into a synchronous bounce through the driver's
        # staging buffer (torch's host allocator re-uses the block once the previous chunk's tensor is gone)
        try:
            images = torch.empty(n, 3, image_size, image_size, dtype=torch.float16, pin_memory=True)
        except RuntimeError:   # host cannot lock that much memory: a pageable destination is only slower
            images = torch.empty(n, 3, image_size, image_size, dtype=torch.float16)
        images.copy_(dst, non_blocking=True)
        stream.synchronize()
    return images


def normalize_frames_arithmetic(frames_u8, img_mean=IMG_MEAN, img_std=IMG_STD):
    """The reference's arithmetic (misc.py:336-359) on already-resized uint8 frames [N,S,S,3] -> fp16 [N,3,S,S]:
    /255 in float64, store to fp16, in-place `-= mean`, `/= std`.  Used to build the gather table and as the
    ground truth of its test."""
    frames_u8 = np.asarray(frames_u8)
    images = torch.zeros(frames_u8.shape[0], 3, frames_u8.shape[1], frames_u8.shape[2], dtype=torch.float16)
    for i in range(frames_u8.shape[0]):
        images[i] = torch.from_numpy(frames_u8[i] / 255.0).permute(2, 0, 1)
    images -= torch.tensor(img_mean, dtype=torch.float32)[:, None, None]
    images /= torch.tensor(img_std, dtype=torch.float32)[:, None, None]
    return images


_LUTS = {}
_SCRATCH = {}


def _normalize_lut(img_mean, img_std):
    """uint16 [3, 256]: fp16 bit pattern of the normalised value of byte v in channel c."""
    key = (img_mean, img_std)
    if key not in _LUTS:
        ramp = np.repeat(np.arange(256, dtype=np.uint8).reshape(1, 256, 1, 1), 3, axis=3)   # [1, 256, 1, 3]
        t = normalize_frames_arithmetic(ramp, img_mean, img_std)                            # [1, 3, 256, 1]
        _LUTS[key] = np.ascontiguousarray(t.numpy().view(np.uint16).reshape(3, 256))
    return _LUTS[key]


def _gather_frame(frame_u8, lut, out16):
    """frame_u8 [S,S,3] uint8 -> out16 [3,S,S] (uint16 view of the fp16 destination)."""
    if frame_u8.dtype != np.uint8 or frame_u8.ndim != 3 or frame_u8.shape[2] != 3:
        raise RuntimeError(f"expected an RGB uint8 frame [H,W,3], got {frame_u8.dtype} {frame_u8.shape}")
    import cv2
    # de-interleave into a re-used scratch (fresh outputs cost 10x more in page faults than the split itself)
    key = frame_u8.shape[:2]
    planes = _SCRATCH.get(key)
    if planes is None:
        planes = _SCRATCH[key] = [np.empty(key, dtype=np.uint8) for _ in range(3)]
        if len(_SCRATCH) > 4:
            _SCRATCH.pop(next(iter(_SCRATCH)))
    cv2.split(np.ascontiguousarray(frame_u8), planes)
    for c in range(3):
        cv2.LUT(planes[c], lut[c], dst=out16[c])   # vectorised byte -> 16-bit table lookup, written in place
