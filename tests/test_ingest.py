"""Frame ingest on the device (csrc/ingest.cu).

CPU part: the numpy restatement of OpenCV's 8-bit bilinear resize (oracle/resize_oracle.py) is pinned bit for bit
against cv2 itself — the reference's own dependency for this step (misc.py:338,345) — over a sweep of frame sizes.
GPU part (-m gpu, through the C ABI): ``ds2_ingest_frames`` against that oracle + the reference's normalisation
arithmetic, and ``load_video_frames`` with device ingest against the host path.  Integer / byte work: bit exact.
"""
import numpy as np
import pytest
import torch

from detsam2_b200.frames import IMG_MEAN, IMG_STD, load_video_frames, normalize_frames_arithmetic, _normalize_lut
from oracle.resize_oracle import resize_u8_bilinear

# (Hv, Wv): 1080p, 720p, the 2x decimation fast path, equal size, odd sizes, up-scaling, one axis at 2x only, tiny
SIZES = [(1080, 1920), (720, 1280), (512, 512), (256, 256), (333, 517), (750, 350), (100, 100), (257, 255),
         (512, 256), (256, 512), (3, 5), (1, 1), (2, 2)]


def _frame(rng, h, w):
    f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    f[: min(h, 16), : min(w, 16)] = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)[: min(h, 16), : min(w, 16)]
    return f


@pytest.mark.parametrize("S", [256, 128, 63])
def test_resize_oracle_is_bit_identical_to_cv2(S):
    import cv2
    rng = np.random.default_rng(11)
    for h, w in SIZES:
        f = _frame(rng, h, w)
        ref = cv2.resize(f, (S, S))
        got = resize_u8_bilinear(f, S, S)
        assert np.array_equal(ref, got), (h, w, S)


def test_resize_oracle_full_size_1080p_to_1024():
    import cv2
    rng = np.random.default_rng(12)
    for h, w in [(1080, 1920), (720, 1280), (2048, 2048), (1024, 1024)]:
        f = _frame(rng, h, w)
        assert np.array_equal(cv2.resize(f, (1024, 1024)), resize_u8_bilinear(f, 1024, 1024)), (h, w)


def test_device_ingest_refuses_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        load_video_frames([np.zeros((8, 8, 3), np.uint8)], 8, compute_device=torch.device("cpu"), device_ingest=True)


# ------------------------------------------------------------------------------------------------------------ GPU
def _expected(frames, S):
    return normalize_frames_arithmetic(np.stack([resize_u8_bilinear(f, S, S) for f in frames]), IMG_MEAN, IMG_STD)


def _dev_lut():
    lut = _normalize_lut(tuple(IMG_MEAN), tuple(IMG_STD))
    return torch.from_numpy(lut.view(np.int16).copy()).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("S", [256, 63])
def test_ingest_kernel_bit_exact_size_sweep(S):
    from detsam2_b200 import ops
    rng = np.random.default_rng(21)
    lut = _dev_lut()
    for h, w in SIZES:
        frames = [_frame(rng, h, w) for _ in range(2)]
        src = torch.from_numpy(np.stack(frames)).cuda()
        out = torch.full((2, 3, S, S), float("nan"), dtype=torch.float16, device="cuda")
        ops.ingest_frames(src, lut, out)
        ref = _expected(frames, S)
        assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16)), (h, w, S)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(1080, 1920), (720, 1280), (2048, 2048), (1024, 1024), (1500, 700)])
def test_ingest_kernel_bit_exact_full_size(h, w):
    """BASELINE sizes: 1024^2 destination from 1080p / 720p (configs[2]) / 2x / equal-size frames."""
    import cv2
    from detsam2_b200 import ops
    rng = np.random.default_rng(22)
    f = _frame(rng, h, w)
    out = torch.empty((1, 3, 1024, 1024), dtype=torch.float16, device="cuda")
    ops.ingest_frames(torch.from_numpy(f[None]).cuda(), _dev_lut(), out)
    ref = normalize_frames_arithmetic(cv2.resize(f, (1024, 1024))[None], IMG_MEAN, IMG_STD)   # cv2 itself
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))


@pytest.mark.gpu
def test_ingest_kernel_strided_source_and_errors():
    from detsam2_b200 import ops
    from detsam2_b200.capi import Ds2Error
    rng = np.random.default_rng(23)
    big = torch.from_numpy(rng.integers(0, 256, (3, 90, 160, 3), dtype=np.uint8)).cuda()
    crop = big[:, 10:70, 20:140]                      # row pitch and frame stride larger than the crop
    out = torch.empty((3, 3, 64, 64), dtype=torch.float16, device="cuda")
    ops.ingest_frames(crop, _dev_lut(), out)
    ref = _expected(list(crop.cpu().numpy()), 64)
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    with pytest.raises(Ds2Error):
        ops.ingest_frames(crop.float(), _dev_lut(), out)
    with pytest.raises(Ds2Error):
        ops.ingest_frames(crop, _dev_lut(), out[:2])
    with pytest.raises(Ds2Error):
        ops.ingest_frames(crop.cpu(), _dev_lut(), out)


@pytest.mark.gpu
@pytest.mark.parametrize("offload", [True, False])
def test_load_video_frames_device_ingest_equals_host_path(offload):
    rng = np.random.default_rng(24)
    frames = [_frame(rng, 135, 240) for _ in range(37)]           # > 2 staging rounds of 16
    frames += [_frame(rng, 64, 64) for _ in range(3)]             # a size change mid-list (own launch)
    dev = torch.device("cuda")
    host, hh, hw = load_video_frames(frames, 128, offload_video_to_cpu=True, compute_device=dev, device_ingest=False)
    got, gh, gw = load_video_frames(frames, 128, offload_video_to_cpu=offload, compute_device=dev, device_ingest=True)
    assert (gh, gw) == (hh, hw) == (135, 240)
    assert got.is_cuda == (not offload)
    assert torch.equal(got.cpu().view(torch.int16), host.view(torch.int16))
    with pytest.raises(RuntimeError):
        load_video_frames([frames[0].astype(np.float32)], 128, compute_device=dev, device_ingest=True)


@pytest.mark.gpu
def test_predictor_state_identical_with_device_ingest(monkeypatch):
    """init_state / update_state through the public API: the `images` tensor of the session is the same bits whichever
    side resized the frames."""
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.synthetic import BilliardVideo
    frames = list(BilliardVideo(num_objects=2, height=360, width=640, num_frames=6, seed=5).frames())
    pred = build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_t.yaml", device="cuda", image_size=512)
    st = pred.init_state(frames[:3])
    st = pred.update_state(frames[3:], st)
    monkeypatch.setenv("DS2_HOST_INGEST", "1")
    st_h = pred.init_state(frames[:3])
    st_h = pred.update_state(frames[3:], st_h)
    assert st["images"].device.type == "cpu" and tuple(st["images"].shape) == (6, 3, 512, 512)
    assert torch.equal(st["images"].view(torch.int16), st_h["images"].view(torch.int16))
