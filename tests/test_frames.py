"""Frame ingest: the gather-table path of detsam2_b200.frames.load_video_frames must be
bit-identical to the reference's arithmetic (misc.py:336-359: /255 in float64 -> fp16, `-= mean`, `/= std`)."""
import numpy as np
import torch

from detsam2_b200.frames import IMG_MEAN, IMG_STD, load_video_frames, normalize_frames_arithmetic


def test_gather_table_is_bit_identical_to_reference_arithmetic():
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (64, 64, 3), dtype=np.uint8) for _ in range(3)]
    frames[0][:16, :16] = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)   # every byte value in every channel
    got, h, w = load_video_frames(frames, 64)
    ref = normalize_frames_arithmetic(np.stack(frames), IMG_MEAN, IMG_STD)
    assert (h, w) == (64, 64) and got.dtype == torch.float16 and tuple(got.shape) == (3, 3, 64, 64)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


def test_resize_path_matches_cv2_then_arithmetic():
    import cv2
    rng = np.random.default_rng(4)
    frames = [rng.integers(0, 256, (45, 80, 3), dtype=np.uint8) for _ in range(2)]
    got, h, w = load_video_frames(frames, 32)
    ref = normalize_frames_arithmetic(np.stack([cv2.resize(f, (32, 32)) for f in frames]), IMG_MEAN, IMG_STD)
    assert (h, w) == (45, 80)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


def test_single_array_and_bad_input():
    import pytest
    f = np.zeros((8, 8, 3), dtype=np.uint8)
    got, h, w = load_video_frames(f, 8)
    assert tuple(got.shape) == (1, 3, 8, 8)
    with pytest.raises(NotImplementedError):
        load_video_frames(12345, 8)
