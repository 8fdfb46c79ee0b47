"""SURVEY.md 8f rank 1, second half: the detector's pre-processing on the device.

CPU: the LetterBox geometry of detsam2_b200.detector_preproc equals the oracle's restatement of ultralytics 8.2.82 and
the hand-checked cases of its documentation (1080p -> 384 x 640 at imgsz 640, auto padding to the stride).
GPU: ds2_letterbox_frames against the oracle (cv2.resize + cv2.copyMakeBorder + torch's `/= 255`): identical bits, fp32
and fp16, over upscaling / downscaling / exact-2x / equal-size frames; and VideoProcessor with the device pre-processing
hands the detector the same tensor and ends with the same segmentation as the host flow.
"""
import numpy as np
import pytest
import torch

from detsam2_b200.detector_preproc import DeviceLetterbox, letterbox_params
from oracle import letterbox_oracle as LO

SIZES = [(1080, 1920), (720, 1280), (1024, 1024), (480, 640), (1280, 1280), (640, 640), (333, 517), (2160, 3840), (64, 48)]


@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("imgsz,auto", [(640, True), (640, False), (1280, True), ((384, 672), True)])
def test_letterbox_geometry_equals_oracle(hw, imgsz, auto):
    assert letterbox_params(hw, imgsz, auto) == LO.letterbox_params(hw, imgsz, auto)


def test_letterbox_geometry_known_cases():
    # 1080p at imgsz 640, minimum-rectangle padding: 640 x 360 resized, padded to 640 x 384 (12 rows top and bottom)
    assert letterbox_params((1080, 1920), 640, True)[:6] == (360, 640, 12, 0, 384, 640)
    # square padding (auto=False): 140 rows top and bottom
    assert letterbox_params((1080, 1920), 640, False)[:6] == (360, 640, 140, 0, 640, 640)
    # a square frame needs no border
    assert letterbox_params((1024, 1024), 640, True)[:6] == (640, 640, 0, 0, 640, 640)
    b = DeviceLetterbox(640).unletterbox([[0, 12, 640, 372]], (1080, 1920))
    assert np.allclose(b, [[0, 0, 1920, 1080]])
    assert np.allclose(LO.unletterbox_boxes([[0, 12, 640, 372]], (1080, 1920)), b)


def _frames(n, hw, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(hw[0] // 8 + 2, hw[1] // 8 + 2, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))[:hw[0], :hw[1]]      # blocky + noise: edges and flats
    out = []
    for i in range(n):
        out.append(np.ascontiguousarray((img.astype(np.int16) + rng.integers(-20, 21, size=img.shape)).clip(0, 255).astype(np.uint8)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("half", [False, True])
def test_device_letterbox_is_bit_exact(hw, half):
    frames = _frames(2, hw, 5)
    want = LO.preprocess(frames, 640, half=half)
    dev = torch.from_numpy(np.stack(frames)).cuda()
    got = DeviceLetterbox(640, half=half)(dev)
    torch.cuda.synchronize()
    assert got.dtype == want.dtype and tuple(got.shape) == tuple(want.shape)
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_device_letterbox_other_model_sizes_and_strided_source():
    frames = _frames(3, (720, 1280), 9)
    big = torch.from_numpy(np.stack(frames)).cuda()
    for imgsz, auto in ((1280, True), (640, False), ((384, 672), True)):
        want = torch.stack([LO.preprocess([f], imgsz, auto=auto)[0] for f in frames])
        got = DeviceLetterbox(imgsz, auto=auto)(big)
        assert torch.equal(got.cpu(), want), (imgsz, auto)
    # a frame that is a window into a larger buffer (row pitch > 3 * width)
    canvas = torch.zeros(1, 800, 1400, 3, dtype=torch.uint8, device="cuda")
    canvas[0, 40:760, 60:1340] = big[1]
    got = DeviceLetterbox(640)(canvas[:, 40:760, 60:1340])
    assert torch.equal(got.cpu(), LO.preprocess([frames[1]], 640))


@pytest.mark.gpu
def test_video_processor_with_device_detector_preprocessing():
    """Same stream through VideoProcessor twice: host flow (BGR ndarrays to the detector) and device flow (one upload per
    chunk shared by detector pre-processing and SAM 2 ingest).  The detector sees the tensor ultralytics would have built
    on the host, its boxes come back in frame pixels, and the segmentation is identical."""
    from detsam2_b200.config import get_config
    from detsam2_b200.engine import CudaEngine
    from detsam2_b200.predictor import SAM2VideoPredictor
    from detsam2_b200.synthetic import BilliardVideo, GroundTruthDetector
    from detsam2_b200.video_processor import VideoProcessor
    from detsam2_b200.weights import synthetic_state_dict
    cfg = get_config("tiny", image_size=512)
    eng = CudaEngine(cfg, synthetic_state_dict(cfg, 0), device="cuda:0")
    H, W, n = 360, 640, 10
    vid = BilliardVideo(num_objects=2, height=H, width=W, num_frames=n, seed=17)
    lb = DeviceLetterbox(640)
    seen = []

    class TensorDetector:
        """Stands in for YOLO fed with a ready tensor: checks what it is given, answers in TENSOR pixels."""
        def __init__(self):
            self.gt = GroundTruthDetector(vid, detect_interval=4)

        def __call__(self, x):
            assert isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (3, 384, 640)
            seen.append(x.clone())
            new_h, new_w, top, left, _, _, r = letterbox_params((H, W), 640, True)
            out = []
            for dets in self.gt([None] * x.shape[0]):
                out.append([dict(d, coordinates=np.asarray(d["coordinates"], np.float32) * r + np.float32([left, top, left, top]))
                            for d in dets])
            return out

    segs = {}
    with torch.inference_mode():
        for mode in ("host", "device"):
            det = GroundTruthDetector(vid, detect_interval=4) if mode == "host" else TensorDetector()
            vp = VideoProcessor(predictor=SAM2VideoPredictor(eng, fill_hole_area=8), detector=det, frame_buffer_size=4,
                                detect_interval=4, max_frame_num_to_track=6, max_inference_state_frames=6, skip_classes=set(),
                                detector_preproc=lb if mode == "device" else None)
            segs[mode] = vp.run(frames=(vid.frame(t) for t in range(n)))
    assert len(seen) == 3                                     # frames 0, 4, 8
    want = LO.preprocess([vid.frame(0)], 640)
    assert torch.equal(seen[0].cpu(), want)
    assert sorted(segs["host"]) == sorted(segs["device"]) == list(range(n))
    for t in range(n):
        for oid in segs["host"][t]:
            assert np.array_equal(segs["host"][t][oid], segs["device"][t][oid]), (t, oid)
