"""Det-SAM2's stream driver: this repo's VideoProcessor (over the CPU oracle engine) against the
UNMODIFIED reference VideoProcessor.run (golden fixture tests/golden/video_processor.npz, made by
oracle/gen_golden.py with the reference's third-party imports stubbed): same frames get segmented,
same object ids per frame (incl. an id that appears mid-stream), same masks, same window contents."""
import os

import numpy as np
import pytest
import torch

from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.video_processor import VideoProcessor
from detsam2_b200.weights import synthetic_state_dict
from oracle import sam2_oracle as O
from oracle import scenarios


def _make_vp_factory(engine):
    def make_vp(detector, **kw):
        return VideoProcessor(predictor=SAM2VideoPredictor(engine, fill_hole_area=0), detector=detector, **kw)
    return make_vp


def test_video_processor_matches_reference_golden():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold, _ = scenarios.load_golden("video_processor")
    cfg = scenarios.scenario_config("video_processor")
    eng = O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False)
    got = scenarios.run_video_processor(_make_vp_factory(eng))
    assert set(got) == set(gold)
    bad = scenarios.compare(got, gold, 1e-4, iou_min=0.999)
    assert not bad, "\n".join(bad)
    # the third object is only reported from frame 4 on, yet chunk-1's reverse pass re-tracks frames >= 2 with it
    assert gold["f0.obj_ids"].tolist() == [0, 1] and gold["f10.obj_ids"].tolist() == [0, 1, 2]


def test_constructor_contract():
    class Eng:  # never used for compute in this test
        cfg = scenarios.scenario_config("stream")
        device = torch.device("cpu")
    pred = SAM2VideoPredictor(Eng())
    with pytest.raises(AssertionError):   # det_sam2_RT.py:67-68
        VideoProcessor(predictor=pred, save_inference_state_path="/tmp/x.pkl", max_inference_state_frames=60)
    with pytest.raises(NotImplementedError):
        VideoProcessor(predictor=pred, vis_frame_stride=5)
    vp = VideoProcessor(predictor=pred, detect_interval=30)
    with pytest.raises(RuntimeError):
        vp.detect_predict([np.zeros((8, 8, 3), np.uint8)], 0)    # detection requested, no detector
    vp = VideoProcessor(predictor=pred, detect_interval=-1)
    assert vp.detect_predict([np.zeros((8, 8, 3), np.uint8)], 0) == {}
    with pytest.raises(ValueError):
        vp.run()


def test_detect_predict_grid_and_special_classes():
    class Eng:
        cfg = scenarios.scenario_config("stream")
        device = torch.device("cpu")
    seen = []

    def det(frames):
        seen.append(len(frames))
        return [[{"coordinates": np.array([1, 2, 3, 4], np.float32), "class": np.array([11.0]), "confidence": np.array([0.9])},
                 {"coordinates": np.array([5, 6, 7, 8], np.float32), "class": np.array([11.0]), "confidence": np.array([0.9])},
                 {"coordinates": np.array([0, 0, 9, 9], np.float32), "class": np.array([3.0]), "confidence": np.array([0.9])}]
                for _ in frames]
    vp = VideoProcessor(predictor=SAM2VideoPredictor(Eng()), detector=det, detect_interval=5)
    res = vp.detect_predict([np.zeros((8, 8, 3), np.uint8)] * 7, past_num_frames=8)   # absolute 8..14 -> frame 10
    assert list(res) == ["frame_10"] and seen == [1]
    assert len(vp.special_classes_detection) == 2 and vp.special_classes_count == 2
