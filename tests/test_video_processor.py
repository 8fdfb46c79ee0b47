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


@pytest.mark.parametrize("name", sorted(scenarios.VP_SCENARIOS))
def test_video_processor_matches_reference_golden(name):
    """`video_processor`: the plain stream.  `video_processor_dup`: the detector reports one class twice per detection
    frame — the reference prompts that obj_id twice and the second call gets the first call's mask logits as a dense
    prompt (det_sam2_RT.py:288-302 -> svp:470-482); the batched box path must replay the repeats one by one."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold, _ = scenarios.load_golden(name)
    cfg = scenarios.scenario_config(name)
    eng = O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False)
    got = scenarios.run_video_processor(_make_vp_factory(eng), **scenarios.VP_SCENARIOS[name])
    assert set(got) == set(gold)
    bad = scenarios.compare(got, gold, 1e-4, iou_min=0.999)
    assert not bad, "\n".join(bad)
    # the third object is only reported from frame 4 on, yet chunk-1's reverse pass re-tracks frames >= 2 with it
    assert gold["f0.obj_ids"].tolist() == [0, 1] and gold["f10.obj_ids"].tolist() == [0, 1, 2]


def test_repeated_class_is_not_a_replacement():
    """The fixture with a repeated class must differ from the plain one on the repeated object (otherwise the scenario
    would not distinguish "second box replaces the first" from the reference's behaviour)."""
    a, _ = scenarios.load_golden("video_processor")
    b, _ = scenarios.load_golden("video_processor_dup")
    assert any(scenarios.packed_mask_iou(a[k], b[k]) < 0.999 for k in a if k.endswith("masks_packed"))


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).available(), reason="/root/reference not present")
def test_pickled_bank_is_loadable_by_the_reference(tmp_path):
    """det_sam2_RT.py:489-503: a bank saved by this repo's VideoProcessor with a non-.ds2bank suffix is a plain pickle of
    the reference's state schema.  It must unpickle WITHOUT this package importable (no engine handles inside) and the
    UNMODIFIED reference predictor must be able to continue a stream from it."""
    import pickle
    import subprocess
    import sys
    from detsam2_b200.synthetic import BilliardVideo, GroundTruthDetector
    from oracle import ref_shim
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = scenarios.scenario_config("preload")
    sd = synthetic_state_dict(cfg, 0)
    eng = O.OracleEngine(cfg, sd, fill_holes=False)
    vid = BilliardVideo(num_objects=2, height=160, width=224, num_frames=6, seed=4)
    bank = str(tmp_path / "bank.pkl")
    with torch.inference_mode():
        vp = VideoProcessor(predictor=SAM2VideoPredictor(eng, fill_hole_area=0, feature_cache_frames=8),
                            detector=GroundTruthDetector(vid, detect_interval=1), frame_buffer_size=3, detect_interval=1,
                            max_frame_num_to_track=3, max_inference_state_frames=-1, save_inference_state_path=bank)
        vp.run(frames=(vid.frame(t) for t in range(3)))
    # (1) no class of this package inside the pickle: a process that cannot import detsam2_b200 loads it
    code = ("import pickle, sys; sys.modules['detsam2_b200'] = None; st = pickle.load(open(sys.argv[1], 'rb')); "
            "assert st['cached_features'] == {}; print(sorted(st['output_dict']['cond_frame_outputs']))")
    out = subprocess.run([sys.executable, "-c", code, bank], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "[0, 1, 2]"
    # (2) the unmodified reference continues the stream from it (svp:123-156, det_sam2_RT.py:540-560)
    ref = ref_shim.build_reference_predictor(cfg, sd, device="cpu")
    with open(bank, "rb") as f:
        st = pickle.load(f)
    with torch.inference_mode():
        st["preloading_memory_cond_frame_idx"] = list(st["output_dict"]["cond_frame_outputs"].keys())
        st["preloading_memory_non_cond_frames_idx"] = list(st["output_dict"]["non_cond_frame_outputs"].keys())
        ref.init_preloading_state(st, offload_video_to_cpu=True, offload_state_to_cpu=False)
        st = ref.update_state([vid.frame(t) for t in range(3, 6)], st)
        got = {f: m for f, _, m in ref.propagate_in_video(st, start_frame_idx=2, max_frame_num_to_track=3)}
    assert sorted(got) == [2, 3, 4, 5]
    # the balls are still tracked by the reference on the last frame
    for oid, (x0, y0, x1, y1) in vid.boxes(5).items():
        m = got[5][oid, 0] > 0
        assert bool(m.any()) and bool(m[int(y0):int(y1) + 1, int(x0):int(x1) + 1].any())


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
