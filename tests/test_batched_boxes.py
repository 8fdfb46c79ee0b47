"""``SAM2VideoPredictor.add_new_boxes`` (all boxes of a frame in one B-wide decoder call) must leave
the inference state exactly as the reference's sequence of per-object ``add_new_points_or_box`` calls does
(det_sam2_RT.py:285-316), and tracking from that state must give the same masks.  Host logic on the CPU oracle
engine here; the CUDA engine is covered by tests/test_engine_gpu.py::test_batched_boxes_on_cuda_engine."""
import numpy as np
import pytest
import torch

from detsam2_b200.config import get_config
from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.synthetic import BilliardVideo
from detsam2_b200.weights import synthetic_state_dict


def _run(pred, vid, batched, nframes=3):
    with torch.inference_mode():
        st = pred.init_state([vid.frame(t) for t in range(nframes)])
        boxes = {oid: np.asarray(b, dtype=np.float32) for oid, b in vid.boxes(0).items()}
        if batched:
            f, ids, m = pred.add_new_boxes(st, 0, boxes)
        else:
            for oid, b in boxes.items():
                f, ids, m = pred.add_new_points_or_box(st, 0, oid, box=b)
        prompt = (f, list(ids), m.clone())
        temp = {oi: {k: (None if v is None else v.clone()) for k, v in d["cond_frame_outputs"][0].items()}
                for oi, d in st["temp_output_dict_per_obj"].items()}
        tracked = [(fi, list(oids), mm.clone()) for fi, oids, mm in pred.propagate_in_video(st)]
    return prompt, temp, tracked, st


def check_equivalent(make_predictor, vid, atol):
    a = _run(make_predictor(), vid, batched=False)
    b = _run(make_predictor(), vid, batched=True)
    assert a[0][0] == b[0][0] and a[0][1] == b[0][1]
    assert torch.allclose(a[0][2].float().cpu(), b[0][2].float().cpu(), atol=atol)
    assert a[1].keys() == b[1].keys()
    for oi in a[1]:
        for k, v in a[1][oi].items():
            w = b[1][oi][k]
            assert (v is None) == (w is None), (oi, k)
            if v is not None:
                assert v.shape == w.shape and torch.allclose(v.float().cpu(), w.float().cpu(), atol=atol), (oi, k)
    assert len(a[2]) == len(b[2])
    for (fa, ia, ma), (fb, ib, mb) in zip(a[2], b[2]):
        assert fa == fb and ia == ib
        assert torch.allclose(ma.float().cpu(), mb.float().cpu(), atol=atol)
    for key in ("obj_id_to_idx", "obj_idx_to_id", "obj_ids"):
        assert a[3][key] == b[3][key]
    for oi in a[3]["point_inputs_per_obj"]:
        pa, pb = a[3]["point_inputs_per_obj"][oi][0], b[3]["point_inputs_per_obj"][oi][0]
        assert torch.equal(pa["point_coords"].cpu(), pb["point_coords"].cpu())
        assert torch.equal(pa["point_labels"].cpu(), pb["point_labels"].cpu())


def test_batched_boxes_equal_sequential_prompts_on_oracle_engine():
    from oracle import sam2_oracle as O
    cfg = get_config("tiny")
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=3, height=160, width=224, num_frames=3, seed=13)
    check_equivalent(lambda: SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0), vid, atol=2e-4)


def test_batched_boxes_fall_back_and_validate():
    from oracle import sam2_oracle as O
    cfg = get_config("tiny")
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=2, height=160, width=224, num_frames=2, seed=14)
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0)
    with torch.inference_mode():
        st = pred.init_state([vid.frame(t) for t in range(2)])
        with pytest.raises(ValueError):
            pred.add_new_boxes(st, 0, {})
        boxes = {oid: np.asarray(b, dtype=np.float32) for oid, b in vid.boxes(0).items()}
        pred.add_new_boxes(st, 0, boxes)
        # a second prompt on the same frame finds earlier results there: sequential path, same API result shape
        f, ids, m = pred.add_new_boxes(st, 0, boxes)
        assert f == 0 and list(ids) == [0, 1] and tuple(m.shape) == (2, 1, 160, 224)
