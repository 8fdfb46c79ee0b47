"""Live pin of the oracle against the UNMODIFIED reference imported from /root/reference (build
container only; skipped on the GPU box where the reference does not exist — the committed fixtures
under tests/golden/ carry the same comparison there)."""
import os

import pytest
import torch

from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.weights import synthetic_state_dict
from oracle import ref_shim, scenarios
from oracle import sam2_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


def test_preload_scenario_live():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = scenarios.scenario_config("preload")
    sd = synthetic_state_dict(cfg, 0)
    ref = scenarios.run_preload(ref_shim.build_reference_predictor(cfg, sd, device="cpu"))
    got = scenarios.run_preload(SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0))
    assert set(got) == set(ref)
    bad = scenarios.compare(got, ref, 1e-4, iou_min=0.999)
    assert not bad, "\n".join(bad)


def test_error_behaviour_matches_reference():
    """ValueError / RuntimeError on the same bad calls (svp:363-366, 385-388, 943-944)."""
    import numpy as np
    from detsam2_b200.synthetic import BilliardVideo
    cfg = scenarios.scenario_config("stream")
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=1, height=128, width=128, num_frames=2, seed=0)
    frames = list(vid.frames())
    preds = [ref_shim.build_reference_predictor(cfg, sd, device="cpu"),
             SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False))]
    for p in preds:
        st = p.init_state(frames)
        with pytest.raises(ValueError):
            p.add_new_points_or_box(st, 0, 0, points=np.zeros((1, 2), np.float32))          # labels missing
        with pytest.raises(ValueError):
            p.add_new_points_or_box(st, 0, 0)                                                # nothing given
        with pytest.raises(ValueError):
            p.add_new_points_or_box(st, 0, 0, box=np.array([1, 2, 30, 40], np.float32), clear_old_points=False)
        with pytest.raises(RuntimeError):
            next(iter(p.propagate_in_video(st)))                                             # no prompts yet
        with pytest.raises(AssertionError):
            p.update_state([np.zeros((64, 64, 3), np.uint8)], st)                            # size mismatch


def test_non_overlapping_constraints_match_reference():
    """sam2_base.py:934-952 (off by default; `non_overlap_masks=True` sessions): same tensor, value for value."""
    ref_shim.install()
    from sam2.modeling.sam2_base import SAM2Base
    torch.manual_seed(0)
    m = torch.randn(5, 1, 17, 23) * 8
    m[:, :, :4] = m[0:1, :, :4]            # ties: argmax must pick the same (first) object on both sides
    ref = SAM2Base._apply_non_overlapping_constraints(None, m.clone())
    got = SAM2VideoPredictor._apply_non_overlapping_constraints(None, m.clone())
    assert torch.equal(ref, got)
    one = torch.randn(1, 1, 4, 4)
    assert torch.equal(SAM2VideoPredictor._apply_non_overlapping_constraints(None, one), one)


def test_select_closest_cond_frames_matches_reference_on_random_inputs():
    """sam2_utils.py:19-66 incl. Det-SAM2's forced preload frames: same keys, same ORDER (the order is the memory row
    order), for 400 random configurations."""
    import random
    ref_shim.install()
    from sam2.modeling.sam2_utils import select_closest_cond_frames as ref_select
    from detsam2_b200.memory_bank import select_closest_cond_frames as our_select
    rnd = random.Random(7)
    for _ in range(400):
        n = rnd.randint(1, 40)
        keys = rnd.sample(range(0, 120), n)
        if rnd.random() < 0.5:
            keys.sort()
        cond = {t: ("out", t) for t in keys}
        frame_idx = rnd.randint(0, 130)
        cap = rnd.choice([-1, 2, 3, 5, 20])
        preload = None if rnd.random() < 0.5 else rnd.sample(keys, rnd.randint(0, min(6, n)))
        rs, ru = ref_select(frame_idx, dict(cond), cap, preload)
        gs, gu = our_select(frame_idx, dict(cond), cap, preload)
        assert list(rs.items()) == list(gs.items()) and list(ru.items()) == list(gu.items()), (keys, frame_idx, cap, preload)
