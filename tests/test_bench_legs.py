"""bench.py's baseline legs (host logic only, no GPU): the pre-filled steady-state session that the CPU and GPU-library
baselines time must really present a FULL memory bank to the first timed frame — 1 conditioning frame + 6 recent frames
+ 16 object pointers, at the real object count — so that those legs time real steps instead of extrapolating."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_steady_state_session_presents_a_full_bank():
    import bench
    from detsam2_b200.config import get_config
    from detsam2_b200.memory_bank import plan_memory
    from detsam2_b200.predictor import SAM2VideoPredictor
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False), fill_hole_area=0)
    args = argparse.Namespace(objects=3, model="tiny")
    with torch.inference_mode():
        st, first = bench._steady_state_session(pred, args, steps=2, warmup=1)
    assert first == 17 and st["num_frames"] == 1 + 16 + 1 + 2
    od = st["output_dict"]
    assert sorted(od["cond_frame_outputs"]) == [0] and sorted(od["non_cond_frame_outputs"]) == list(range(1, 17))
    plan = plan_memory(first, od, st["num_frames"], False, None, cfg.num_maskmem, cfg.max_cond_frames_in_attn,
                       cfg.max_obj_ptrs_in_encoder)
    assert len(plan.frames) == 7 and len(plan.ptrs) == 16          # N = 7 * T + 4 * 16 tokens per object
    for _, out in plan.frames:
        assert out["maskmem_features"].shape[0] == 3               # stored at the real object count
    # and the session tracks: one real step through every seam
    with torch.inference_mode():
        f, ids, m = next(pred.propagate_in_video(st, start_frame_idx=first))
    assert f == first and list(ids) == [0, 1, 2] and tuple(m.shape) == (3, 1, 256, 256)
