"""Pins the CPU oracle (oracle/sam2_oracle.py driven through this repo's predictor) against outputs
of the UNMODIFIED reference, committed under tests/golden/ by oracle/gen_golden.py.

fp32 on both sides; the only differences are operation order (e.g. folded position terms), hence the
tight tolerance.  Integer bookkeeping (object ids, window contents after release_old_frames,
conditioning-frame sets) must be identical.
"""
import os

import numpy as np
import pytest
import torch

from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.weights import synthetic_state_dict
from oracle import sam2_oracle as O
from oracle import scenarios
from oracle.gen_golden import fingerprint

RTOL_RMS = 2e-4  # fp32 oracle vs fp32 reference, relative RMS per array


def oracle_predictor(name):
    cfg = scenarios.scenario_config(name)
    sd = synthetic_state_dict(cfg, 0)
    # the reference's CPU path skips hole filling (misc.py:389-391) -> compare before hole fill
    return SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0), sd


# large_1024 / bplus_720p: the full-size models (Hiera-L at 1024^2; base_plus with zero-padded 14/7 windows on 720p frames)
@pytest.mark.parametrize("name", ["stream", "preload", "offline", "mask_prompt", "points_api", "refine_click",
                                  "large_1024", "bplus_720p"])
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold, fp = scenarios.load_golden(name)
    pred, sd = oracle_predictor(name)
    np.testing.assert_allclose(fingerprint(sd), fp, rtol=1e-9,
                               err_msg="synthetic weights drifted from the ones the fixture was made with")
    got = scenarios.SCENARIOS[name](pred)
    assert set(got) == set(gold), (sorted(set(got) ^ set(gold)))
    bad = scenarios.compare(got, gold, RTOL_RMS, iou_min=0.999)
    assert not bad, "\n".join(bad)
