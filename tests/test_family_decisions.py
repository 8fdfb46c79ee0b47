"""CPU half of tests/test_configs_gpu.py::test_model_family_tracked_frames_match_oracle.

Round 1 ended with the GPU suite red because the oracle's decision log gained a second kind of entry after the
last GPU run and the GPU-only test still asserted the old length.  This test runs the ORACLE half of that GPU
test (same helper, same seeds, same assertions on the log) in the GPU-less container, so a change of the log's
shape, or a seed whose discrete choices are near ties, fails here first."""
import os

import pytest
import torch

from test_configs_gpu import FAMILY_CASES, check_family_decisions, run_family_case


@pytest.mark.parametrize("name,hw,seed", FAMILY_CASES)
def test_oracle_half_of_family_case(name, hw, seed):
    from detsam2_b200.config import get_config
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg = get_config(name)
    sd = synthetic_state_dict(cfg, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    O.DECISION_LOG = []
    try:
        masks, st = run_family_case(O.OracleEngine(cfg, sd, fill_holes=True), cfg, hw, seed)
    finally:
        decisions, O.DECISION_LOG = O.DECISION_LOG, None
    check_family_decisions(decisions, cfg)
    assert sorted(masks) == [0, 1, 2]
    for f in (1, 2):
        out = st["output_dict"]["non_cond_frame_outputs"][f]
        assert tuple(out["pred_masks"].shape) == (2, 1, 4 * cfg.feat_size, 4 * cfg.feat_size)
        assert tuple(masks[f].shape) == (2, 1, hw[0], hw[1])
