"""``SAM2ImagePredictor`` (sam2_image_predictor.py) over an engine.  The fixture
tests/golden/image_predictor.npz holds outputs of the UNMODIFIED reference class (oracle/gen_golden.py image_predictor):
a click with three candidate masks, a box (single mask, stability fallback), a refinement with the previous logits as
a dense prompt, and two boxes in one batched call — on a 480x640 image (antialiased resize to the model resolution).
CPU: the fp32 oracle engine against the fixture.  GPU: the CUDA engine against the fixture (bf16 band)."""
import os

import numpy as np
import pytest
import torch

from detsam2_b200.image_predictor import SAM2ImagePredictor
from detsam2_b200.weights import synthetic_state_dict
from oracle import scenarios


def _compare(got, gold, bound):
    assert set(got) == set(k for k in gold if not k.startswith("__"))
    lines, bad = [], []
    for k in sorted(got):
        r, g = gold[k], got[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        if np.issubdtype(r.dtype, np.integer) and not k.endswith("masks_packed"):
            assert np.array_equal(g, r), k
            continue
        if k.endswith("masks_packed"):
            lines.append(f"{k}: raw mask IoU {scenarios.packed_mask_iou(g, r):.4f}")
            continue
        g64, r64 = g.astype(np.float64), r.astype(np.float64)
        rel = np.sqrt(np.mean((g64 - r64) ** 2)) / max(np.sqrt(np.mean(r64 ** 2)), 1e-12)
        lines.append(f"{k}: rel-rms {rel:.2e}")
        if rel > bound(k):
            bad.append(lines[-1] + f" > {bound(k):.2e}")
    return lines, bad


def test_oracle_image_predictor_matches_reference_golden():
    from oracle import sam2_oracle as O
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold, _ = scenarios.load_golden("image_predictor")
    cfg = scenarios.image_predictor_config()
    ip = SAM2ImagePredictor(O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False))
    got = scenarios.run_image_predictor(ip)
    # fp32 against fp32; the refinement step takes the click step's logits as a dense prompt and amplifies their 1e-4
    # difference about five-fold
    lines, bad = _compare(got, gold, lambda k: 2e-3 if k.startswith("refine") else 2e-4)
    assert not bad, "\n".join(bad)
    # error behaviour (sam2_image_predictor.py:279-282)
    ip.reset_predictor()
    with pytest.raises(RuntimeError):
        ip.predict(point_coords=np.zeros((1, 2), np.float32), point_labels=np.ones(1, np.int32))
    with pytest.raises(NotImplementedError):
        ip.set_image("not an image")


@pytest.mark.gpu
def test_cuda_image_predictor_matches_reference_golden():
    from detsam2_b200.engine import CudaEngine
    gold, _ = scenarios.load_golden("image_predictor")
    cfg = scenarios.image_predictor_config()
    ip = SAM2ImagePredictor(CudaEngine(cfg, synthetic_state_dict(cfg, 0), device="cuda:0"))
    got = scenarios.run_image_predictor(ip)
    torch.cuda.synchronize()
    # bf16 operands / fp32 accumulation against the fp32 reference: per array no worse than 1.25x what the UNMODIFIED
    # reference shows against itself when run under bf16 autocast (tests/golden/ref_bf16_deviation_image_predictor.json;
    # click prompts are chaotic with random weights: the reference's own bf16 run deviates by 0.31 there)
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_bf16_deviation_image_predictor.json")) as f:
        calib = json.load(f)["arrays"]
    # (an `ious` array is one to three numbers — an extreme-value statistic: bounded by the worst of the four calibrations)
    iou_cal = max(v for k, v in calib.items() if k.endswith(".ious"))
    lines, bad = _compare(got, gold, lambda k: 1.25 * (iou_cal if k.endswith(".ious") else calib[k]) + 1e-3)
    print("\n".join(lines))
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_image_predictor.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    assert not bad, "\n".join(bad)
