"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, through oracle/ref_shim.py) on CPU fp32 over the scenarios of oracle/scenarios.py.

Run in the build container only (the reference does not exist on the GPU box):

    python -m oracle.gen_golden            # all scenarios
    python -m oracle.gen_golden stream     # one scenario

Inputs are not stored: they are regenerated from seeds (detsam2_b200.weights.synthetic_state_dict
seed 0, detsam2_b200.synthetic.BilliardVideo with the scenario's seed); a fingerprint of the weights
and of frame 0 is stored so that a drifted generator is detected instead of silently mis-compared.
Float arrays are stored as fp32 (fp16 for the large video-res logits, whose only consumer is the
``> 0`` threshold and a loose rel-rms check), compressed.

The reference's CPU path silently skips hole filling (sam2._C is a CUDA extension,
misc.py:389-391), so these fixtures are "before hole fill"; hole filling is pinned separately
(tests/test_cc_oracle.py against scipy.ndimage.label).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from detsam2_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_shim, scenarios  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# VideoProcessor.run scenarios: fixture name -> keyword arguments of scenarios.run_video_processor
VP_SCENARIOS = scenarios.VP_SCENARIOS


def fingerprint(sd):
    """Cheap, order-stable fingerprint of a state dict (sum and abs-sum of every 37th tensor)."""
    keys = sorted(sd)[::37]
    return np.asarray([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in keys], dtype=np.float64)


def generate_image_predictor():
    """tests/golden/image_predictor.npz from the UNMODIFIED reference SAM2ImagePredictor (sam2_image_predictor.py)."""
    ref_shim.install()
    from sam2.sam2_image_predictor import SAM2ImagePredictor
    cfg = scenarios.image_predictor_config()
    sd = synthetic_state_dict(cfg, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref_shim.build_reference_predictor(cfg, sd, device="cpu")   # a SAM2Base subclass: what the predictor wraps
    rec = scenarios.run_image_predictor(SAM2ImagePredictor(model))
    rec["__weights_fingerprint"] = fingerprint(sd)
    rec["__torch_version"] = np.asarray(torch.__version__)
    path = os.path.join(GOLDEN_DIR, "image_predictor.npz")
    np.savez_compressed(path, **rec)
    print(f"image_predictor: {len(rec)} arrays, {os.path.getsize(path) / 1e6:.2f} MB -> {path}")


def generate(name):
    if name == "image_predictor":
        return generate_image_predictor()
    cfg = scenarios.scenario_config(name)
    sd = synthetic_state_dict(cfg, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ref_shim.build_reference_predictor(cfg, sd, device="cpu")
    t0 = time.time()
    if name in VP_SCENARIOS:
        def make_vp(detector, **kw):
            cls = ref_shim.reference_video_processor_class(ref, detector)
            return cls(sam2_checkpoint=None, model_cfg=None, detect_model_weights=None, **kw)
        rec = scenarios.run_video_processor(make_vp, **VP_SCENARIOS[name])
    else:
        rec = scenarios.SCENARIOS[name](ref)
    dt = time.time() - t0
    out = {}
    for k, v in rec.items():
        if k.endswith("video_res_masks") or k.endswith("maskmem_features"):
            v = v.astype(np.float16)  # maskmem_features are bf16 in the state: fp16 holds them exactly (|x| >= 6e-5)
        out[k] = v
    out["__weights_fingerprint"] = fingerprint(sd)
    out["__torch_version"] = np.asarray(torch.__version__)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(rec)} arrays, reference ran {dt:.1f} s, {os.path.getsize(path) / 1e6:.2f} MB -> {path}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(scenarios.SCENARIOS) + list(VP_SCENARIOS) + ["image_predictor"]
    for n in names:
        generate(n)
