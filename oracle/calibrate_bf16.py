"""TEST INFRASTRUCTURE — measures the UNMODIFIED reference's own bf16-autocast-vs-fp32 deviation on the
golden scenarios (CPU autocast, build container only) and writes tests/golden/ref_bf16_deviation.json.
The GPU parity tests use these numbers as the calibration of "within bf16 tolerance": the CUDA engine
(bf16 operands, fp32 accumulation) must not deviate from the fp32 reference by more than the reference
itself does when run the way Det-SAM2 runs it (torch.autocast bf16, det_sam2_RT.py:101-103).

    python -m oracle.calibrate_bf16               # all scenarios
    python -m oracle.calibrate_bf16 points_api    # only the named ones (the others keep their entries)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from detsam2_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_shim, scenarios  # noqa: E402


def deviations(got, ref):
    """per array-kind (suffix after the last '.'): max over arrays of rel-rms; min IoU for masks."""
    out = {}
    for k, r in ref.items():
        if np.issubdtype(r.dtype, np.integer) or k not in got:
            continue
        kind = k.rsplit(".", 1)[-1]
        g64, r64 = got[k].astype(np.float64), r.astype(np.float64)
        err = float(np.sqrt(np.mean((g64 - r64) ** 2)) / max(np.sqrt(np.mean(r64 ** 2)), 1e-12))
        d = out.setdefault(kind, {"rel_rms_max": 0.0, "rel_rms_mean": [], "iou_min": 1.0, "iou_mean": []})
        d["rel_rms_max"] = max(d["rel_rms_max"], err)
        d["rel_rms_mean"].append(err)
        if kind in ("video_res_masks", "pred_masks"):
            a, b = got[k] > 0, r > 0
            u = np.logical_or(a, b).sum()
            iou = 1.0 if u == 0 else float(np.logical_and(a, b).sum() / u)
            d["iou_min"] = min(d["iou_min"], iou)
            d["iou_mean"].append(iou)
    for d in out.values():
        d["rel_rms_mean"] = float(np.mean(d["rel_rms_mean"]))
        d["iou_mean"] = float(np.mean(d["iou_mean"])) if d["iou_mean"] else None
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    path = os.path.join(ROOT, "tests", "golden", "ref_bf16_deviation.json")
    names = sys.argv[1:] or ["stream", "preload", "offline", "mask_prompt", "points_api", "refine_click"]
    res = {}
    if sys.argv[1:] and os.path.exists(path):     # re-calibrating some scenarios keeps the others
        with open(path) as f:
            res = json.load(f)["scenarios"]
    for name in names:
        cfg = scenarios.scenario_config(name)
        sd = synthetic_state_dict(cfg, 0)
        ref = ref_shim.build_reference_predictor(cfg, sd, device="cpu")
        gold, _ = scenarios.load_golden(name)
        # autocast only around the compute seams: CPU autocast cannot "prioritize" the fp16 frame tensors
        # that update_state concatenates (svp:196)
        for meth in ("forward_image", "track_step", "_encode_new_memory"):
            def wrap(fn):
                def inner(*a, **k):
                    with torch.autocast("cpu", dtype=torch.bfloat16):
                        return fn(*a, **k)
                return inner
            setattr(ref, meth, wrap(getattr(ref, meth)))
        got = scenarios.SCENARIOS[name](ref)
        res[name] = deviations(got, gold)
        print(name, json.dumps(res[name], indent=1))
    with open(path, "w") as f:
        json.dump({"what": "reference under torch.autocast(cpu, bf16) vs reference fp32, per scenario and array kind",
                   "torch": torch.__version__, "scenarios": res}, f, indent=1)


if __name__ == "__main__":
    main()
