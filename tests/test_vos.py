"""the per-object mask-prompt VOS driver (detsam2_b200/vos.py) against the UNMODIFIED reference
function ``vos_inference`` of tools/vos_inference.py, both on CPU fp32 (reference predictor vs this repo's predictor over
the oracle engine), on a small synthetic DAVIS-style dataset: JPEG frames + a palette PNG of two objects on frame 0."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from detsam2_b200 import vos
from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.synthetic import BilliardVideo
from detsam2_b200.weights import synthetic_state_dict
from oracle import ref_shim, scenarios
from oracle import sam2_oracle as O


def _dataset(root, per_obj=False):
    from PIL import Image
    vid = BilliardVideo(num_objects=2, height=160, width=224, num_frames=3, seed=13)
    vdir = os.path.join(root, "JPEGImages", "clip")
    os.makedirs(vdir)
    for t in range(3):
        Image.fromarray(vid.frame(t)).save(os.path.join(vdir, f"{t:05d}.jpg"), quality=95)
    yy, xx = np.mgrid[0:160, 0:224]
    ids = np.zeros((160, 224), np.uint8)
    for oid, (cx, cy) in enumerate(vid.centers(0)):
        ids[(xx - cx) ** 2 + (yy - cy) ** 2 <= (vid.radius - 1) ** 2] = oid + 1
    adir = os.path.join(root, "Annotations", "clip")
    os.makedirs(adir)
    vos.save_ann_png(os.path.join(adir, "00000.png"), ids, vos.DAVIS_PALETTE)
    return os.path.join(root, "JPEGImages"), os.path.join(root, "Annotations")


def test_helpers_round_trip(tmp_path):
    ids = np.zeros((6, 8), np.uint8)
    ids[1:3, 1:4] = 1
    ids[2:5, 3:7] = 2
    per = vos.split_objects(ids)
    assert sorted(per) == [1, 2]
    per[1] = per[1] | per[2]                      # overlap: the smaller id wins
    merged = vos.merge_objects(per, 6, 8)
    assert (merged[2:5, 3:7] == 1).all() and (merged[1:3, 1:4] == 1).all()
    p = str(tmp_path / "m.png")
    vos.save_ann_png(p, ids, vos.DAVIS_PALETTE)
    back, pal = vos.load_ann_png(p)
    assert np.array_equal(back, ids) and bytes(pal[:768]) == vos.DAVIS_PALETTE


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_vos_driver_matches_reference_function(tmp_path):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = scenarios.scenario_config("stream")
    sd = synthetic_state_dict(cfg, 0)
    frames_dir, ann_dir = _dataset(str(tmp_path))
    # ---- the reference: tools/vos_inference.py imported unmodified (its hydra-based builder import is stubbed) ----
    ref_pred = ref_shim.build_reference_predictor(cfg, sd, device="cpu")
    sys.modules["sam2.build_sam"] = types.ModuleType("sam2.build_sam")
    sys.modules["sam2.build_sam"].build_sam2_video_predictor = lambda *a, **k: ref_pred
    tools = os.path.join(ref_shim.REF_ROOT, "tools")
    sys.path.insert(0, tools)
    try:
        sys.modules.pop("vos_inference", None)
        import vos_inference as ref_vos
    finally:
        sys.path.remove(tools)
    ref_out = str(tmp_path / "ref_out")
    ref_vos.vos_inference(predictor=ref_pred, base_video_dir=frames_dir, input_mask_dir=ann_dir, output_mask_dir=ref_out,
                          video_name="clip")
    # ---- this repo ----
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0)
    our_out = str(tmp_path / "our_out")
    segs = vos.vos_inference(pred, frames_dir, ann_dir, our_out, "clip")
    assert sorted(segs) == [0, 1, 2] and sorted(segs[2]) == [1, 2]
    for t in range(3):
        a, pa = vos.load_ann_png(os.path.join(ref_out, "clip", f"{t:05d}.png"))
        b, pb = vos.load_ann_png(os.path.join(our_out, "clip", f"{t:05d}.png"))
        assert pa == pb
        assert (a != b).mean() < 1e-3, (t, float((a != b).mean()))       # fp32 vs fp32: at most a stray threshold pixel
    # error behaviour: a mask directory without the first frame's annotation
    with pytest.raises(RuntimeError):
        vos.vos_inference(pred, frames_dir, str(tmp_path / "nowhere"), our_out, "clip")
