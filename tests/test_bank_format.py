"""the compact preload-bank format (detsam2_b200/bank_format.py) is lossless w.r.t. the reference's
pickle (det_sam2_RT.py:489-503), restores the aliasing between batched and per-object outputs, is several times smaller,
and a stream continued from it gives the same result."""
import os
import pickle

import numpy as np
import pytest
import torch

from detsam2_b200.bank_format import BankFormatError, load_bank, save_bank
from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.synthetic import BilliardVideo
from detsam2_b200.weights import synthetic_state_dict
from oracle import sam2_oracle as O
from oracle import scenarios


def _deep_equal(a, b, path="state"):
    assert type(a) is type(b) or (isinstance(a, dict) and isinstance(b, dict)), (path, type(a), type(b))
    if isinstance(a, torch.Tensor):
        assert a.dtype == b.dtype and a.shape == b.shape, path
        assert torch.equal(a.cpu(), b.cpu()), path
    elif isinstance(a, dict):
        assert list(a.keys()) == list(b.keys()), (path, list(a.keys()), list(b.keys()))
        for k in a:
            _deep_equal(a[k], b[k], f"{path}[{k!r}]")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _deep_equal(x, y, f"{path}[{i}]")
    elif isinstance(a, torch.device):
        assert a.type == b.type, path
    else:
        assert a == b, (path, a, b)


@pytest.fixture(scope="module")
def bank_state():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg = scenarios.scenario_config("preload")
    sd = synthetic_state_dict(cfg, 0)
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0)
    vid = BilliardVideo(num_objects=2, height=192, width=256, num_frames=5, seed=7)
    with torch.inference_mode():
        st = pred.init_state([vid.frame(t) for t in range(3)])
        for t in range(3):
            for oid, box in vid.boxes(t).items():
                pred.add_new_points_or_box(st, t, oid, box=np.asarray(box, dtype=np.float32))
        for _ in pred.propagate_in_video(st, start_frame_idx=2, max_frame_num_to_track=3, reverse=True):
            pass
    return pred, vid, st


def test_round_trip_equals_pickle_round_trip(bank_state, tmp_path):
    pred, vid, st = bank_state
    path = str(tmp_path / "bank.ds2bank")
    save_bank(st, path)
    got = load_bank(path, map_location="cpu")
    ref = dict(pickle.loads(pickle.dumps(st)))
    ref["cached_features"] = {}
    _deep_equal(ref, got)
    # aliasing restored: a per-object entry is a view of the batched tensor of the same frame
    f = next(iter(got["output_dict"]["cond_frame_outputs"]))
    whole = got["output_dict"]["cond_frame_outputs"][f]["pred_masks"]
    part = got["output_dict_per_obj"][1]["cond_frame_outputs"][f]["pred_masks"]
    assert part.untyped_storage().data_ptr() == whole.untyped_storage().data_ptr()
    assert torch.equal(part, whole[1:2])
    # and it is much smaller than the pickle
    pk = len(pickle.dumps(st))
    assert os.path.getsize(path) * 2 < pk, (os.path.getsize(path), pk)


def test_stream_continued_from_compact_bank_equals_pickled_bank(bank_state, tmp_path):
    pred, vid, st = bank_state
    path = str(tmp_path / "bank.ds2bank")
    save_bank(st, path)
    outs = []
    for loaded in (pickle.loads(pickle.dumps(st)), load_bank(path, map_location="cpu")):
        with torch.inference_mode():
            loaded["preloading_memory_cond_frame_idx"] = list(loaded["output_dict"]["cond_frame_outputs"].keys())
            loaded["preloading_memory_non_cond_frames_idx"] = list(loaded["output_dict"]["non_cond_frame_outputs"].keys())
            pred.init_preloading_state(loaded, offload_video_to_cpu=True, offload_state_to_cpu=False)
            loaded = pred.update_state([vid.frame(t) for t in range(3, 5)], loaded)
            outs.append([m.clone() for _, _, m in pred.propagate_in_video(loaded, start_frame_idx=2, max_frame_num_to_track=2)])
    assert len(outs[0]) == len(outs[1]) > 0
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.ds2bank"
    import zipfile
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("meta.json", '{"magic": "ds2bank/999", "state": null}')
    with pytest.raises(BankFormatError):
        load_bank(str(p))
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("other.txt", "hi")
    with pytest.raises(BankFormatError):
        load_bank(str(p))
    with pytest.raises(BankFormatError):
        save_bank({"bad": object()}, str(tmp_path / "y.ds2bank"))
