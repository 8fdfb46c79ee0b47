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


def _last_json(path):
    import json
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_bench_lines_keep_the_contract():
    """The bench line of the final commit and its reference arm (profiles/) carry every key the driver reads."""
    line = _last_json(os.path.join(ROOT, "profiles", "r2_s24_bench.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    assert abs(line["value"] - 1e3 * line["n_gpus"] / line["ms_per_step"]) < 0.05 * line["value"]
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["h2d_bytes_per_step"] == 3 * 1024 * 1024 * 2 and e2e["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    roof = line["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-3 and roof["traffic"] > 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
    ref = _last_json(os.path.join(ROOT, "profiles", "r2_s20_bench_reference.json"))
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    assert ref["cpu_baseline"]["value"] == ref["value"]


def test_clock_sampler_uses_only_samples_inside_the_timed_region():
    """The sampler runs from before the warm-up (nvidia-smi's start-up must not fall into the timed region); rows that
    arrived outside [begin, end + one period] are ignored."""
    import bench
    s = bench.ClockSampler(0)

    class _Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s.proc = _Done()
    s.t0, s.t1 = 100.0, 101.0
    row = "{}, 1965, {}, Not Active, Not Active, Not Active, {}"
    s.rows = [(99.0, row.format(300, 150.0, "Not Active")),          # warm-up: ignored
              (100.2, row.format(1900, 900.0, "Active")), (100.6, row.format(1800, 980.0, "Active")),
              (101.05, row.format(1850, 990.0, "Active")),           # covers the last 100 ms of the region
              (102.0, row.format(500, 200.0, "Not Active"))]         # after the region: ignored
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1850.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w"] == 980.0
