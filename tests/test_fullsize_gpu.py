"""Full-size parity at the headline configuration (BASELINE.json configs[1]): sam2.1_hiera_large, 1024^2, 16 box-prompted
objects, tracked until the memory bank is at its steady state (1 conditioning + 6 recent frames + 16 object pointers:
N = 7 * 4096 + 64 = 28 736 memory tokens), CUDA engine against the fp32 oracle on the same seeded inputs.

The oracle runs here on torch's own fp32 CUDA kernels (TF32 off): the same restated functions that are pinned on CPU
against the reference fixtures (tests/test_oracle_golden.py; CPU-vs-CUDA agreement of the oracle itself:
test_oracle_cuda_path_equals_cpu_path below) — on the host cores this case takes ~22 s per frame, on the device
seconds for the whole sequence.

Tracked frames choose one of three candidate masks per object by argmax over predicted IoUs (mask_decoder.py:150-158).
With plain seeded random weights the three predicted IoUs of an object converge as tracking goes on (measured: top-2
margin 0.012 on frame 1, 0.001 on frame 16, 0.0000 on frame 17 — profiles/r2_parity_large16_random_iou_head.txt), a bf16
implementation then legitimately picks another candidate, and from that frame on the object's memory differs and
nothing about it is comparable any more (1 of 16 objects was still comparable on frame 18).  A trained checkpoint has
decisive IoU predictions; the test restores that property with ONE change to the seeded weights, applied to both
sides: the last-layer bias of the IoU head gets +2 / 0 / -2 on the three multimask outputs, which makes every argmax
decisive (margin >= `MIN_MARGIN`, asserted) without touching the mask / pointer / memory arithmetic under test.  The
test still records both sides' choices and requires them to be identical for all 16 objects on every frame.

Hole filling is switched off on both sides here: it is a discrete post-process (components of area <= 8,
misc.py:365-393) with its own bit-exact tests, and a component of area 8 on one side and 9 on the other would flip
"decided" pixels for a reason that has nothing to do with the arithmetic compared here.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIN_MARGIN = 0.05   # smallest top-2 predicted-IoU margin the oracle may show with the decisive IoU-head bias


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def _iou(a, b):
    a, b = (a > 0).cpu(), (b > 0).cpu()
    u = (a | b).sum().item()
    return 1.0 if u == 0 else (a & b).sum().item() / u


# A pixel is "confident" when its fp32 logit exceeds BAND x the rms logit of its mask.  The band is 5x the rel-rms error
# bound the respective test asserts on the same logits (0.03 at the steady-state test, 0.06 at the reference fixtures),
# i.e. it follows from the STATED tolerance, not from the measured error.
BAND = 0.15
BAND_FIXTURE = 0.30


def _confident_iou(got, ref, band=BAND):
    """North-star criterion (mask-pixel IoU >= 0.999) on the pixels the fp32 reference is confident about.

    With seeded random weights a mask is a smooth random field whose logits hover around the threshold over a sizeable
    area, and the raw IoU of two correct implementations is 0.95-0.99 (the unmodified reference under bf16 autocast
    scores 0.88-0.99 against its own fp32 run, tests/golden/ref_bf16_deviation.json).  The band is FIXED relative to the
    reference mask (a fraction of the mask's rms logit; 0.15 x 6.8 ~ 1 logit unit at the large model = probabilities
    outside 0.27-0.73), not derived from the measured error, so the statement is not true by construction: every
    pixel outside the band must keep its sign, however the error is distributed (it is heavy-tailed: p99.9 is ~6x the
    rms error; the worst flipped pixel seen sits at 0.037 x rms, profiles/r2_parity_large16_steady_state.txt).
    Returns (IoU over confident pixels, fraction of confident pixels, diagnostics of the flipped pixels or None)."""
    got, ref = got.double(), ref.double()
    ref_rms = ref.pow(2).mean().sqrt().item()
    conf = ref.abs() > band * ref_rms
    a, b = (got > 0) & conf, (ref > 0) & conf
    u = (a | b).sum().item()
    iou = 1.0 if u == 0 else (a & b).sum().item() / u
    flip = (got > 0) != (ref > 0)
    info = None
    if flip.any():
        err = (got - ref).abs()
        info = (int(flip.sum()), err.pow(2).mean().sqrt().item(), ref.abs()[flip].max().item(), err[flip].max().item(),
                err.max().item(), torch.quantile(err.flatten()[::16].float(), 0.999).item(), ref_rms)
    return iou, conf.double().mean().item(), info


def _run(engine, cfg, vid, frames, log_best, fill_hole_area=8, **init_kw):
    from detsam2_b200.predictor import SAM2VideoPredictor
    pred = SAM2VideoPredictor(engine, fill_hole_area=fill_hole_area)
    orig = engine.sam_heads

    def sam_heads(*a, **kw):
        out = orig(*a, **kw)
        log_best(out)
        return out

    engine.sam_heads = sam_heads
    outs = {}
    try:
        with torch.inference_mode():
            st = pred.init_state(frames, **init_kw)
            pred.add_new_boxes(st, 0, {oid: np.asarray(b, np.float32) for oid, b in vid.boxes(0).items()})
            for f, _, m in pred.propagate_in_video(st):
                o = st["output_dict"]["cond_frame_outputs"].get(f) or st["output_dict"]["non_cond_frame_outputs"][f]
                outs[f] = {"video": m.float().cpu(), "pred_masks": o["pred_masks"].float().cpu(),
                           "obj_ptr": o["obj_ptr"].float().cpu(), "score": o["object_score_logits"].float().cpu(),
                           "maskmem": o["maskmem_features"].float().cpu()}
    finally:
        engine.sam_heads = orig
    return outs, st


def _decisive_iou_head(sd):
    sd = dict(sd)
    k = "sam_mask_decoder.iou_prediction_head.layers.2.bias"
    sd[k] = sd[k].clone() + torch.tensor([0.0, 2.0, 0.0, -2.0])
    return sd


def test_large_16_objects_steady_state_bank_vs_fp32_oracle():
    from detsam2_b200.config import get_config
    from detsam2_b200.engine import CudaEngine
    from detsam2_b200.synthetic import BilliardVideo
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg = get_config("large")
    sd = _decisive_iou_head(synthetic_state_dict(cfg, 0))
    B, nfr = 16, 19
    vid = BilliardVideo(num_objects=B, height=1024, width=1024, num_frames=nfr, seed=21)
    frames = list(vid.frames())

    eng = CudaEngine(cfg, sd, device="cuda:0")
    eng.kernel_timers = []                      # records the key count N of every cross-attention launch
    cuda_best = []
    got, _ = _run(eng, cfg, vid, frames, lambda out: cuda_best.append(out["_best_idx"].clone().cpu())
                  if "_best_idx" in out else None, fill_hole_area=0)
    keys = sorted({meta["N"] for _, _, _, meta in eng.kernel_timers})
    eng.kernel_timers = None
    del eng
    torch.cuda.empty_cache()

    ora = O.OracleEngine(cfg, sd, fill_holes=False, device="cuda:0")
    O.DECISION_LOG = []
    try:
        ref, _ = _run(ora, cfg, vid, frames, lambda out: None, fill_hole_area=0)
    finally:
        decisions, O.DECISION_LOG = O.DECISION_LOG, None
    del ora
    torch.cuda.empty_cache()

    # the bank reached its steady state: 7 stored frames + 16 pointers
    assert keys[-1] == 7 * 4096 + 4 * 16 == 28736, keys
    tracked = [d for d in decisions if "multimask_top2_margin" in d]
    assert len(tracked) == nfr - 1 and len(cuda_best) >= nfr - 1
    cuda_tracked = cuda_best[-(nfr - 1):]       # the batched box prompt on frame 0 is a single-mask decode

    lines, bad, diag = [], [], []
    for i, f in enumerate(range(1, nfr)):
        ob = np.asarray(tracked[i]["best"])
        cb = cuda_tracked[i].numpy().astype(np.int64) - 1     # engine indexes the 4 decoder masks, oracle masks 1..3
        margin = np.asarray(tracked[i]["multimask_top2_margin"])
        if margin.min() < MIN_MARGIN:
            bad.append(f"frame {f}: oracle top-2 IoU margin {margin.min():.4f} < {MIN_MARGIN} despite the decisive head")
        if not np.array_equal(ob, cb):
            bad.append(f"frame {f}: multimask choices differ: oracle {ob.tolist()} cuda {cb.tolist()}")
            break
        g, r = got[f], ref[f]
        row = {k: _rel(g[k], r[k]) for k in ("pred_masks", "video", "obj_ptr", "maskmem")}
        row["score"] = (g["score"] - r["score"]).abs().max().item()
        ious, ious_dec, frac_dec = [], [], []
        for o in range(B):
            ious.append(_iou(g["video"][o], r["video"][o]))
            iou_d, frac, flip_info = _confident_iou(g["video"][o], r["video"][o])
            ious_dec.append(iou_d)
            frac_dec.append(frac)
            if flip_info is not None:
                diag.append((f, o) + flip_info)
        lines.append(f"frame {f:2d}: rel-rms pred_masks {row['pred_masks']:.4f} video {row['video']:.4f} obj_ptr "
                     f"{row['obj_ptr']:.4f} maskmem {row['maskmem']:.4f}  score |d| {row['score']:.4f}  raw IoU mean "
                     f"{np.mean(ious):.4f} min {np.min(ious):.4f}  confident-pixel IoU min {np.min(ious_dec):.5f} "
                     f"({100 * np.mean(frac_dec):.1f} % of pixels confident)  min oracle margin {margin.min():.4f}")
        # bf16 operands / fp32 accumulation against fp32, through the rolling memory bank: measured 0.005-0.010 /
        # 0.0077 / 0.010-0.014 / <= 0.013 (profiles/r2_parity_large16_steady_state.txt); bounds = 3x / 2x of that
        if row["pred_masks"] > 0.03 or row["maskmem"] > 0.015 or row["obj_ptr"] > 0.04 or row["score"] > 0.05:
            bad.append(lines[-1])
        if np.min(ious_dec) < 0.999:
            bad.append(lines[-1])
    n_frames_compared = len(lines)
    diag.sort(key=lambda d: -d[4] / max(d[8], 1e-12))
    lines.append("flipped pixels, worst objects (frame, object, flips, rms err, max |ref| among flips, max err among flips, "
                 "max err, p99.9 err, ref rms):")
    lines += [f"  f{d[0]} o{d[1]}: {d[2]} flips, rms {d[3]:.4f}, |ref|max@flip {d[4]:.3f}, err@flip {d[5]:.3f}, err max {d[6]:.3f}, "
              f"p99.9 {d[7]:.3f}, ref rms {d[8]:.2f}" for d in diag[:12]]
    report = "\n".join(lines + [f"key counts seen: {keys}"])
    print(report)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_large16_steady_state.txt"), "w") as fh:
        fh.write(report + "\n")
    assert not bad, "\n".join(bad)
    assert n_frames_compared == nfr - 1


def test_oracle_cuda_path_equals_cpu_path():
    """The fp32 oracle on torch's CUDA kernels (TF32 off) is the same function as the CPU oracle that is pinned against
    the reference fixtures: tiny model, 2 objects, two tracked frames, every stored output."""
    from detsam2_b200.config import get_config
    from detsam2_b200.synthetic import BilliardVideo
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=512)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=2, height=192, width=256, num_frames=3, seed=5)
    frames = list(vid.frames())
    torch.set_num_threads(os.cpu_count() or 1)
    a, _ = _run(O.OracleEngine(cfg, sd, fill_holes=True, device="cuda:0"), cfg, vid, frames, lambda out: None)
    b, _ = _run(O.OracleEngine(cfg, sd, fill_holes=True), cfg, vid, frames, lambda out: None)
    for f in (1, 2):
        for k in ("pred_masks", "video", "obj_ptr", "maskmem"):
            assert _rel(a[f][k], b[f][k]) < (8e-3 if k == "maskmem" else 2e-4), (f, k, _rel(a[f][k], b[f][k]))
        assert (a[f]["score"] - b[f]["score"]).abs().max().item() < 1e-3


@pytest.mark.parametrize("offload_state", [True])
def test_offload_state_to_cpu_on_cuda_engine(offload_state):
    """`offload_state_to_cpu=True` (svp:44-53, the default of init_preloading_state svp:123-156): stored maskmem features
    and masks live on the host and come back over PCIe for every memory read.  Same results as the device-resident
    state, bit for bit (the storage device does not enter the arithmetic)."""
    from detsam2_b200.config import get_config
    from detsam2_b200.engine import CudaEngine
    from detsam2_b200.synthetic import BilliardVideo
    from detsam2_b200.weights import synthetic_state_dict
    cfg = get_config("tiny", image_size=512)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=3, height=192, width=256, num_frames=6, seed=8)
    frames = list(vid.frames())
    eng = CudaEngine(cfg, sd, device="cuda:0")
    dev, st_dev = _run(eng, cfg, vid, frames, lambda out: None)
    off, st_off = _run(eng, cfg, vid, frames, lambda out: None, offload_state_to_cpu=True)
    assert st_off["storage_device"].type == "cpu" and st_dev["storage_device"].type == "cuda"
    o = st_off["output_dict"]["non_cond_frame_outputs"][3]
    assert o["maskmem_features"].device.type == "cpu" and o["pred_masks"].device.type == "cpu"
    assert o["obj_ptr"].device.type == "cuda"                 # pointers stay on the compute device (svp:1357-1363)
    for f in dev:
        for k in dev[f]:
            assert torch.equal(dev[f][k], off[f][k]), (f, k)


@pytest.mark.parametrize("name", ["large_1024", "bplus_720p"])
def test_cuda_engine_matches_reference_fixture_at_full_size(name):
    """CUDA engine against outputs of the UNMODIFIED reference (fp32) at the full-size models: tests/golden/large_1024.npz
    (Hiera-L, 1024^2) and bplus_720p.npz (base_plus, 1280x720 frames), generated by oracle/gen_golden.py.  Hole filling is
    off on both sides (the reference's CPU path skips it, misc.py:389-391)."""
    from detsam2_b200.engine import CudaEngine
    from detsam2_b200.predictor import SAM2VideoPredictor
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import scenarios
    gold, _ = scenarios.load_golden(name)
    cfg = scenarios.scenario_config(name)
    eng = CudaEngine(cfg, synthetic_state_dict(cfg, 0), device="cuda:0")
    got = scenarios.SCENARIOS[name](SAM2VideoPredictor(eng, fill_hole_area=0))
    torch.cuda.synchronize()
    assert set(got) == set(gold)
    lines, bad = [], []
    # bf16 operands / fp32 accumulation against the fp32 reference; measured values in profiles/r2_parity_fullsize_fixtures.txt
    # (the second tracked frame reads a bank that already carries bf16 noise: wider band, as in test_configs_gpu.py)
    bounds = {"pred_masks": 0.06, "obj_ptr": 0.08, "maskmem_features": 0.02, "object_score_logits": 0.02}
    late = {"pred_masks": 0.10}
    for k in sorted(gold):
        r = gold[k]
        kind = k.rsplit(".", 1)[-1]
        if np.issubdtype(r.dtype, np.integer) and kind != "masks_packed":
            assert np.array_equal(got[k], r), k
            continue
        if kind == "masks_packed":
            # reported only: on near-degenerate random-weight masks the raw IoU says little (see _confident_iou); the
            # logits these bits come from are compared below
            lines.append(f"{name} {k}: raw mask IoU {scenarios.packed_mask_iou(got[k], r):.4f}")
            continue
        g64, r64 = got[k].astype(np.float64), r.astype(np.float64)
        rms = np.sqrt(np.mean((g64 - r64) ** 2))
        rel = rms / max(np.sqrt(np.mean(r64 ** 2)), 1e-12)
        msg = f"{name} {k}: rel-rms {rel:.4f}"
        if kind == "pred_masks":
            worst, frac = 1.0, []
            for o in range(r.shape[0]):
                iou_c, fr, _ = _confident_iou(torch.from_numpy(g64[o]), torch.from_numpy(r64[o]), BAND_FIXTURE)
                worst = min(worst, iou_c)
                frac.append(fr)
            a, b = g64 > 0, r64 > 0
            msg += (f", raw IoU {(a & b).sum() / max((a | b).sum(), 1):.4f}, confident-pixel IoU min {worst:.5f} "
                    f"({100 * np.mean(frac):.1f} % of pixels confident)")
            if worst < 0.999:
                bad.append(msg)
        lines.append(msg)
        if rel > (late.get(kind, bounds[kind]) if ".f2." in k else bounds[kind]):
            bad.append(msg)
    print("\n".join(lines))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_fixture_{name}.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    assert not bad, "\n".join(bad)
