"""GPU bring-up of the CudaEngine: every seam against the CPU oracle on the same seeded inputs, then a
short end-to-end predictor run with both engines.  Writes gpurun_out/engine_check_<cfg>.json.

usage: python tools/engine_check.py [tiny|large|base_plus] [num_objects]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from detsam2_b200.config import get_config  # noqa: E402
from detsam2_b200.engine import CudaEngine  # noqa: E402
from detsam2_b200.predictor import SAM2VideoPredictor  # noqa: E402
from detsam2_b200.synthetic import BilliardVideo  # noqa: E402
from detsam2_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import sam2_oracle as O  # noqa: E402


def stats(name, got, ref, res):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    d = (got - ref).abs()
    r = {"max_abs": d.max().item(), "mean_abs": d.mean().item(), "ref_absmax": ref.abs().max().item(),
         "ref_rms": ref.pow(2).mean().sqrt().item(), "nan": bool(torch.isnan(got).any())}
    r["rel_rms"] = (d.pow(2).mean().sqrt() / max(r["ref_rms"], 1e-12)).item()
    res[name] = r
    print(f"{name:34s} max {r['max_abs']:.3e} mean {r['mean_abs']:.3e} refmax {r['ref_absmax']:.3e} relrms {r['rel_rms']:.3e}"
          f"{' NaN!' if r['nan'] else ''}", flush=True)


def iou(a, b):
    a, b = (a > 0).cpu(), (b > 0).cpu()
    u = (a | b).sum().item()
    return 1.0 if u == 0 else (a & b).sum().item() / u


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = get_config(name)
    sd = synthetic_state_dict(cfg, 0)
    res = {"config": name, "B": B}
    t0 = time.time()
    cu = CudaEngine(cfg, sd)
    orc = O.OracleEngine(cfg, sd, fill_holes=True)
    print("engines built", time.time() - t0, flush=True)
    S = cfg.image_size
    vid = BilliardVideo(num_objects=B, height=S, width=S, num_frames=6, seed=0)
    frames = list(vid.frames())
    imgs, _, _ = O.load_frames(frames, S)
    T = cfg.feat_size ** 2
    with torch.inference_mode():
        # ---- seam 1 ----
        t0 = time.time()
        fo = orc.encode_image(imgs[0])
        print("oracle encode", time.time() - t0, flush=True)
        fc = cu.encode_image(imgs[0])
        torch.cuda.synchronize()
        stats("enc.vis(64^2)", fc.vis_f32, fo["fpn"][2][0].flatten(1).t(), res)
        stats("enc.feat_s1(128^2)", fc.feat_s1, fo["fpn"][1][0].flatten(1).t(), res)
        stats("enc.feat_s0(256^2)", fc.feat_s0, fo["fpn"][0][0].flatten(1).t(), res)
        # cuda feats carrying the ORACLE values, to test the later seams in isolation
        from detsam2_b200.engine import FrameFeats
        fx = FrameFeats(fo["fpn"][2][0].flatten(1).t().contiguous().cuda(),
                        fo["fpn"][2][0].flatten(1).t().contiguous().cuda().bfloat16(),
                        fo["fpn"][0][0].flatten(1).t().contiguous().cuda(),
                        fo["fpn"][1][0].flatten(1).t().contiguous().cuda())
        # ---- seam 3 on an init-cond frame (box prompt, B = 1) ----
        box = vid.boxes(0)[0]
        pts = torch.tensor([[[box[0], box[1]], [box[2], box[3]]]], dtype=torch.float32)
        lab = torch.tensor([[2, 3]], dtype=torch.int32)
        po = orc.condition_on_memory(fo, 1, 0, True, {}, 6, False, None)
        pc = cu.condition_on_memory(fx, 1, 0, True, {}, 6, False, None)
        stats("init.pix_feat", pc[0], po[0].flatten(1).t(), res)
        so = orc.sam_heads(po, fo, 1, pts, lab, None, False)
        sc = cu.sam_heads(pc, fx, 1, pts.cuda(), lab.cuda(), None, False)
        torch.cuda.synchronize()
        for k in ("pred_masks", "ious", "obj_ptr", "object_score_logits"):
            stats("box." + k, sc[k], so[k], res)
        res["box.mask_iou"] = iou(sc["pred_masks"], so["pred_masks"])
        print("box mask iou", res["box.mask_iou"], "pos frac", (so["pred_masks"] > 0).float().mean().item(), flush=True)
        # ---- seam 4 ----
        lowB = torch.randn(B, 1, 4 * cfg.feat_size, 4 * cfg.feat_size) * 4
        scoreB = torch.tensor([[3.0]] * B)
        if B > 1:
            scoreB[1, 0] = -2.0
        for frompts in (True, False):
            mo, pe_o = orc.encode_memory(fo, B, lowB, scoreB, frompts)
            mc, pe_c = cu.encode_memory(fx, B, lowB.cuda(), scoreB.cuda(), frompts)
            torch.cuda.synchronize()
            stats(f"memenc(pts={frompts})", mc, mo, res)
            stats(f"memenc.pos(pts={frompts})", pe_c[0], pe_o[0], res)
        # ---- seam 2 + 3 on a tracking step with a synthetic bank ----
        def entry(seed):
            g = torch.Generator().manual_seed(seed)
            mf = (torch.randn(B, 64, cfg.feat_size, cfg.feat_size, generator=g) * 2).bfloat16()
            ptr = torch.randn(B, 256, generator=g)
            return mf, ptr

        od_o = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        od_c = {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}
        pos_o = [O.sine_pe_2d(64, cfg.feat_size, cfg.feat_size)[None].expand(B, -1, -1, -1)]
        for t, key in ((0, "cond_frame_outputs"), (1, "non_cond_frame_outputs"), (2, "non_cond_frame_outputs"),
                       (3, "non_cond_frame_outputs")):
            mf, ptr = entry(t)
            od_o[key][t] = {"maskmem_features": mf, "maskmem_pos_enc": pos_o, "obj_ptr": ptr}
            od_c[key][t] = {"maskmem_features": mf.cuda(), "maskmem_pos_enc": None, "obj_ptr": ptr.cuda()}
        t0 = time.time()
        po = orc.condition_on_memory(fo, B, 4, False, od_o, 6, False, None)
        print("oracle memattn", time.time() - t0, flush=True)
        pc = cu.condition_on_memory(fx, B, 4, False, od_c, 6, False, None)
        torch.cuda.synchronize()
        stats("memattn.pix_feat", pc, po.flatten(2).transpose(1, 2), res)
        # decoder on the ORACLE's conditioned features (isolation) — multimask tracking mode
        so = orc.sam_heads(po, fo, B, None, None, None, True)
        sc = cu.sam_heads(po.flatten(2).transpose(1, 2).contiguous().cuda(), fx, B, None, None, None, True)
        torch.cuda.synchronize()
        for k in ("pred_masks", "ious", "obj_ptr", "object_score_logits"):
            stats("track." + k, sc[k], so[k], res)
        res["track.mask_iou"] = iou(sc["pred_masks"], so["pred_masks"])
        print("track mask iou", res["track.mask_iou"], flush=True)
        stats("track.multimasks", sc["_all_masks"][:, 1:4], so["_multimasks"], res)
        stats("track.multi_ious", sc["_all_ious"][:, 1:4], so["ious"], res)
        best_o = so["ious"].argmax(-1)
        top2 = so["ious"].topk(2, dim=-1).values
        print("oracle best", best_o.tolist(), "cuda best", (sc["_best_idx"].cpu() - 1).tolist(),
              "oracle top-2 iou margin", (top2[:, 0] - top2[:, 1]).tolist(), flush=True)
        res["track.multimask_iou"] = iou(sc["_all_masks"][:, 1:4], so["_multimasks"])
        print("track multimask iou", res["track.multimask_iou"], flush=True)
        # ---- seam 5 ----
        m = torch.randn(B, 1, 256, 256)
        stats("fill_holes", cu.fill_holes(m.cuda(), 8), orc.fill_holes(m, 8), res)
        stats("resize", cu.resize_masks(m.cuda(), 300, 420), orc.resize_masks(m, 300, 420), res)

        # ---- end to end: same predictor, both engines ----
        if os.environ.get("DS2_SKIP_E2E") != "1":
            outs = {}
            for tag, eng in (("cuda", cu), ("oracle", orc)):
                pred = SAM2VideoPredictor(eng, fill_hole_area=8)
                st = pred.init_state(frames)
                for oid, bx in vid.boxes(0).items():
                    pred.add_new_points_or_box(st, 0, oid, box=bx)
                t0 = time.time()
                outs[tag] = {f: m.detach().float().cpu() for f, _, m in pred.propagate_in_video(st)}
                torch.cuda.synchronize()
                print(tag, "propagate", time.time() - t0, flush=True)
                outs[tag + "_state"] = st
            e2e = {}
            for f in outs["oracle"]:
                a, b = outs["cuda"][f], outs["oracle"][f]
                e2e[f] = {"iou": iou(a, b), "max_abs": (a - b).abs().max().item(),
                          "pos_frac": (b > 0).float().mean().item()}
                print("e2e frame", f, e2e[f], flush=True)
            res["e2e"] = e2e
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"engine_check_{name}.json"), "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
