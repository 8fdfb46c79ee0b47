"""Device time of each engine seam inside a steady-state tracked frame (large, 16 objects, CUDA-graph replays), measured
with CUDA events around the engine calls on the launching stream.  Complements the ncu launch list (cold, serialised)
with warm in-pipeline numbers.

usage: python tools/seam_times.py [--objects 16] [--steps 24] [--enc-batch 4]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--prefill", type=int, default=20)
    ap.add_argument("--enc-batch", type=int, default=4)
    args = ap.parse_args()
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.synthetic import BilliardVideo
    pred = build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_l.yaml", device="cuda", seed=0,
                                      encoder_batch_frames=args.enc_batch)
    # per-seam accounting needs the seams one after the other on one stream: the encoder passes stay on the tracker's stream
    # here (in the product they run on the engine's encoder stream and overlap the tracker, predictor.encoder_overlap)
    pred.encoder_overlap = False
    eng = pred.engine
    S = pred.cfg.image_size
    n = 1 + args.prefill + args.steps
    vid = BilliardVideo(num_objects=args.objects, height=S, width=S, num_frames=n, seed=0)
    st = pred.init_state(list(vid.frames()), offload_video_to_cpu=False)
    for oid, box in vid.boxes(0).items():
        pred.add_new_points_or_box(st, 0, oid, box=box)
    events = collections.defaultdict(list)
    recording = [False]

    def wrap(name):
        fn = getattr(eng, name)

        def inner(*a, **kw):
            if not recording[0]:
                return fn(*a, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            events[name].append((e0, e1))
            return out
        setattr(eng, name, inner)

    for name in ("encode_image", "encode_images", "condition_on_memory", "sam_heads", "encode_memory", "fill_holes",
                 "resize_masks"):
        wrap(name)
    gen = pred.propagate_in_video(st)
    for _ in range(1 + args.prefill):
        next(gen)
    pred.drop_encoded_ahead()
    torch.cuda.synchronize()
    recording[0] = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        next(gen)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1)
    out = {"objects": args.objects, "steps": args.steps, "encoder_batch_frames": pred.encoder_batch_frames,
           "encoder_overlap": False,
           "ms_per_step": round(total / args.steps, 3), "seams": {}}
    acc = 0.0
    for name, evs in events.items():
        ms = sum(a.elapsed_time(b) for a, b in evs)
        acc += ms
        out["seams"][name] = {"calls": len(evs), "ms_per_call": round(ms / len(evs), 3), "ms_per_step": round(ms / args.steps, 3)}
    out["between_seams_ms_per_step"] = round((total - acc) / args.steps, 3)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"seam_times_e{pred.encoder_batch_frames}.json"), "w") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
