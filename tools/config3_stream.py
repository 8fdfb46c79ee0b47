"""BASELINE.json configs[2] at full length: sam2.1_hiera_base_plus, 1280x720 frames, a preload memory bank pickled from a
10-frame all-conditioning-frame run (detect_interval = 1, max_inference_state_frames = -1, det_sam2_RT.py:67-68), then
a 2000-frame stream against that bank through the constant-memory window (VideoProcessor: K = 30 frames per chunk,
reverse window M = 60, state window S = 60; detect_interval = -1: the bank is the only prompt source, or 30).

Reports video frames/s (wall clock: ingest, preflight, reverse re-tracking, release and the D2H hand-off of the boolean
masks included; frame synthesis excluded by pre-generating each chunk), track steps/s, and the device-memory trace per
chunk; asserts that allocated memory after chunk 3 never grows.

usage: python tools/config3_stream.py [--frames 2000] [--objects 16] [--detect-interval -1] [--bank .pkl|.ds2bank]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--model", default="base_plus")
    ap.add_argument("--hw", type=int, nargs=2, default=(720, 1280))
    ap.add_argument("--pre", type=int, default=10, help="frames of the preload bank")
    ap.add_argument("--detect-interval", type=int, default=-1)
    ap.add_argument("--bank", default=".pkl", choices=[".pkl", ".ds2bank"])
    ap.add_argument("--state-on-host", action="store_true", help="the reference's offload_state_to_cpu=True for the loaded bank")
    args = ap.parse_args()
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.synthetic import BilliardVideo, GroundTruthDetector
    from detsam2_b200.video_processor import VideoProcessor
    yaml = {"tiny": "t", "small": "s", "base_plus": "b+", "large": "l"}[args.model]
    dev = torch.device("cuda", 0)
    H, W = args.hw
    K, M, S = 30, 60, 60
    predictor = build_sam2_video_predictor(f"configs/sam2.1/sam2.1_hiera_{yaml}.yaml", device=dev, seed=0,
                                           feature_cache_frames=M + 4)
    vid = BilliardVideo(num_objects=args.objects, height=H, width=W, num_frames=args.pre + args.frames, seed=9)
    tmp = tempfile.mkdtemp()
    bank = os.path.join(tmp, "bank" + args.bank)
    with torch.inference_mode():
        t0 = time.perf_counter()
        vp = VideoProcessor(predictor=predictor, detector=GroundTruthDetector(vid, detect_interval=1),
                            frame_buffer_size=args.pre, detect_interval=1, max_frame_num_to_track=args.pre,
                            max_inference_state_frames=-1, save_inference_state_path=bank, skip_classes=set())
        vp.run(frames=(vid.frame(t) for t in range(args.pre)))
        bank_s = time.perf_counter() - t0
        bank_mb = os.path.getsize(bank) / 2 ** 20
        del vp
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

        det = None if args.detect_interval == -1 else GroundTruthDetector(vid, detect_interval=args.detect_interval,
                                                                           first_frame=args.pre + (-args.pre) % args.detect_interval)
        vp = VideoProcessor(predictor=predictor, detector=det, frame_buffer_size=K, detect_interval=args.detect_interval,
                            max_frame_num_to_track=M, max_inference_state_frames=S, load_inference_state_path=bank,
                            skip_classes=set(), offload_state_to_cpu=True if args.state_on_host else None)
        steps = [0]
        orig_step = predictor._run_single_frame_inference

        def counted(*a, **kw):
            if not kw.get("is_init_cond_frame", False):
                steps[0] += 1
            return orig_step(*a, **kw)

        predictor._run_single_frame_inference = counted
        trace = []
        orig = vp.Detect_and_SAM2_inference

        def chunk(frame_idx):
            orig(frame_idx)
            torch.cuda.synchronize()
            st = vp.inference_state
            trace.append({"frame": frame_idx, "allocated_mb": round(torch.cuda.memory_allocated() / 2 ** 20, 1),
                          "reserved_mb": round(torch.cuda.memory_reserved() / 2 ** 20, 1),
                          "frames_in_state": len(st["images_idx"]),
                          "non_cond_outputs": len(st["output_dict"]["non_cond_frame_outputs"]),
                          "cond_outputs": len(st["output_dict"]["cond_frame_outputs"])})
            # results are handed to the consumer chunk by chunk in a real stream: do not let them pile up on the host
            for t in [t for t in vp.video_segments if t < frame_idx - 2 * K]:
                vp.video_segments.pop(t)

        vp.Detect_and_SAM2_inference = chunk
        # frames are synthesised chunk by chunk OUTSIDE the timed sections
        synth = 0.0
        torch.cuda.synchronize()
        t_start = time.perf_counter()
        vp.load_inference_state_path = bank
        gen_t = [0.0]

        def frames():
            for t in range(args.pre, args.pre + args.frames):
                a = time.perf_counter()
                f = vid.frame(t)
                gen_t[0] += time.perf_counter() - a
                yield f

        segs_total = [0]
        vp.run(frames=frames())
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_start - gen_t[0]
    alloc = [c["allocated_mb"] for c in trace]
    base = alloc[2] if len(alloc) > 2 else alloc[-1]
    flat = max(alloc[2:]) <= base + 32.0 if len(alloc) > 2 else True
    out = {"config": f"BASELINE configs[2]: sam2.1_hiera_{args.model}, {W}x{H} frames, {args.objects} objects, preload bank of "
                     f"{args.pre} conditioning frames ({args.bank}, {bank_mb:.0f} MB, built in {bank_s:.1f} s), {args.frames}-frame "
                     f"stream, K={K} M={M} S={S}, detect_interval={args.detect_interval}, loaded bank kept on the "
                     f"{'host (reference default)' if args.state_on_host else 'device'}",
           "video_fps": round(args.frames / wall, 2), "track_steps_per_s": round(steps[0] / wall, 2), "track_steps": steps[0],
           "wall_s": round(wall, 2), "frame_synthesis_s_excluded": round(gen_t[0], 2), "chunks": len(trace),
           "allocated_mb_after_chunk3": base, "allocated_mb_max_after_chunk3": max(alloc[2:]) if len(alloc) > 2 else None,
           "allocated_mb_last": alloc[-1], "reserved_mb_max": max(c["reserved_mb"] for c in trace),
           "peak_allocated_mb": round(torch.cuda.max_memory_allocated() / 2 ** 20, 1), "memory_flat": flat,
           "state_window_max": {"frames_in_state": max(c["frames_in_state"] for c in trace[2:]) if len(trace) > 2 else None,
                                "non_cond_outputs": max(c["non_cond_outputs"] for c in trace[2:]) if len(trace) > 2 else None},
           "chunk_phase_s": {k: round(v, 3) for k, v in vp.timings.items()},
           "trace_every_8th_chunk": trace[::8]}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "config3_stream.json"), "w") as f:
        f.write(json.dumps(out) + "\n")
    assert flat, ("device memory grew over the stream", alloc)


if __name__ == "__main__":
    main()
