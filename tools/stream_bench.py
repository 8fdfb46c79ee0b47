"""Det-SAM2 stream mode at BASELINE configs[1] size (drive mode B): this repo's
VideoProcessor (det_sam2_RT.py semantics) on sam2.1_hiera_large, 1024x1024, 16 objects, ground-truth boxes as the
detector: K = frame_buffer_size = 30, detect every 30 frames, reverse window M = 60, state window S = 60.
fps = video frames / wall time (prompting, preflight, reverse re-tracking, release, D2H of the boolean masks
included; frame synthesis excluded); also track-steps/s.

usage: python tools/stream_bench.py [--frames 150] [--objects 16] [--model large]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--model", default="large")
    ap.add_argument("--cache", type=int, default=64, help="backbone-feature cache (frames); the reference keeps 1")
    ap.add_argument("--skip-single-cache", action="store_true", help="only run the --cache setting")
    ap.add_argument("--frames-on-device", type=int, default=1, help="keep the session's fp16 frames in HBM")
    args = ap.parse_args()
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.synthetic import BilliardVideo, GroundTruthDetector
    from detsam2_b200.video_processor import VideoProcessor
    yaml = {"tiny": "t", "small": "s", "base_plus": "b+", "large": "l"}[args.model]
    dev = torch.device("cuda", 0)
    res = {}
    for cache in ((args.cache,) if args.skip_single_cache else (1, args.cache)):
        predictor = build_sam2_video_predictor(f"configs/sam2.1/sam2.1_hiera_{yaml}.yaml", device=dev, seed=0,
                                               feature_cache_frames=cache)
        S = predictor.cfg.image_size
        vid = BilliardVideo(num_objects=args.objects, height=S, width=S, num_frames=args.frames, seed=0)
        frames = [vid.frame(t) for t in range(args.frames)]
        steps = [0]
        orig = predictor._run_single_frame_inference

        def counted(*a, **kw):
            if not kw.get("is_init_cond_frame", False):
                steps[0] += 1
            return orig(*a, **kw)

        predictor._run_single_frame_inference = counted
        # coarse host-side wall time per predictor entry point (GPU work is asynchronous: a phase that synchronises
        # also absorbs the device time still queued before it)
        phase = {}

        def timed(name):
            fn = getattr(predictor, name)

            def wrap(*a, **kw):
                t = time.perf_counter()
                try:
                    return fn(*a, **kw)
                finally:
                    phase[name] = phase.get(name, 0.0) + time.perf_counter() - t
            setattr(predictor, name, wrap)

        for name in ("init_state", "update_state", "add_new_boxes", "add_new_points_or_box",
                     "propagate_in_video_preflight", "release_old_frames"):
            if hasattr(predictor, name):
                timed(name)
        for rep in range(2):   # first pass warms up graphs / allocator; second is timed
            steps[0] = 0
            gr = predictor.engine.graphs
            g0 = (gr.captures, gr.replays, sum(gr.hits.values()))
            phase.clear()
            vp = VideoProcessor(predictor=predictor, detector=GroundTruthDetector(vid, detect_interval=30),
                                frame_buffer_size=30, detect_interval=30, max_frame_num_to_track=60,
                                max_inference_state_frames=60, skip_classes=set(),
                                frames_on_device=bool(args.frames_on_device))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.inference_mode():
                segs = vp.run(frames=iter(frames))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        missing = [t for t in range(args.frames) if t not in segs]
        short = {t: len(segs[t]) for t in segs if len(segs[t]) != args.objects}
        if missing or short:
            print(f"WARNING cache={cache}: frames without a result {missing[:20]}; frames with fewer objects {dict(list(short.items())[:10])}",
                  flush=True)
        res[f"feature_cache_{cache}"] = {"video_fps": round(args.frames / dt, 2), "track_steps_per_s": round(steps[0] / dt, 2),
                                         "track_steps": steps[0], "wall_s": round(dt, 3),
                                         "frames_with_result": len(segs),
                                         "host_phase_s": {k: round(v, 3) for k, v in phase.items()},
                                         "chunk_phase_s": {k: round(v, 3) for k, v in vp.timings.items()},
                                         "graphs_in_timed_pass": {"captured": gr.captures - g0[0], "replays": gr.replays - g0[1],
                                                                  "eager_seam_runs": sum(gr.hits.values()) - g0[2],
                                                                  "graphs_total": len(gr.graphs)}}
        del predictor, vp
        torch.cuda.empty_cache()
    line = {"mode": "Det-SAM2 stream (VideoProcessor: K=30, detect every 30, M=60, S=60, reverse)", "model": args.model,
            "objects": args.objects, "frames": args.frames, "frames_on_device": bool(args.frames_on_device),
            "host_ingest": os.environ.get("DS2_HOST_INGEST", "0") == "1", "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), **res}
    print(json.dumps(line))
    with open(os.path.join(ROOT, "gpurun_out", "stream_bench.json"), "w") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
