"""TEST INFRASTRUCTURE — API-level scenarios driven through *any* object that presents the reference's
``SAM2VideoPredictor`` interface (/root/reference/sam2/sam2_video_predictor.py): the unmodified
reference (through oracle/ref_shim.py, build container only), this repo's predictor over the CPU
oracle engine, or this repo's predictor over the sm_100a CUDA engine.  Because the predictor API is
the drop-in boundary, ONE driver serves all three; parity = the recorded arrays agree.

Every scenario returns ``{name: np.ndarray}``.  Nothing in the product imports this module.

Scenarios
  offline   BASELINE.json configs[0]: sam2.1_hiera_tiny, 4 synthetic 512x512 frames, 1 box prompt on
            frame 0 -> init_state -> add_new_points_or_box -> propagate_in_video.
  stream    Det-SAM2 streaming semantics (det_sam2_RT.py:342-411) on a reduced-resolution tiny model:
            init_state(chunk) / update_state(chunk), box prompts on the last frame of each chunk,
            reverse propagate with max_frame_num_to_track, an online NEW object id while tracking
            (svp:250-327), release_old_frames with release_images (svp:1215-1277).
  mask_prompt  add_new_mask with a disc, a box on the same frame and an empty mask, then tracking (svp:527-600).
  points_api  click prompts (one click / positive + negative / box + click), clear_all_prompts_in_frame,
            remove_object, reset_state and a new object id afterwards (svp:344-520, 1061-1170, 1438-1553).
  refine_click  three accumulating clicks on one frame: the previous logits return as a dense prompt (svp:463-480).
  preload   preload memory bank (det_sam2_RT.py:489-503, svp:123-156): every frame of a short clip is
            a conditioning frame, state is pickled, re-loaded, init_preloading_state, new frames
            appended with update_state and tracked against the bank.
"""
import io
import pickle

import numpy as np
import torch

from detsam2_b200.config import get_config
from detsam2_b200.synthetic import BilliardVideo

# reduced resolution for the plumbing-heavy scenarios (T = 32*32 tokens): same code paths, 4x cheaper
SMALL = dict(image_size=512)


# full-size model scenarios (fixture name -> model, frame size, objects, frames, seed): the oracle is pinned against the
# unmodified reference at the shapes the headline lives on — Hiera-L (16-token windows, global blocks 23/33/43,
# head_dim 72) at 1024^2 and base_plus (hieradet.py:177-202 defaults, 14/7-token zero-padded windows) on 720p frames
FULLSIZE = {"large_1024": dict(model="large", height=1024, width=1024, num_objects=1, num_frames=3, seed=31),
            "bplus_720p": dict(model="base_plus", height=720, width=1280, num_objects=2, num_frames=3, seed=2)}


def scenario_config(name):
    if name in FULLSIZE:
        return get_config(FULLSIZE[name]["model"])
    if name == "offline":
        return get_config("tiny")
    return get_config("tiny", **SMALL)


def _np(t):
    return t.detach().to(torch.float32).cpu().numpy().copy()


def _record_frame(rec, tag, st, frame_idx, masks):
    """Per yielded frame: the video-res logits (as the caller sees them) and what the state keeps."""
    rec[f"{tag}.f{frame_idx}.video_res_masks"] = _np(masks)
    od = st["output_dict"]
    out = od["cond_frame_outputs"].get(frame_idx) or od["non_cond_frame_outputs"].get(frame_idx)
    if out is None:
        return
    rec[f"{tag}.f{frame_idx}.pred_masks"] = _np(out["pred_masks"])
    rec[f"{tag}.f{frame_idx}.obj_ptr"] = _np(out["obj_ptr"])
    rec[f"{tag}.f{frame_idx}.object_score_logits"] = _np(out["object_score_logits"])
    if out.get("maskmem_features") is not None:
        # every 4th channel: keeps the committed fixtures small, still covers all pixels / objects
        rec[f"{tag}.f{frame_idx}.maskmem_features"] = _np(out["maskmem_features"][:, ::4])


def run_offline(predictor, num_frames=4, size=512, num_objects=1, seed=3):
    vid = BilliardVideo(num_objects=num_objects, height=size, width=size, num_frames=num_frames, seed=seed)
    frames = list(vid.frames())
    rec = {}
    with torch.inference_mode():
        st = predictor.init_state(frames)
        for oid, box in vid.boxes(0).items():
            f, ids, m = predictor.add_new_points_or_box(st, 0, oid, box=np.asarray(box, dtype=np.float32))
            rec[f"prompt.obj{oid}.video_res_masks"] = _np(m)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "track", st, f, m)
        rec["obj_ids"] = np.asarray(list(st["obj_ids"]), dtype=np.int64)
    return rec


def run_fullsize(predictor, model, height, width, num_objects, num_frames, seed):
    """Box prompts on frame 0, forward tracking, at a full-size model.  Compact record (the fixtures are committed):
    low-resolution logits, pointers, scores, every 4th memory channel, and the video-resolution masks as packed bits."""
    vid = BilliardVideo(num_objects=num_objects, height=height, width=width, num_frames=num_frames, seed=seed)
    rec = {}
    with torch.inference_mode():
        st = predictor.init_state(list(vid.frames()))
        for oid, box in vid.boxes(0).items():
            f, ids, m = predictor.add_new_points_or_box(st, 0, oid, box=np.asarray(box, dtype=np.float32))
        rec["prompt.masks_packed"] = np.packbits((_np(m) > 0).astype(np.uint8))
        for f, ids, m in predictor.propagate_in_video(st):
            od = st["output_dict"]
            out = od["cond_frame_outputs"].get(f) or od["non_cond_frame_outputs"][f]
            rec[f"track.f{f}.pred_masks"] = _np(out["pred_masks"])
            rec[f"track.f{f}.obj_ptr"] = _np(out["obj_ptr"])
            rec[f"track.f{f}.object_score_logits"] = _np(out["object_score_logits"])
            rec[f"track.f{f}.maskmem_features"] = _np(out["maskmem_features"][:, ::4])
            rec[f"track.f{f}.masks_packed"] = np.packbits((_np(m) > 0).astype(np.uint8))
        rec["obj_ids"] = np.asarray(list(st["obj_ids"]), dtype=np.int64)
    return rec


def run_stream(predictor, chunk=3, chunks=3, height=192, width=256, seed=5, max_track=5, window=4):
    """Chunked stream: chunk 0 -> objects {0,1}; chunk 1 -> re-prompt + NEW object 2; chunk 2 ->
    re-prompt all; after every chunk reverse-propagate `max_track` frames and release old frames."""
    nobj = 3
    vid = BilliardVideo(num_objects=nobj, height=height, width=width, num_frames=chunk * chunks, seed=seed)
    rec = {}
    st = None
    with torch.inference_mode():
        for c in range(chunks):
            buf = [vid.frame(t) for t in range(c * chunk, (c + 1) * chunk)]
            st = predictor.init_state(buf) if st is None else predictor.update_state(buf, st)
            last = (c + 1) * chunk - 1
            ids = (0, 1) if c == 0 else (0, 1, 2)
            boxes = vid.boxes(last)
            for oid in ids:
                predictor.add_new_points_or_box(st, last, oid, box=np.asarray(boxes[oid], dtype=np.float32))
            for f, oids, m in predictor.propagate_in_video(st, start_frame_idx=last, max_frame_num_to_track=max_track,
                                                           reverse=True):
                _record_frame(rec, f"c{c}", st, f, m)
                rec[f"c{c}.f{f}.obj_ids"] = np.asarray(list(oids), dtype=np.int64)
            predictor.release_old_frames(st, last, window, 0, release_images=True)
            rec[f"c{c}.images_idx"] = np.asarray(st["images_idx"], dtype=np.int64)
            rec[f"c{c}.cond_keys"] = np.asarray(sorted(st["output_dict"]["cond_frame_outputs"]), dtype=np.int64)
            rec[f"c{c}.non_cond_keys"] = np.asarray(sorted(st["output_dict"]["non_cond_frame_outputs"]), dtype=np.int64)
            rec[f"c{c}.num_frames"] = np.asarray([st["num_frames"], len(st["images"])], dtype=np.int64)
    return rec


def run_preload(predictor, pre=3, extra=3, height=192, width=256, seed=7):
    nobj = 2
    vid = BilliardVideo(num_objects=nobj, height=height, width=width, num_frames=pre + extra, seed=seed)
    rec = {}
    with torch.inference_mode():
        # ---- build the bank: every frame a conditioning frame (detect_interval = 1) ----
        st = predictor.init_state([vid.frame(t) for t in range(pre)])
        for t in range(pre):
            for oid, box in vid.boxes(t).items():
                predictor.add_new_points_or_box(st, t, oid, box=np.asarray(box, dtype=np.float32))
        for f, oids, m in predictor.propagate_in_video(st, start_frame_idx=pre - 1, max_frame_num_to_track=pre,
                                                       reverse=True):
            _record_frame(rec, "bank", st, f, m)
        # ---- pickle round trip (det_sam2_RT.py:489-503) ----
        blob = io.BytesIO()
        pickle.dump(st, blob)
        rec["bank.cond_keys"] = np.asarray(sorted(st["output_dict"]["cond_frame_outputs"]), dtype=np.int64)
        st = pickle.loads(blob.getvalue())
        # det_sam2_RT.py:127-133: mark the preload frames
        st["preloading_memory_cond_frame_idx"] = list(st["output_dict"]["cond_frame_outputs"].keys())
        st["preloading_memory_non_cond_frames_idx"] = list(st["output_dict"]["non_cond_frame_outputs"].keys())
        predictor.init_preloading_state(st, offload_video_to_cpu=True, offload_state_to_cpu=False)
        # ---- stream new frames against the bank, no new prompts (detect_interval = -1) ----
        st = predictor.update_state([vid.frame(t) for t in range(pre, pre + extra)], st)
        last = pre + extra - 1
        # forward pass over the appended frames starting from the last preload frame
        for f, oids, m in predictor.propagate_in_video(st, start_frame_idx=pre - 1, max_frame_num_to_track=extra,
                                                       reverse=False):
            _record_frame(rec, "live", st, f, m)
        predictor.release_old_frames(st, last, 2, pre, release_images=True)
        rec["live.images_idx"] = np.asarray(st["images_idx"], dtype=np.int64)
        rec["live.cond_keys"] = np.asarray(sorted(st["output_dict"]["cond_frame_outputs"]), dtype=np.int64)
        rec["live.non_cond_keys"] = np.asarray(sorted(st["output_dict"]["non_cond_frame_outputs"]), dtype=np.int64)
    return rec


def run_mask_prompt(predictor, num_frames=3, height=192, width=256, seed=9):
    """Dense mask prompt (svp:527-600 add_new_mask -> sam2_base.py:399-448 _use_mask_as_output): object 0 is
    prompted with its ground-truth disc as a boolean video-resolution mask (antialiased resize to the model
    resolution + 0.5 threshold inside the predictor), object 1 with a box on the same frame, object 2 with an
    EMPTY mask (never appears: pointer = no_obj_ptr, score -10); then forward tracking."""
    vid = BilliardVideo(num_objects=2, height=height, width=width, num_frames=num_frames, seed=seed)
    rec = {}
    with torch.inference_mode():
        st = predictor.init_state([vid.frame(t) for t in range(num_frames)])
        x0, y0, x1, y1 = [float(v) for v in vid.boxes(0)[0]]
        yy, xx = np.mgrid[0:height, 0:width]
        cx, cy, r = (x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0) / 2 - 2
        disc = ((xx - cx) ** 2 + (yy - cy) ** 2) <= r * r
        f, ids, m = predictor.add_new_mask(st, 0, 0, disc)
        rec["prompt.mask.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 1, box=np.asarray(vid.boxes(0)[1], dtype=np.float32))
        rec["prompt.box.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_mask(st, 0, 2, np.zeros((height, width), dtype=bool))
        rec["prompt.empty.video_res_masks"] = _np(m)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "track", st, f, m)
        rec["obj_ids"] = np.asarray(list(st["obj_ids"]), dtype=np.int64)
    return rec


def run_points_api(predictor, num_frames=4, height=192, width=256, seed=23):
    """Point prompts and session management through the public API (svp:344-520, 1061-1170, 1438-1553):
    object 0 = ONE positive click (the multimask path of an initial conditioning frame, sam2_base.py:922-932),
    object 1 = a positive and a negative click, object 2 = a box plus a positive click (box corners become the first
    two points, svp:398-413); forward tracking; clear_all_prompts_in_frame + remove_object of object 1 (the state is
    re-indexed, svp:1438-1553) and tracking again; reset_state, a click for a NEW object id on frame 1, tracking from
    there."""
    vid = BilliardVideo(num_objects=3, height=height, width=width, num_frames=num_frames, seed=seed)
    rec = {}
    pt = lambda xy: np.asarray([xy], dtype=np.float32)  # noqa: E731
    with torch.inference_mode():
        st = predictor.init_state([vid.frame(t) for t in range(num_frames)])
        c = vid.centers(0)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 0, points=pt(c[0]), labels=np.array([1], np.int32))
        rec["prompt.click.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 1, points=np.asarray([c[1], c[1] + 30.0], dtype=np.float32),
                                                    labels=np.array([1, 0], np.int32))
        rec["prompt.posneg.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 2, points=pt(c[2]), labels=np.array([1], np.int32),
                                                    box=np.asarray(vid.boxes(0)[2], dtype=np.float32))
        rec["prompt.boxclick.video_res_masks"] = _np(m)
        rec["prompt.obj_ids"] = np.asarray(list(ids), dtype=np.int64)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "track", st, f, m)
        predictor.clear_all_prompts_in_frame(st, 0, 1)
        ids, _ = predictor.remove_object(st, 1)
        rec["removed.obj_ids"] = np.asarray(list(ids), dtype=np.int64)
        rec["removed.idx_map"] = np.asarray(sorted(st["obj_id_to_idx"].items()), dtype=np.int64)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "removed", st, f, m)
            rec[f"removed.f{f}.obj_ids"] = np.asarray(list(ids), dtype=np.int64)
        predictor.reset_state(st)
        rec["reset.obj_ids"] = np.asarray(list(st["obj_ids"]), dtype=np.int64)
        rec["reset.flags"] = np.asarray([int(st["tracking_has_started"]), len(st["output_dict"]["cond_frame_outputs"]),
                                         len(st["output_dict"]["non_cond_frame_outputs"])], dtype=np.int64)
        f, ids, m = predictor.add_new_points_or_box(st, 1, 5, points=pt(vid.centers(1)[0]), labels=np.array([1], np.int32))
        rec["reset.prompt.video_res_masks"] = _np(m)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "reset", st, f, m)
            rec[f"reset.f{f}.obj_ids"] = np.asarray(list(ids), dtype=np.int64)
    return rec


def run_refine_click(predictor, num_frames=3, height=192, width=256, seed=29):
    """Interactive refinement on a prompted frame (svp:463-480): a second and a third click on the SAME frame with
    clear_old_points=False — the previous low-resolution logits of that frame, clamped to [-32, 32], come back as a dense
    prompt through PromptEncoder.mask_downscaling (sam2_base.py:306-329) and the clicks accumulate; then tracking."""
    vid = BilliardVideo(num_objects=2, height=height, width=width, num_frames=num_frames, seed=seed)
    rec = {}
    with torch.inference_mode():
        st = predictor.init_state([vid.frame(t) for t in range(num_frames)])
        c = vid.centers(0)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 0, points=np.asarray([c[0]], dtype=np.float32),
                                                    labels=np.array([1], np.int32))
        rec["click1.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 0, points=np.asarray([c[0] + 4.0], dtype=np.float32),
                                                    labels=np.array([1], np.int32), clear_old_points=False)
        rec["click2.video_res_masks"] = _np(m)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 0, points=np.asarray([c[0] + 40.0], dtype=np.float32),
                                                    labels=np.array([0], np.int32), clear_old_points=False)
        rec["click3.video_res_masks"] = _np(m)
        pts = st["point_inputs_per_obj"][0][0]
        rec["clicks.point_labels"] = pts["point_labels"].detach().cpu().numpy().astype(np.int64)
        f, ids, m = predictor.add_new_points_or_box(st, 0, 1, box=np.asarray(vid.boxes(0)[1], dtype=np.float32))
        rec["box.video_res_masks"] = _np(m)
        for f, ids, m in predictor.propagate_in_video(st):
            _record_frame(rec, "track", st, f, m)
        rec["obj_ids"] = np.asarray(list(st["obj_ids"]), dtype=np.int64)
    return rec


def run_video_processor(make_vp, num_frames=11, height=160, width=224, seed=11, repeat_class=None):
    """Det-SAM2's own driver (det_sam2_RT.py VideoProcessor.run) over a frame folder: K = 4 frames per
    chunk, detection every 4 frames, reverse window M = 6, state window S = 6 with image release, a
    third object that the detector only reports from frame 4 on (online new id), and a 3-frame tail.
    `make_vp(detector, **ctor_kwargs)` builds the reference's or this repo's VideoProcessor.
    `repeat_class` ({obj_id: (dx, dy)}) makes the detector report a second, shifted box of that class on every
    detection frame: the reference then prompts that obj_id twice on one frame and the second call receives the first
    call's mask logits as a dense prompt (det_sam2_RT.py:288-302 -> svp:470-482)."""
    import os
    import tempfile

    import cv2
    from detsam2_b200.synthetic import GroundTruthDetector
    vid = BilliardVideo(num_objects=3, height=height, width=width, num_frames=num_frames, seed=seed)
    det = GroundTruthDetector(vid, detect_interval=4, appear_at={2: 4}, repeat_class=repeat_class)
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        fdir = os.path.join(tmp, "frames")
        os.makedirs(fdir)
        for t in range(num_frames):
            cv2.imwrite(os.path.join(fdir, f"{t:05d}.png"), cv2.cvtColor(vid.frame(t), cv2.COLOR_RGB2BGR))
        vp = make_vp(det, output_dir=os.path.join(tmp, "out"), frame_buffer_size=4, detect_interval=4,
                     max_frame_num_to_track=6, max_inference_state_frames=6, skip_classes={11, 14, 15, 19})
        with torch.inference_mode():
            vp.run(frame_dir=fdir, output_video_segments_pkl_path=os.path.join(tmp, "seg.pkl"),
                   output_special_classes_detection_pkl_path=os.path.join(tmp, "special.pkl"))
        with open(os.path.join(tmp, "seg.pkl"), "rb") as f:
            segs = pickle.load(f)
    rec["frames"] = np.asarray(sorted(segs), dtype=np.int64)
    for t in sorted(segs):
        ids = sorted(segs[t])
        rec[f"f{t}.obj_ids"] = np.asarray(ids, dtype=np.int64)
        rec[f"f{t}.masks_packed"] = np.packbits(np.stack([segs[t][i] for i in ids]).astype(np.uint8))
    st = vp.inference_state
    rec["images_idx"] = np.asarray(st["images_idx"], dtype=np.int64)
    rec["cond_keys"] = np.asarray(sorted(st["output_dict"]["cond_frame_outputs"]), dtype=np.int64)
    rec["non_cond_keys"] = np.asarray(sorted(st["output_dict"]["non_cond_frame_outputs"]), dtype=np.int64)
    return rec


def run_image_predictor(ip, height=480, width=640, seed=41):
    """SAM2ImagePredictor (sam2_image_predictor.py): set_image on a non-square image (antialiased resize), then
    (1) one click, three candidate masks; (2) a box, single mask with the stability fallback; (3) box + click +
    the previous low-resolution logits as a dense prompt; (4) two boxes in one batched call (repeat_image)."""
    vid = BilliardVideo(num_objects=3, height=height, width=width, num_frames=1, seed=seed)
    img = vid.frame(0)
    c, boxes = vid.centers(0), vid.boxes(0)
    rec = {}

    def put(tag, out):
        masks, ious, low = out
        rec[f"{tag}.masks_packed"] = np.packbits((np.asarray(masks) > 0).astype(np.uint8))
        rec[f"{tag}.mask_logits"] = np.asarray(masks, dtype=np.float32)[..., ::4, ::4].copy()
        rec[f"{tag}.ious"] = np.asarray(ious, dtype=np.float32)
        rec[f"{tag}.pred_masks"] = np.asarray(low, dtype=np.float32)

    with torch.inference_mode():
        ip.set_image(img)
        rec["embedding_shape"] = np.asarray(tuple(ip.get_image_embedding().shape), dtype=np.int64)
        out = ip.predict(point_coords=np.asarray([c[0]], np.float32), point_labels=np.array([1], np.int32),
                         multimask_output=True, return_logits=True)
        put("click", out)
        out2 = ip.predict(box=np.asarray(boxes[1], np.float32), multimask_output=False, return_logits=True)
        put("box", out2)
        # refinement: the same box plus a click, with the box step's low-resolution logits as the dense prompt (the click
        # step's candidates are not used here: with random weights they are chaotic — the reference's own bf16 run
        # deviates by 0.3 on them — and would hand different dense prompts to the implementations under comparison)
        out3 = ip.predict(point_coords=np.asarray([c[1]], np.float32), point_labels=np.array([1], np.int32),
                          box=np.asarray(boxes[1], np.float32), mask_input=out2[2][0][None], multimask_output=False,
                          return_logits=True)
        put("refine", out3)
        out4 = ip.predict(box=np.asarray([boxes[1], boxes[2]], np.float32), multimask_output=False, return_logits=True)
        put("boxes2", out4)
        m = ip.predict(box=np.asarray(boxes[2], np.float32), multimask_output=False)[0]
        # thresholded output: the reference hands back float32 0.0 / 1.0, not booleans (sam2_image_predictor.py:299)
        rec["plain.is_float32_01"] = np.asarray([int(m.dtype == np.float32 and set(np.unique(m)) <= {0.0, 1.0})], dtype=np.int64)
        rec["plain.shape"] = np.asarray(m.shape, dtype=np.int64)
    return rec


def image_predictor_config():
    # the reference hard-codes the backbone feature sizes of a 1024^2 input (sam2_image_predictor.py:60-64)
    return get_config("tiny")


def packed_mask_iou(a, b):
    """IoU of two np.packbits arrays."""
    a, b = np.unpackbits(a), np.unpackbits(b)
    u = np.logical_or(a, b).sum()
    return 1.0 if u == 0 else float(np.logical_and(a, b).sum()) / float(u)


def load_golden(name):
    """tests/golden/<name>.npz -> (arrays, weights fingerprint)."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"{name}.npz")
    assert os.path.exists(path), f"{path} missing: run `python -m oracle.gen_golden` in the build container"
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d.pop("__torch_version", None)
    return d, d.pop("__weights_fingerprint")


# VideoProcessor.run scenarios: fixture name -> keyword arguments of run_video_processor.  `video_processor_dup`: the
# detector reports object 1 twice per detection frame (second box shifted by 6 px)
VP_SCENARIOS = {"video_processor": {}, "video_processor_dup": {"repeat_class": {1: (6.0, -4.0)}}}

SCENARIOS = {**{_n: (lambda predictor, _kw=_kw: run_fullsize(predictor, **_kw)) for _n, _kw in FULLSIZE.items()},
             "offline": run_offline, "stream": run_stream, "preload": run_preload, "mask_prompt": run_mask_prompt,
             "points_api": run_points_api, "refine_click": run_refine_click}


def compare(got, ref, rtol_rms, iou_min=None, int_exact=True):
    """Returns a list of human-readable mismatches (empty = parity).  Float arrays are compared by
    relative RMS error; integer arrays must be identical; `*.video_res_masks` additionally by the
    IoU of the thresholded masks when `iou_min` is given."""
    bad = []
    for k, r in ref.items():
        if k not in got:
            bad.append(f"{k}: missing")
            continue
        g = got[k]
        if g.shape != r.shape:
            bad.append(f"{k}: shape {g.shape} != {r.shape}")
            continue
        if k.endswith("masks_packed"):
            iou = packed_mask_iou(g, r)
            # boolean masks are all the reference's driver hands out: a logit that is 0 to fp32 rounding may land on
            # either side of the `> 0` threshold, so ONE differing pixel per frame is not a mismatch (on the small
            # VideoProcessor frames one pixel of a ~600-pixel union already reads as IoU 0.9984)
            flips = int((np.unpackbits(g) != np.unpackbits(r)).sum())
            if iou < (iou_min if iou_min is not None else 1.0) and flips > 1:
                bad.append(f"{k}: packed-mask IoU {iou:.5f} ({flips} differing pixels)")
            continue
        if np.issubdtype(r.dtype, np.integer):
            if int_exact and not np.array_equal(g, r):
                bad.append(f"{k}: integer mismatch {g.tolist()} != {r.tolist()}")
            continue
        g64, r64 = g.astype(np.float64), r.astype(np.float64)
        den = max(np.sqrt(np.mean(r64 ** 2)), 1e-12)
        err = np.sqrt(np.mean((g64 - r64) ** 2)) / den
        # fixtures stored as fp16 carry their own quantisation (2^-11 relative, ~3e-4 rms)
        tol = rtol_rms + (6e-4 if r.dtype == np.float16 else 0.0)
        if not np.isfinite(err) or err > tol:
            bad.append(f"{k}: rel-rms {err:.3e} > {tol:.1e}")
        if iou_min is not None and k.endswith("video_res_masks"):
            a, b = g > 0, r > 0
            u = np.logical_or(a, b).sum()
            iou = 1.0 if u == 0 else np.logical_and(a, b).sum() / u
            if iou < iou_min:
                bad.append(f"{k}: mask IoU {iou:.5f} < {iou_min}")
    return bad
