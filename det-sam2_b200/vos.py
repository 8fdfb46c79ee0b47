"""Semi-supervised VOS driver over the drop-in predictor: the per-object mask-prompt path of the reference's
``tools/vos_inference.py`` (``vos_inference``, :117-247): JPEG frames ``<video>/<index>.jpg``, object masks as palette
PNGs (one PNG per frame with object ids as pixel values, or one ``<object id:03d>/`` folder per object), every object of
the input frame(s) prompted with ``add_new_mask``, forward ``propagate_in_video``, results written back as palette PNGs.
SURVEY.md 8f rank 4.  Host glue only: all arithmetic is the predictor's engine.
"""
import os

import numpy as np
import torch

# the 256-entry DAVIS colour map (bit-interleaved object ids): what the reference falls back to when the input PNGs
# carry no palette (tools/vos_inference.py:16)
def _davis_palette():
    pal = np.zeros((256, 3), dtype=np.uint8)
    for i in range(256):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        pal[i] = (r, g, b)
    return pal.tobytes()


DAVIS_PALETTE = _davis_palette()


def load_ann_png(path):
    """PNG -> (uint8 id map, palette or None)."""
    from PIL import Image
    im = Image.open(path)
    return np.array(im).astype(np.uint8), im.getpalette()


def save_ann_png(path, mask, palette):
    from PIL import Image
    assert mask.dtype == np.uint8 and mask.ndim == 2
    out = Image.fromarray(mask)
    out.putpalette(palette)
    out.save(path)


def split_objects(id_map):
    """id map -> {object id: boolean mask}; 0 is background."""
    return {int(i): id_map == i for i in np.unique(id_map) if i > 0}


def merge_objects(per_obj_mask, height, width):
    """{object id: boolean mask} -> id map; where masks overlap the SMALLER object id wins (the reference paints ids in
    descending order, tools/vos_inference.py:47-55)."""
    out = np.zeros((height, width), dtype=np.uint8)
    for oid in sorted(per_obj_mask, reverse=True):
        out[np.asarray(per_obj_mask[oid]).reshape(height, width)] = oid
    return out


def load_masks_from_dir(input_mask_dir, video_name, frame_name, per_obj_png_file, allow_missing=False):
    if not per_obj_png_file:
        path = os.path.join(input_mask_dir, video_name, f"{frame_name}.png")
        if allow_missing and not os.path.exists(path):
            return {}, None
        id_map, palette = load_ann_png(path)
        return split_objects(id_map), palette
    per_obj, palette = {}, None
    for object_name in os.listdir(os.path.join(input_mask_dir, video_name)):
        path = os.path.join(input_mask_dir, video_name, object_name, f"{frame_name}.png")
        if allow_missing and not os.path.exists(path):
            continue
        m, palette = load_ann_png(path)
        per_obj[int(object_name)] = m > 0
    return per_obj, palette


def save_masks_to_dir(output_mask_dir, video_name, frame_name, per_obj_output_mask, height, width, per_obj_png_file,
                      output_palette):
    os.makedirs(os.path.join(output_mask_dir, video_name), exist_ok=True)
    if not per_obj_png_file:
        save_ann_png(os.path.join(output_mask_dir, video_name, f"{frame_name}.png"),
                     merge_objects(per_obj_output_mask, height, width), output_palette)
        return
    for oid, m in per_obj_output_mask.items():
        d = os.path.join(output_mask_dir, video_name, f"{oid:03d}")
        os.makedirs(d, exist_ok=True)
        save_ann_png(os.path.join(d, f"{frame_name}.png"), np.asarray(m).reshape(height, width).astype(np.uint8),
                     output_palette)


@torch.inference_mode()
def vos_inference(predictor, base_video_dir, input_mask_dir, output_mask_dir, video_name, score_thresh=0.0,
                  use_all_masks=False, per_obj_png_file=False):
    """tools/vos_inference.py:117-247 for one video.  Returns {frame index: {object id: bool [1,H,W]}} (also written as
    PNGs).  Same RuntimeErrors as the reference for missing masks / object ids that first appear on a later frame."""
    video_dir = os.path.join(base_video_dir, video_name)
    frame_names = [os.path.splitext(p)[0] for p in os.listdir(video_dir)
                   if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG")]
    frame_names.sort(key=lambda p: int(os.path.splitext(p)[0]))
    state = predictor.init_state(video_path=video_dir, async_loading_frames=False)
    height, width = state["video_height"], state["video_width"]
    palette = None
    if not use_all_masks:
        input_frame_inds = [0]
    else:
        root = os.path.join(input_mask_dir, video_name)
        if not per_obj_png_file:
            input_frame_inds = [i for i, n in enumerate(frame_names) if os.path.exists(os.path.join(root, f"{n}.png"))]
        else:
            input_frame_inds = [i for o in os.listdir(root) for i, n in enumerate(frame_names)
                                if os.path.exists(os.path.join(root, o, f"{n}.png"))]
        if not input_frame_inds:
            raise RuntimeError(f"In {video_name=}, got no input masks in {input_mask_dir=}. "
                               "Please make sure the input masks are available in the correct format.")
        input_frame_inds = sorted(set(input_frame_inds))
    object_ids = None
    for idx in input_frame_inds:
        try:
            per_obj, palette = load_masks_from_dir(input_mask_dir, video_name, frame_names[idx], per_obj_png_file)
        except FileNotFoundError as e:
            raise RuntimeError(f"In {video_name=}, failed to load input mask for frame input_frame_idx={idx}.") from e
        if object_ids is None:
            object_ids = set(per_obj)
        for oid, m in per_obj.items():
            if oid not in object_ids:
                raise RuntimeError(f"In {video_name=}, got a new object_id={oid} appearing only in a later "
                                   f"input_frame_idx={idx} (but not appearing in the first frame).")
            predictor.add_new_mask(inference_state=state, frame_idx=idx, obj_id=oid, mask=m)
    if not object_ids:
        raise RuntimeError(f"In {video_name=}, got no object ids on {input_frame_inds=}.")
    os.makedirs(os.path.join(output_mask_dir, video_name), exist_ok=True)
    out_palette = palette or DAVIS_PALETTE
    segments = {}
    for f, ids, logits in predictor.propagate_in_video(state):
        m = (logits > score_thresh).cpu().numpy()        # one threshold + one copy for all objects
        segments[f] = {oid: m[i] for i, oid in enumerate(ids)}
    for f, per_obj in segments.items():
        save_masks_to_dir(output_mask_dir, video_name, frame_names[f], per_obj, height, width, per_obj_png_file, out_palette)
    return segments
