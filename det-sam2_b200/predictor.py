"""Host-side mirror of the reference's ``SAM2VideoPredictor`` (/root/reference/sam2/
sam2_video_predictor.py) — same method names, arguments, return values, error behaviour and the same
``inference_state`` dictionary schema (which is also the pickle format of Det-SAM2's "preload memory
bank", det_sam2_RT.py:489-503) — over an *engine* that owns all arithmetic.

The product engine is ``detsam2_b200.engine.CudaEngine`` (hand-written sm_100a kernels behind the C
ABI of include/detsam2.h).  The predictor itself only does bookkeeping: which frames are
conditioning frames, which memories a step may read, per-object views, the constant-memory window.
It never computes on the CPU on behalf of a missing CUDA library.

Engine seams (each mirrors one reference function; see engine.py):
  encode_image(image_f16)                               svp:1174-1212 / sam2_base.py:450-477
  condition_on_memory(feats, B, frame_idx, is_init, output_dict, num_frames, reverse, preload)
                                                        sam2_base.py:479-690
  sam_heads(pix_feat, feats, B, coords, labels, mask_inputs, multimask) / mask_as_output(...)
                                                        sam2_base.py:254-448
  encode_memory(feats, B, low_res_masks, scores, is_mask_from_pts)   sam2_base.py:692-743
  fill_holes(pred_masks, max_area), resize_masks(masks, H, W)        svp:1341-1348, 618-642
"""
import gc
import os
from collections import OrderedDict

import torch

from .frames import load_video_frames

NO_OBJ_SCORE = -1024.0  # sam2_base.py:17


def concat_points(old_point_inputs, new_points, new_labels):
    """misc.py:396-405."""
    if old_point_inputs is None:
        points, labels = new_points, new_labels
    else:
        points = torch.cat([old_point_inputs["point_coords"], new_points], dim=1)
        labels = torch.cat([old_point_inputs["point_labels"], new_labels], dim=1)
    return {"point_coords": points, "point_labels": labels}


def _empty_frame_dict():
    return {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}


class SAM2VideoPredictor:
    """Drop-in for the reference class of the same name (svp:20-41).  ``engine`` supplies compute."""

    def __init__(self, engine, fill_hole_area=0, non_overlap_masks=False, clear_non_cond_mem_around_input=False,
                 clear_non_cond_mem_for_multi_obj=False, add_all_frames_to_correct_as_cond=False,
                 feature_cache_frames=1, encoder_batch_frames=None, encoder_overlap=None, verbose=False):
        self.engine = engine
        self.cfg = engine.cfg
        self.fill_hole_area = fill_hole_area
        self.non_overlap_masks = non_overlap_masks
        self.clear_non_cond_mem_around_input = clear_non_cond_mem_around_input
        self.clear_non_cond_mem_for_multi_obj = clear_non_cond_mem_for_multi_obj
        self.add_all_frames_to_correct_as_cond = add_all_frames_to_correct_as_cond
        # the reference caches exactly one frame's backbone features (svp:1190); a larger cache is a
        # pure speed-up for Det-SAM2's reverse re-tracking (each frame is visited M/K times)
        self.feature_cache_frames = max(1, int(feature_cache_frames))
        # frames encoded per pass of the image encoder while propagating: the frames still to be tracked are known
        # (processing order of propagate_in_video) and their backbone features do not depend on the tracker, so the
        # next few are encoded together (engine.encode_images; bit-identical per frame, a better shape for the GPU).
        # 1 = the reference's frame-at-a-time behaviour; DS2_ENC_BATCH overrides the default of 4.
        if encoder_batch_frames is None:
            encoder_batch_frames = int(os.environ.get("DS2_ENC_BATCH", "4"))
        self.encoder_batch_frames = max(1, int(encoder_batch_frames)) if hasattr(engine, "encode_images") else 1
        # while the tracker works through the frames of one encoder pass, the pass over the NEXT frames of the processing
        # order runs on the engine's encoder stream (engine.encode_images_async) and fills the SMs the tracker's small
        # kernels and partial waves leave idle.  Same kernels and bits; DS2_ENC_OVERLAP=0 (or encoder_overlap=False)
        # keeps every pass on the caller's stream.
        if encoder_overlap is None:
            encoder_overlap = os.environ.get("DS2_ENC_OVERLAP", "1") != "0"
        self.encoder_overlap = bool(encoder_overlap) and hasattr(engine, "encode_images_async")
        self._upcoming = None      # (id(state), frames the running propagate call will still encode, in order)
        self._prefetched = {}      # frame_idx -> features encoded ahead of their step (same propagate call only)
        self._pending = None       # (frames, PendingFeats): the pass in flight on the encoder stream, at most one
        self.verbose = verbose

    @classmethod
    def from_pretrained(cls, model_id, **kwargs):
        """svp:208-222: build from a Hugging Face model id (``facebook/sam2.1-hiera-{tiny,small,base-plus,large}``)."""
        from .build_sam import build_sam2_video_predictor_hf
        return build_sam2_video_predictor_hf(model_id, **kwargs)

    # ---- attributes callers read on the reference model ------------------------------------------
    @property
    def device(self):
        return self.engine.device

    @property
    def image_size(self):
        return self.cfg.image_size

    @property
    def hidden_dim(self):
        return self.cfg.hidden_dim

    @property
    def num_maskmem(self):
        return self.cfg.num_maskmem

    def _log(self, msg):
        if self.verbose:
            print(msg)

    # ---- session state ---------------------------------------------------------------------------
    @torch.inference_mode()
    def init_state(self, video_path, offload_video_to_cpu=True, offload_state_to_cpu=False,
                   async_loading_frames=False):
        """svp:44-120."""
        compute_device = self.device
        images, video_height, video_width = load_video_frames(
            video_path, self.image_size, offload_video_to_cpu, compute_device,
            async_loading_frames=async_loading_frames, pin=compute_device.type == "cuda")
        st = {}
        st["images"] = images
        st["num_frames"] = len(images)
        st["images_idx"] = list(range(len(images)))
        st["offload_video_to_cpu"] = offload_video_to_cpu
        st["offload_state_to_cpu"] = offload_state_to_cpu
        st["video_height"] = video_height
        st["video_width"] = video_width
        st["device"] = compute_device
        st["storage_device"] = torch.device("cpu") if offload_state_to_cpu else compute_device
        st["point_inputs_per_obj"] = {}
        st["mask_inputs_per_obj"] = {}
        st["cached_features"] = {}
        st["constants"] = {}
        st["obj_id_to_idx"] = OrderedDict()
        st["obj_idx_to_id"] = OrderedDict()
        st["obj_ids"] = []
        st["output_dict"] = _empty_frame_dict()
        st["output_dict_per_obj"] = {}
        st["temp_output_dict_per_obj"] = {}
        st["consolidated_frame_inds"] = {"cond_frame_outputs": set(), "non_cond_frame_outputs": set()}
        st["tracking_has_started"] = False
        st["frames_already_tracked"] = {}
        st["preloading_memory_cond_frame_idx"] = None
        st["preloading_memory_non_cond_frames_idx"] = None
        st["max_update_length_for_new_obj_id"] = 100
        self._get_image_feature(st, frame_idx=0)  # warm-up, caches frame 0 (svp:119)
        return st

    def init_preloading_state(self, inference_state, offload_video_to_cpu=True, offload_state_to_cpu=True):
        """svp:123-156: re-home a loaded preload bank to this session's storage settings.  As in the
        reference the last preload frame is skipped and every preload frame must be a cond frame."""
        st = inference_state
        if offload_video_to_cpu:
            st["images"] = st["images"].to("cpu")
        elif self.device.type == "cuda":
            # addition: a bank saved from a host-frames session continues with its frames in HBM
            st["images"] = st["images"].to(self.device)
            st["offload_video_to_cpu"] = False
        st["device"] = self.device
        st["storage_device"] = torch.device("cpu") if offload_state_to_cpu else self.device
        dev = st["storage_device"]
        st["cached_features"] = {}
        for frame_idx in range(st["num_frames"] - 1):
            cur = st["output_dict"]["cond_frame_outputs"][frame_idx]
            cur["maskmem_features"] = cur["maskmem_features"].to(dev)
            cur["pred_masks"] = cur["pred_masks"].to(dev)
            for obj_idx in st["obj_idx_to_id"].keys():
                o = st["output_dict_per_obj"][obj_idx]["cond_frame_outputs"][frame_idx]
                o["maskmem_features"] = o["maskmem_features"].to(dev)
                o["pred_masks"] = o["pred_masks"].to(dev)

    @torch.inference_mode()
    def update_state(self, video_path, inference_state, async_loading_frames=False):
        """svp:160-205: append frames to a running session."""
        st = inference_state
        new_images, nh, nw = load_video_frames(
            video_path, self.image_size, st["offload_video_to_cpu"], self.device,
            async_loading_frames=async_loading_frames, pin=False)
        assert st["video_height"] == nh and st["video_width"] == nw, "new frames must match the video size"
        last = st["images_idx"][-1]
        st["images_idx"].extend(range(last + 1, last + 1 + len(new_images)))
        images = st["images"]
        assert images.shape[1:] == new_images.shape[1:]
        # (re-packing host-resident frames into pinned memory was measured and dropped: 21.7 against 28.0 video
        # frames/s in stream mode, profiles/r1_s16_stream_modes.txt; frames in HBM — offload_video_to_cpu=False — is
        # the fast configuration)
        st["images"] = torch.cat((images, new_images.to(images.device)), dim=0)
        st["num_frames"] += len(new_images)
        return st

    # ---- object ids ------------------------------------------------------------------------------
    def _obj_id_to_idx(self, st, obj_id):
        """svp:224-327 incl. Det-SAM2's online new-ID path (re-consolidate recent + preload cond frames)."""
        idx = st["obj_id_to_idx"].get(obj_id, None)
        if idx is not None:
            return idx
        started = st["tracking_has_started"]
        idx = len(st["obj_id_to_idx"])
        st["obj_id_to_idx"][obj_id] = idx
        st["obj_idx_to_id"][idx] = obj_id
        st["obj_ids"] = list(st["obj_id_to_idx"])
        st["point_inputs_per_obj"][idx] = {}
        st["mask_inputs_per_obj"][idx] = {}
        st["output_dict_per_obj"][idx] = _empty_frame_dict()
        st["temp_output_dict_per_obj"][idx] = _empty_frame_dict()
        if started:
            output_dict = st["output_dict"]
            cond_idx = sorted(output_dict["cond_frame_outputs"].keys())
            mx = st["max_update_length_for_new_obj_id"]
            if mx > 0:
                cond_idx = cond_idx[-mx:]
            pre = st["preloading_memory_cond_frame_idx"]
            if pre is not None:
                for t in pre:
                    if t not in cond_idx:
                        cond_idx.append(t)
            self._log(f"new object id {obj_id} while tracking: re-consolidating {len(cond_idx)} cond frames")
            for t in cond_idx:
                out = self._consolidate_temp_output_across_obj(st, t, is_cond=True, run_mem_encoder=True,
                                                               consolidate_at_video_res=False)
                output_dict["cond_frame_outputs"][t] = out
                self._add_output_per_object(st, t, out, "cond_frame_outputs")
        return idx

    def _obj_idx_to_id(self, st, obj_idx):
        return st["obj_idx_to_id"][obj_idx]

    def _get_obj_num(self, st):
        return len(st["obj_idx_to_id"])

    # ---- prompts ---------------------------------------------------------------------------------
    @torch.inference_mode()
    def add_new_points_or_box(self, inference_state, frame_idx, obj_id, points=None, labels=None,
                              clear_old_points=True, normalize_coords=True, box=None):
        """svp:344-520."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        point_inputs_per_frame = st["point_inputs_per_obj"][obj_idx]
        mask_inputs_per_frame = st["mask_inputs_per_obj"][obj_idx]
        if (points is not None) != (labels is not None):
            raise ValueError("points and labels must be provided together")
        if points is None and box is None:
            raise ValueError("at least one of points or box must be provided as input")
        if points is None:
            points = torch.zeros(0, 2, dtype=torch.float32)
        elif not isinstance(points, torch.Tensor):
            points = torch.tensor(points, dtype=torch.float32)
        if labels is None:
            labels = torch.zeros(0, dtype=torch.int32)
        elif not isinstance(labels, torch.Tensor):
            labels = torch.tensor(labels, dtype=torch.int32)
        if points.dim() == 2:
            points = points.unsqueeze(0)
        if labels.dim() == 1:
            labels = labels.unsqueeze(0)
        if box is not None:
            if not clear_old_points:
                raise ValueError("cannot add box without clearing old points, since box prompt must be provided "
                                 "before any point prompt (please use clear_old_points=True instead)")
            if not isinstance(box, torch.Tensor):
                box = torch.tensor(box, dtype=torch.float32, device=points.device)
            box_coords = box.reshape(1, 2, 2)
            box_labels = torch.tensor([2, 3], dtype=torch.int32, device=labels.device).reshape(1, 2)
            points = torch.cat([box_coords, points], dim=1)
            labels = torch.cat([box_labels, labels], dim=1)
        if normalize_coords:
            points = points / torch.tensor([st["video_width"], st["video_height"]]).to(points.device)
        points = points * self.image_size
        points = points.to(st["device"])
        labels = labels.to(st["device"])
        point_inputs = None if clear_old_points else point_inputs_per_frame.get(frame_idx, None)
        point_inputs = concat_points(point_inputs, points, labels)
        point_inputs_per_frame[frame_idx] = point_inputs
        mask_inputs_per_frame.pop(frame_idx, None)

        is_init_cond_frame = frame_idx not in st["frames_already_tracked"]
        reverse = False if is_init_cond_frame else st["frames_already_tracked"][frame_idx]["reverse"]
        obj_output_dict = st["output_dict_per_obj"][obj_idx]
        obj_temp_output_dict = st["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"

        prev_sam_mask_logits = None
        prev_out = obj_temp_output_dict[storage_key].get(frame_idx)
        if prev_out is None:
            prev_out = obj_output_dict["cond_frame_outputs"].get(frame_idx)
            if prev_out is None:
                prev_out = obj_output_dict["non_cond_frame_outputs"].get(frame_idx)
        if prev_out is not None and prev_out["pred_masks"] is not None:
            prev_sam_mask_logits = torch.clamp(prev_out["pred_masks"].to(st["device"]), -32.0, 32.0)

        current_out, _ = self._run_single_frame_inference(
            st, obj_output_dict, frame_idx, batch_size=1, is_init_cond_frame=is_init_cond_frame,
            point_inputs=point_inputs, mask_inputs=None, reverse=reverse, run_mem_encoder=False,
            prev_sam_mask_logits=prev_sam_mask_logits)
        obj_temp_output_dict[storage_key][frame_idx] = current_out

        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond,
                                                                run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    def add_new_points(self, *args, **kwargs):
        return self.add_new_points_or_box(*args, **kwargs)

    @torch.inference_mode()
    def add_new_boxes(self, inference_state, frame_idx, boxes, normalize_coords=True):
        """Addition: box prompts for several objects of ONE frame in one B-wide decoder
        call.  Equivalent to calling ``add_new_points_or_box(frame_idx, obj_id, box=...)`` for the items of
        ``boxes`` ({obj_id: xyxy}) in order — the reference's driver does exactly that, one B = 1 decode and one
        consolidation per detection (det_sam2_RT.py:285-316, svp:485-498) — and returns what the last of those
        calls would return.  The batched path applies when the frame has not been tracked yet and no object has an
        earlier prompt result on it (the detector case); otherwise it falls back to the sequential calls."""
        st = inference_state
        items = list(boxes.items())
        if not items:
            raise ValueError("boxes must hold at least one {obj_id: box} item")
        obj_idxs = [self._obj_id_to_idx(st, oid) for oid, _ in items]   # same registration order / side effects
        batched = frame_idx not in st["frames_already_tracked"] and len(set(obj_idxs)) == len(obj_idxs) and len(items) > 1
        if batched:
            for oi in obj_idxs:
                for d in (st["temp_output_dict_per_obj"][oi], st["output_dict_per_obj"][oi]):
                    if frame_idx in d["cond_frame_outputs"] or frame_idx in d["non_cond_frame_outputs"]:
                        batched = False
        if not batched:
            out = None
            for oid, box in items:
                out = self.add_new_points_or_box(st, frame_idx, oid, box=box, normalize_coords=normalize_coords)
            return out
        scale = torch.tensor([st["video_width"], st["video_height"]], dtype=torch.float32)
        pts, lbs = [], []
        for (oid, box), oi in zip(items, obj_idxs):
            b = box if isinstance(box, torch.Tensor) else torch.tensor(box, dtype=torch.float32)
            p = b.to(torch.float32).reshape(1, 2, 2)
            if normalize_coords:
                p = p / scale
            p = (p * self.image_size).to(st["device"])
            l = torch.tensor([2, 3], dtype=torch.int32).reshape(1, 2).to(st["device"])
            st["point_inputs_per_obj"][oi][frame_idx] = {"point_coords": p, "point_labels": l}
            st["mask_inputs_per_obj"][oi].pop(frame_idx, None)
            pts.append(p)
            lbs.append(l)
        point_inputs = {"point_coords": torch.cat(pts, 0), "point_labels": torch.cat(lbs, 0)}
        # an un-tracked frame is an initial conditioning frame: no memory is read, so the per-object output dict that
        # _run_single_frame_inference would consult for memory is irrelevant and one batched call is exact
        current_out, _ = self._run_single_frame_inference(
            st, st["output_dict_per_obj"][obj_idxs[0]], frame_idx, batch_size=len(items), is_init_cond_frame=True,
            point_inputs=point_inputs, mask_inputs=None, reverse=False, run_mem_encoder=False, prev_sam_mask_logits=None)
        for i, oi in enumerate(obj_idxs):
            sl = slice(i, i + 1)
            st["temp_output_dict_per_obj"][oi]["cond_frame_outputs"][frame_idx] = {
                "maskmem_features": None, "maskmem_pos_enc": None, "pred_masks": current_out["pred_masks"][sl],
                "obj_ptr": current_out["obj_ptr"][sl], "object_score_logits": current_out["object_score_logits"][sl]}
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=True, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    @torch.inference_mode()
    def add_new_mask(self, inference_state, frame_idx, obj_id, mask):
        """svp:527-600."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        if not isinstance(mask, torch.Tensor):
            mask = torch.tensor(mask, dtype=torch.bool)
        assert mask.dim() == 2
        mh, mw = mask.shape
        m = mask[None, None].float().to(st["device"])
        if mh != self.image_size or mw != self.image_size:
            m = torch.nn.functional.interpolate(m, size=(self.image_size, self.image_size), align_corners=False,
                                                mode="bilinear", antialias=True)
            m = (m >= 0.5).float()
        st["mask_inputs_per_obj"][obj_idx][frame_idx] = m
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        is_init_cond_frame = frame_idx not in st["frames_already_tracked"]
        reverse = False if is_init_cond_frame else st["frames_already_tracked"][frame_idx]["reverse"]
        is_cond = is_init_cond_frame or self.add_all_frames_to_correct_as_cond
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        current_out, _ = self._run_single_frame_inference(
            st, st["output_dict_per_obj"][obj_idx], frame_idx, batch_size=1, is_init_cond_frame=is_init_cond_frame,
            point_inputs=None, mask_inputs=m, reverse=reverse, run_mem_encoder=False)
        st["temp_output_dict_per_obj"][obj_idx][storage_key][frame_idx] = current_out
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond,
                                                                run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    # ---- output shaping --------------------------------------------------------------------------
    def _apply_non_overlapping_constraints(self, pred_masks):
        """sam2_base.py:934-952."""
        bs = pred_masks.size(0)
        if bs == 1:
            return pred_masks
        max_obj = torch.argmax(pred_masks, dim=0, keepdim=True)
        keep = max_obj == torch.arange(bs, device=pred_masks.device)[:, None, None, None]
        return torch.where(keep, pred_masks, torch.clamp(pred_masks, max=-10.0))

    def _get_orig_video_res_output(self, st, any_res_masks):
        """svp:618-642."""
        any_res_masks = any_res_masks.to(st["device"])
        video_res_masks = self.engine.resize_masks(any_res_masks, st["video_height"], st["video_width"])
        if self.non_overlap_masks:
            video_res_masks = self._apply_non_overlapping_constraints(video_res_masks)
        sd = st["storage_device"]
        return any_res_masks.to(sd), video_res_masks.to(sd)

    def _consolidate_temp_output_across_obj(self, st, frame_idx, is_cond, run_mem_encoder,
                                            consolidate_at_video_res=False):
        """svp:644-767."""
        B = self._get_obj_num(st)
        storage_key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if consolidate_at_video_res:
            assert not run_mem_encoder, "memory encoder cannot run at video resolution"
            H, W, mask_key = st["video_height"], st["video_width"], "pred_masks_video_res"
        else:
            H = W = self.image_size // 4
            mask_key = "pred_masks"
        out = {
            "maskmem_features": None,
            "maskmem_pos_enc": None,
            mask_key: torch.full((B, 1, H, W), NO_OBJ_SCORE, dtype=torch.float32, device=st["storage_device"]),
            "obj_ptr": torch.full((B, self.hidden_dim), NO_OBJ_SCORE, dtype=torch.float32, device=st["device"]),
            "object_score_logits": torch.full((B, 1), 10.0, dtype=torch.float32, device=st["device"]),
        }
        empty_mask_ptr = None
        for obj_idx in range(B):
            tmp = st["temp_output_dict_per_obj"][obj_idx]
            per = st["output_dict_per_obj"][obj_idx]
            o = tmp[storage_key].get(frame_idx, None)
            if o is None:
                o = per["cond_frame_outputs"].get(frame_idx, None)
            if o is None:
                o = per["non_cond_frame_outputs"].get(frame_idx, None)
            if o is None:
                if run_mem_encoder:
                    if empty_mask_ptr is None:
                        empty_mask_ptr = self._get_empty_mask_ptr(st, frame_idx)
                    out["obj_ptr"][obj_idx:obj_idx + 1] = empty_mask_ptr
                continue
            obj_mask = o["pred_masks"]
            dst = out[mask_key]
            if obj_mask.shape[-2:] == dst.shape[-2:]:
                dst[obj_idx:obj_idx + 1] = obj_mask
            else:
                r = self.engine.resize_masks(obj_mask.to(st["device"]), dst.shape[-2], dst.shape[-1])
                dst[obj_idx:obj_idx + 1] = r.to(dst.device)
            out["obj_ptr"][obj_idx:obj_idx + 1] = o["obj_ptr"]
            out["object_score_logits"][obj_idx:obj_idx + 1] = o["object_score_logits"]
        if run_mem_encoder:
            masks = out["pred_masks"].to(st["device"])
            if getattr(self.cfg, "non_overlap_masks_for_mem_enc", False):
                raise NotImplementedError("non_overlap_masks_for_mem_enc is False in every sam2.1 config")
            mf, pe = self._run_memory_encoder(st, frame_idx, B, masks, out["object_score_logits"],
                                              is_mask_from_pts=True)
            out["maskmem_features"] = mf
            out["maskmem_pos_enc"] = pe
        return out

    def _get_empty_mask_ptr(self, st, frame_idx):
        """svp:769-804: pointer of an all-zero mask prompt (objects absent from a prompted frame)."""
        m = torch.zeros((1, 1, self.image_size, self.image_size), dtype=torch.float32, device=st["device"])
        feats = self._get_image_feature(st, frame_idx)
        return self.engine.mask_as_output(feats, m)["obj_ptr"]

    # ---- propagation -----------------------------------------------------------------------------
    @torch.inference_mode()
    def propagate_in_video_preflight(self, inference_state):
        """svp:807-893."""
        st = inference_state
        st["tracking_has_started"] = True
        B = self._get_obj_num(st)
        temp = st["temp_output_dict_per_obj"]
        output_dict = st["output_dict"]
        cons = st["consolidated_frame_inds"]
        for is_cond in (False, True):
            key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
            frames = set()
            for t in temp.values():
                frames.update(t[key].keys())
            cons[key].update(frames)
            for frame_idx in frames:
                out = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond, run_mem_encoder=True)
                output_dict[key][frame_idx] = out
                self._add_output_per_object(st, frame_idx, out, key)
                if self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1):
                    self._clear_non_cond_mem_around_input(st, frame_idx)
            for t in temp.values():
                t[key].clear()
        for frame_idx in output_dict["cond_frame_outputs"]:
            output_dict["non_cond_frame_outputs"].pop(frame_idx, None)
        for per in st["output_dict_per_obj"].values():
            for frame_idx in per["cond_frame_outputs"]:
                per["non_cond_frame_outputs"].pop(frame_idx, None)
        for frame_idx in cons["cond_frame_outputs"]:
            assert frame_idx in output_dict["cond_frame_outputs"]
            cons["non_cond_frame_outputs"].discard(frame_idx)

    @torch.inference_mode()
    def propagate_in_video(self, inference_state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """svp:911-1025 (generator)."""
        st = inference_state
        self.propagate_in_video_preflight(st)
        output_dict = st["output_dict"]
        cons = st["consolidated_frame_inds"]
        obj_ids = st["obj_ids"]
        num_frames = st["num_frames"]
        B = self._get_obj_num(st)
        if len(output_dict["cond_frame_outputs"]) == 0:
            raise RuntimeError("No points are provided; please add points first")
        clear_non_cond_mem = self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1)
        if start_frame_idx is None:
            start_frame_idx = min(output_dict["cond_frame_outputs"])
        if max_frame_num_to_track is None:
            max_frame_num_to_track = num_frames
        if reverse:
            end = max(start_frame_idx - max_frame_num_to_track + 1, 0)  # Det-SAM2's "+1" (svp:962)
            order = range(start_frame_idx, end - 1, -1) if start_frame_idx > 0 else []
        else:
            end = min(start_frame_idx + max_frame_num_to_track, num_frames - 1)
            order = range(start_frame_idx, end + 1)
        order = list(order)
        if self.encoder_batch_frames > 1:
            self._upcoming = (id(st), [f for f in order if f not in cons["cond_frame_outputs"]
                                       and f not in cons["non_cond_frame_outputs"]])
            self._prefetched = {}
            self._pending = None
        try:
            for frame_idx in order:
                if frame_idx in cons["cond_frame_outputs"]:
                    key = "cond_frame_outputs"
                    current_out = output_dict[key][frame_idx]
                    pred_masks = current_out["pred_masks"]
                    if clear_non_cond_mem:
                        self._clear_non_cond_mem_around_input(st, frame_idx)
                elif frame_idx in cons["non_cond_frame_outputs"]:
                    key = "non_cond_frame_outputs"
                    current_out = output_dict[key][frame_idx]
                    pred_masks = current_out["pred_masks"]
                else:
                    key = "non_cond_frame_outputs"
                    current_out, pred_masks = self._run_single_frame_inference(
                        st, output_dict, frame_idx, batch_size=B, is_init_cond_frame=False, point_inputs=None,
                        mask_inputs=None, reverse=reverse, run_mem_encoder=True)
                    output_dict[key][frame_idx] = current_out
                self._add_output_per_object(st, frame_idx, current_out, key)
                st["frames_already_tracked"][frame_idx] = {"reverse": reverse}
                _, video_res_masks = self._get_orig_video_res_output(st, pred_masks)
                yield frame_idx, obj_ids, video_res_masks
        finally:
            self._upcoming = None
            self._prefetched = {}
            self._pending = None

    def _add_output_per_object(self, st, frame_idx, current_out, storage_key):
        """svp:1027-1058: per-object views sharing storage with the batched output."""
        mf = current_out["maskmem_features"]
        pe = current_out["maskmem_pos_enc"]
        for obj_idx, per in st["output_dict_per_obj"].items():
            s = slice(obj_idx, obj_idx + 1)
            o = {
                "maskmem_features": None,
                "maskmem_pos_enc": None,
                "pred_masks": current_out["pred_masks"][s],
                "obj_ptr": current_out["obj_ptr"][s],
                "object_score_logits": current_out["object_score_logits"][s],
            }
            if mf is not None:
                o["maskmem_features"] = mf[s]
            if pe is not None:
                o["maskmem_pos_enc"] = [x[s] for x in pe]
            per[storage_key][frame_idx] = o

    @torch.inference_mode()
    def clear_all_prompts_in_frame(self, inference_state, frame_idx, obj_id, need_output=True):
        """svp:1061-1131."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        temp = st["temp_output_dict_per_obj"]
        temp[obj_idx]["cond_frame_outputs"].pop(frame_idx, None)
        temp[obj_idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        B = self._get_obj_num(st)
        has_input = any(frame_idx in st["point_inputs_per_obj"][i] or frame_idx in st["mask_inputs_per_obj"][i]
                        for i in range(B))
        if not has_input:
            output_dict = st["output_dict"]
            cons = st["consolidated_frame_inds"]
            cons["cond_frame_outputs"].discard(frame_idx)
            cons["non_cond_frame_outputs"].discard(frame_idx)
            out = output_dict["cond_frame_outputs"].pop(frame_idx, None)
            if out is not None:
                output_dict["non_cond_frame_outputs"][frame_idx] = out
                st["frames_already_tracked"].pop(frame_idx, None)
            for i in range(B):
                per = st["output_dict_per_obj"][i]
                o = per["cond_frame_outputs"].pop(frame_idx, None)
                if o is not None:
                    per["non_cond_frame_outputs"][frame_idx] = o
            if len(output_dict["cond_frame_outputs"]) == 0:
                self._reset_tracking_results(st)
        if not need_output:
            return
        is_cond = any(frame_idx in t["cond_frame_outputs"] for t in temp.values())
        consolidated = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond, run_mem_encoder=False,
                                                                consolidate_at_video_res=True)
        _, video_res_masks = self._get_orig_video_res_output(st, consolidated["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res_masks

    @torch.inference_mode()
    def reset_state(self, inference_state):
        """svp:1134-1143."""
        st = inference_state
        self._reset_tracking_results(st)
        for k in ("obj_id_to_idx", "obj_idx_to_id", "obj_ids", "point_inputs_per_obj", "mask_inputs_per_obj",
                  "output_dict_per_obj", "temp_output_dict_per_obj"):
            st[k].clear()

    def _reset_tracking_results(self, st):
        """svp:1145-1171."""
        for k in ("point_inputs_per_obj", "mask_inputs_per_obj"):
            for v in st[k].values():
                v.clear()
        for k in ("output_dict_per_obj", "temp_output_dict_per_obj"):
            for v in st[k].values():
                v["cond_frame_outputs"].clear()
                v["non_cond_frame_outputs"].clear()
        st["output_dict"]["cond_frame_outputs"].clear()
        st["output_dict"]["non_cond_frame_outputs"].clear()
        st["consolidated_frame_inds"]["cond_frame_outputs"].clear()
        st["consolidated_frame_inds"]["non_cond_frame_outputs"].clear()
        st["tracking_has_started"] = False
        st["frames_already_tracked"].clear()

    # ---- per-frame compute -----------------------------------------------------------------------
    def _get_image_feature(self, st, frame_idx, batch_size=None):
        """svp:1174-1212.  Returns the engine's frame-feature handle; expansion to B objects is a
        view inside the engine (the reference expands with .expand, also views)."""
        cache = st["cached_features"]
        feats = cache.get(frame_idx, None)
        if feats is not None:
            return feats
        E = self.encoder_batch_frames
        in_order = False
        if E > 1 and self._upcoming is not None and self._upcoming[0] == id(st):
            if self._pending is not None and frame_idx in self._pending[0]:
                self._join_pending()
            feats = self._prefetched.pop(frame_idx, None)
            todo = self._upcoming[1]
            in_order = frame_idx in todo
            if in_order:
                del todo[:todo.index(frame_idx) + 1]     # everything up to this frame has been handled
        if feats is None:
            batch = [frame_idx] + (self._next_frames(st, E - 1) if in_order else [])
            if len(batch) > 1:
                for f, o in zip(*self._encode_frames(st, batch, False)):
                    if f == frame_idx:
                        feats = o
                    else:
                        self._prefetched[f] = o
            else:
                row = st["images_idx"].index(frame_idx)
                image = st["images"][row]
                feats = self.engine.encode_image(image)
        if in_order and self.encoder_overlap and self._pending is None:
            # the pass that holds this frame has just started to be consumed: launch the next one behind it
            batch = self._next_frames(st, E)
            if batch:
                self._pending = self._encode_frames(st, batch, True)
        if self.feature_cache_frames <= 1:
            st["cached_features"] = {frame_idx: feats}
        else:
            while len(cache) >= self.feature_cache_frames:
                cache.pop(next(iter(cache)))
            cache[frame_idx] = feats
        return feats

    def _encode_frames(self, st, batch, on_encoder_stream):
        """One encoder pass over the frames of ``batch``: (frames in row order, their features or a PendingFeats)."""
        rows = sorted((st["images_idx"].index(f), f) for f in batch)
        r0, r1 = rows[0][0], rows[-1][0]
        images = st["images"]
        # neighbouring rows (the usual case, forward or reverse) are a view; anything else is gathered
        stack = images[r0:r1 + 1] if r1 - r0 + 1 == len(rows) else images[[r for r, _ in rows]]
        frames = [f for _, f in rows]
        if on_encoder_stream:
            return frames, self.engine.encode_images_async(stack)
        return frames, self.engine.encode_images(stack)

    def _join_pending(self):
        frames, pend = self._pending
        self._pending = None
        for f, o in zip(frames, pend.wait()):
            self._prefetched[f] = o

    def drop_encoded_ahead(self):
        """Forgets features that were encoded ahead of their step (bench.py calls this at the start of its timed region
        so that every frame tracked inside the region is also encoded inside it)."""
        self._prefetched = {}
        self._pending = None

    def _next_frames(self, st, n):
        """Up to ``n`` frames the running propagate call will reach next that have neither cached nor prefetched
        features, are not in the pass in flight, and whose pixels the session still holds."""
        cache, have = st["cached_features"], set(st["images_idx"])
        flying = self._pending[0] if self._pending is not None else ()
        batch = []
        for f in self._upcoming[1]:
            if len(batch) >= n:
                break
            if f not in cache and f not in self._prefetched and f not in flying and f in have:
                batch.append(f)
        return batch

    def _use_multimask(self, is_init_cond_frame, point_inputs):
        """sam2_base.py:922-932 (multimask_output_in_sam and multimask_output_for_tracking are true)."""
        n = 0 if point_inputs is None else point_inputs["point_labels"].size(1)
        return self.cfg.multimask_min_pt_num <= n <= self.cfg.multimask_max_pt_num

    def _run_single_frame_inference(self, st, output_dict, frame_idx, batch_size, is_init_cond_frame, point_inputs,
                                    mask_inputs, reverse, run_mem_encoder, prev_sam_mask_logits=None):
        """svp:1280-1365 over SAM2Base.track_step (sam2_base.py:857-919)."""
        eng = self.engine
        feats = self._get_image_feature(st, frame_idx)
        assert point_inputs is None or mask_inputs is None
        if mask_inputs is not None:
            sam = eng.mask_as_output(feats, mask_inputs)
        else:
            pix_feat = eng.condition_on_memory(
                feats, batch_size, frame_idx, is_init_cond_frame, output_dict, st["num_frames"], reverse,
                st["preloading_memory_cond_frame_idx"])
            coords = labels = None
            if point_inputs is not None:
                coords, labels = point_inputs["point_coords"], point_inputs["point_labels"]
            sam = eng.sam_heads(pix_feat, feats, batch_size, coords, labels, prev_sam_mask_logits,
                                self._use_multimask(is_init_cond_frame, point_inputs))
        maskmem_features = maskmem_pos_enc = None
        if run_mem_encoder and self.num_maskmem > 0:
            maskmem_features, maskmem_pos_enc = eng.encode_memory(
                feats, batch_size, sam["pred_masks"], sam["object_score_logits"],
                is_mask_from_pts=point_inputs is not None)
        sd = st["storage_device"]
        if maskmem_features is not None:
            maskmem_features = maskmem_features.to(torch.bfloat16).to(sd)
        pred_masks_dev = sam["pred_masks"]
        if self.fill_hole_area > 0:
            pred_masks_dev = eng.fill_holes(pred_masks_dev, self.fill_hole_area)
        pred_masks = pred_masks_dev.to(sd)
        maskmem_pos_enc = self._get_maskmem_pos_enc(st, maskmem_pos_enc)
        compact = {
            "maskmem_features": maskmem_features,
            "maskmem_pos_enc": maskmem_pos_enc,
            "pred_masks": pred_masks,
            "obj_ptr": sam["obj_ptr"],
            "object_score_logits": sam["object_score_logits"],
        }
        return compact, pred_masks_dev

    def _run_memory_encoder(self, st, frame_idx, batch_size, low_res_masks, object_score_logits, is_mask_from_pts):
        """svp:1367-1404.  Takes the 256^2 logits: the x4 bilinear to 1024^2 (svp:736-741) is part
        of the engine's memory-encoder seam, where it is fused into the first conv."""
        feats = self._get_image_feature(st, frame_idx)
        mf, pe = self.engine.encode_memory(feats, batch_size, low_res_masks, object_score_logits, is_mask_from_pts)
        mf = mf.to(torch.bfloat16).to(st["storage_device"])
        return mf, self._get_maskmem_pos_enc(st, pe)

    def _get_maskmem_pos_enc(self, st, out_pos_enc):
        """svp:1406-1435: the position encoding is constant; keep one copy in state['constants']."""
        if out_pos_enc is None:
            return None
        consts = st["constants"]
        if "maskmem_pos_enc" not in consts:
            assert isinstance(out_pos_enc, list)
            consts["maskmem_pos_enc"] = [x[0:1].clone() for x in out_pos_enc]
        B = out_pos_enc[0].size(0)
        return [x.expand(B, -1, -1, -1) for x in consts["maskmem_pos_enc"]]

    # ---- window management -----------------------------------------------------------------------
    def release_old_frames(self, inference_state, frame_idx, max_inference_state_frames, pre_frames,
                           release_images=False):
        """svp:1215-1277: drop everything older than the window except the preload bank."""
        st = inference_state
        oldest = frame_idx - max_inference_state_frames
        od = st["output_dict"]
        old_cond = [i for i in od["cond_frame_outputs"].keys() if (pre_frames - 1) < i <= oldest]
        old_non = [i for i in od["non_cond_frame_outputs"].keys() if (pre_frames - 1) < i <= oldest]
        for i in old_non:
            od["non_cond_frame_outputs"].pop(i, None)
            for per in st["output_dict_per_obj"].values():
                per["non_cond_frame_outputs"].pop(i, None)
        for i in old_cond:
            od["cond_frame_outputs"].pop(i, None)
            st["consolidated_frame_inds"]["cond_frame_outputs"].discard(i)
            for per in st["output_dict_per_obj"].values():
                per["cond_frame_outputs"].pop(i, None)
        if release_images:
            old_imgs = set(i for i in st["images_idx"] if (pre_frames - 1) < i <= oldest)
            if old_imgs:
                keep_rows = [r for r, i in enumerate(st["images_idx"]) if i not in old_imgs]
                idx = torch.tensor(keep_rows, dtype=torch.long, device=st["images"].device)
                st["images"] = torch.index_select(st["images"], 0, idx)
                st["images_idx"] = [i for i in st["images_idx"] if i not in old_imgs]
                for i in list(st["cached_features"].keys()):
                    if i in old_imgs:
                        st["cached_features"].pop(i)
            assert len(st["images"]) == len(st["images_idx"])
        # the reference ends with gc.collect() (svp:1277).  Nothing here is cyclic — tensors die with their dict
        # entries by reference count — and a full collection costs 40-80 ms per chunk with torch loaded, so it is
        # only kept for CPU sessions, where it is what the reference does
        if self.device.type != "cuda":
            gc.collect()

    @torch.inference_mode()
    def remove_object(self, inference_state, obj_id, strict=False, need_output=True):
        """svp:1438-1553."""
        st = inference_state
        old_idx = st["obj_id_to_idx"].get(obj_id, None)
        updated = []
        if old_idx is None:
            if not strict:
                return st["obj_ids"], updated
            raise RuntimeError(f"Cannot remove object id {obj_id} as it doesn't exist. "
                               f"All existing object ids: {st['obj_ids']}.")
        if len(st["obj_id_to_idx"]) == 1:
            self.reset_state(st)
            return st["obj_ids"], updated
        input_frames = set(st["point_inputs_per_obj"][old_idx]) | set(st["mask_inputs_per_obj"][old_idx])
        for f in input_frames:
            self.clear_all_prompts_in_frame(st, f, obj_id, need_output=False)
        old_ids = st["obj_ids"]
        old_inds = list(range(len(old_ids)))
        remain = [i for i in old_inds if i != old_idx]
        new_ids = [old_ids[i] for i in remain]
        new_inds = list(range(len(new_ids)))
        remap = dict(zip(remain, new_inds))
        st["obj_id_to_idx"] = dict(zip(new_ids, new_inds))
        st["obj_idx_to_id"] = dict(zip(new_inds, new_ids))
        st["obj_ids"] = new_ids

        def _map_keys(c):
            kv = []
            for k in old_inds:
                v = c.pop(k)
                if k in remap:
                    kv.append((remap[k], v))
            c.update(kv)

        for k in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj"):
            _map_keys(st[k])
        for key in ("cond_frame_outputs", "non_cond_frame_outputs"):
            for f, out in st["output_dict"][key].items():
                out["maskmem_features"] = out["maskmem_features"][remain]
                out["maskmem_pos_enc"] = self._get_maskmem_pos_enc(st, [x[remain] for x in out["maskmem_pos_enc"]])
                out["pred_masks"] = out["pred_masks"][remain]
                out["obj_ptr"] = out["obj_ptr"][remain]
                out["object_score_logits"] = out["object_score_logits"][remain]
                self._add_output_per_object(st, f, out, key)
        if need_output:
            temp = st["temp_output_dict_per_obj"]
            for f in input_frames:
                is_cond = any(f in t["cond_frame_outputs"] for t in temp.values())
                c = self._consolidate_temp_output_across_obj(st, f, is_cond=is_cond, run_mem_encoder=False,
                                                             consolidate_at_video_res=True)
                _, vm = self._get_orig_video_res_output(st, c["pred_masks_video_res"])
                updated.append((f, vm))
        return st["obj_ids"], updated

    def _clear_non_cond_mem_around_input(self, st, frame_idx):
        """svp:1555-1571."""
        r = 1  # memory_temporal_stride_for_eval
        non_cond = st["output_dict"]["non_cond_frame_outputs"]
        for t in range(frame_idx - r * self.num_maskmem, frame_idx + r * self.num_maskmem + 1):
            non_cond.pop(t, None)
            for per in st["output_dict_per_obj"].values():
                per["non_cond_frame_outputs"].pop(t, None)
