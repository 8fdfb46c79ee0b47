"""Host-side mirror of Det-SAM2's stream driver ``VideoProcessor``
(/root/reference/det_sam2_inference/det_sam2_RT.py:25-684): frames are accumulated into a buffer,
every ``frame_buffer_size`` frames the detector produces box prompts for the frames that fall on the
``detect_interval`` grid, the buffer is appended to the predictor state, prompts are added, the last
``max_frame_num_to_track`` frames are re-tracked in REVERSE from the newest frame, and frames older
than ``max_inference_state_frames`` are released (constant-memory window).  An optional pickled
"preload memory bank" seeds the state.

Same constructor keywords, methods (``process_frame``, ``Detect_and_SAM2_inference``,
``detect_predict``, ``Detect_2_SAM2_Prompt``, ``run``, ``clear``, ``save/load_inference_state``) and
result type (``video_segments: {frame_idx: {obj_id: bool ndarray [1,H,W]}}``) as the reference.

Out of scope, as in SURVEY.md §8: the YOLOv8 detector itself ("left untouched, timed separately") is
*injected* — ``detector(frames_bgr) -> [[{"coordinates", "class", "confidence"}, ...], ...]`` — and
loaded from ultralytics only when ``detect_model_weights`` is given and no detector is passed;
matplotlib rendering (``vis_frame_stride``, ``visualize_prompt``) is not provided.
"""
import os
import time
import pickle

import numpy as np
import torch


def _yolo_detector(weights, conf):
    """det_sam2_RT.py:96,230-245: the reference's own detector call, unchanged."""
    try:
        from ultralytics import YOLO
    except ImportError as e:  # fail loudly: there is no stand-in detector in the product
        raise ImportError("ultralytics is required for detect_model_weights; pass detector=... instead") from e
    model = YOLO(weights)

    def detect(frames_bgr):
        out = []
        for result in model(frames_bgr, stream=True, conf=conf, iou=0.1, verbose=False):
            dets = []
            if result.boxes is not None:
                for box in result.boxes:
                    dets.append({"coordinates": box.xyxy[0].cpu().numpy(), "class": box.cls.cpu().numpy(),
                                 "confidence": box.conf.cpu().numpy()})
            out.append(dets)
        return out

    return detect


class VideoProcessor:
    def __init__(self, output_dir=None, sam2_checkpoint=None, model_cfg="configs/sam2.1/sam2.1_hiera_l.yaml",
                 detect_model_weights=None, detect_confidence=0.85, skip_classes=frozenset({11, 14, 15, 19}),
                 vis_frame_stride=-1, visualize_prompt=False, frame_buffer_size=30, detect_interval=30,
                 max_frame_num_to_track=60, max_inference_state_frames=60, load_inference_state_path=None,
                 save_inference_state_path=None, *, predictor=None, detector=None, device="cuda", object_stats=False,
                 frames_on_device=None, offload_state_to_cpu=None, detector_preproc=None):
        if vis_frame_stride != -1 or visualize_prompt:
            raise NotImplementedError("matplotlib rendering is outside the hot path; use vis_frame_stride=-1")
        if save_inference_state_path is not None:
            # det_sam2_RT.py:67-68
            assert max_inference_state_frames == -1, \
                "saving a preload memory bank requires max_inference_state_frames == -1 (nothing may be released)"
        self.output_dir = output_dir
        self.sam2_checkpoint = sam2_checkpoint
        self.model_cfg = model_cfg
        self.detect_model_weights = detect_model_weights
        self.detect_confidence = detect_confidence
        self.skip_classes = set(skip_classes)
        self.vis_frame_stride = vis_frame_stride
        self.visualize_prompt = visualize_prompt
        self.frame_buffer_size = frame_buffer_size
        self.detect_interval = detect_interval
        self.frame_buffer = []
        self.max_frame_num_to_track = max_frame_num_to_track
        self.max_inference_state_frames = max_inference_state_frames
        self.load_inference_state_path = load_inference_state_path
        self.save_inference_state_path = save_inference_state_path
        self.pre_frames = 0
        self.special_classes = 11  # det_sam2_RT.py:71 (pockets: many instances of one class)
        self.special_classes_detection = []
        self.special_classes_count = 0
        if predictor is None:
            from .build_sam import build_sam2_video_predictor
            # the reverse pass re-visits the last max_frame_num_to_track frames of every chunk: keep their backbone
            # features (18 MB per 1024^2 frame) instead of the reference's single-frame cache (svp:1190), which
            # re-encodes each frame M/K times
            cache = max(1, max_frame_num_to_track + 4) if str(device).startswith("cuda") else 1
            predictor = build_sam2_video_predictor(model_cfg, sam2_checkpoint, device=device, feature_cache_frames=cache)
        self.predictor = predictor
        if detector is None and detect_model_weights is not None:
            detector = _yolo_detector(detect_model_weights, detect_confidence)
        self.detect_model = detector
        self.video_segments = {}
        # addition: per frame and object (area, centroid x, centroid y) of the thresholded mask,
        # computed on the GPU in integer arithmetic (ds2_mask_pack_stats) — what Det-SAM2's post-processor derives
        # from the boolean masks with cv2.moments (postprocess_det_sam2.py:331-343); None for an empty mask
        self.object_stats = bool(object_stats)
        # addition: keep the session's fp16 frames in HBM (the reference's `offload_video_to_cpu=False`, svp:44-53)
        # instead of its default host tensor: frames are ingested on the device (ds2_ingest_frames) and neither the
        # per-chunk torch.cat (svp:196) nor the per-step upload (svp:1184-1186) touches host memory.  A window of
        # S + K = 90 frames is 0.57 GB of the 180 GB.  None = on for a CUDA predictor; False = the reference's
        # host-resident frames.
        if frames_on_device is None:
            frames_on_device = getattr(getattr(predictor, "device", None), "type", "cpu") == "cuda"
        self.frames_on_device = bool(frames_on_device)
        # Where a preloaded bank lives once it is loaded (init_preloading_state, svp:123-156).  The reference moves it to
        # the host by default (offload_state_to_cpu=True: its constant-memory design targets 24 GB cards) and every memory
        # read then drags the stored features back over PCIe — measured on configs[2] (base_plus, 720p, 16 objects, 10-frame
        # bank, 2000-frame stream): 8.6 video frames/s host-bound (profiles/r2_config3_stream_2000_frames_state_on_host.json).
        # A 10-frame bank is 0.12 GB of the 180 GB: None = keep it in HBM for a CUDA predictor (results are bit-identical,
        # tests/test_fullsize_gpu.py::test_offload_state_to_cpu_on_cuda_engine); True = the reference's behaviour.
        if offload_state_to_cpu is None:
            offload_state_to_cpu = getattr(getattr(predictor, "device", None), "type", "cpu") != "cuda"
        self.offload_state_to_cpu = bool(offload_state_to_cpu)
        # addition: `detector_preproc` = a detsam2_b200.detector_preproc.DeviceLetterbox.  The chunk's
        # uint8 frames are then uploaded ONCE; the detector is called with the letterboxed [n, 3, H, W] tensor built on
        # the device from that copy (ultralytics skips its host pre-processing for tensor input) and must return its
        # boxes in the pixels of that tensor — they are mapped back to frame pixels here; the SAM 2 ingest reads the same
        # device copy.  None = the reference's flow (BGR ndarrays to the detector).
        self.detector_preproc = detector_preproc
        self._chunk_dev_u8 = None
        self.video_stats = {}
        self.timings = {}
        self.inference_state = None
        if output_dir:
            os.makedirs(output_dir, exist_ok=True)

    # ---- detector seam (det_sam2_RT.py:201-269) ----------------------------------------------------
    def detect_predict(self, images, past_num_frames):
        detection_results = {}
        if self.detect_interval == -1:
            return detection_results
        selected, absolute = [], []
        for i, image in enumerate(images):
            frame_idx = past_num_frames + i
            if frame_idx % self.detect_interval == 0:
                selected.append(i if self.detector_preproc is not None else
                                np.ascontiguousarray(image[..., ::-1]))  # RGB -> BGR, what YOLO was trained on
                absolute.append(frame_idx)
        if not selected:
            return detection_results
        if self.detect_model is None:
            raise RuntimeError("detect_interval != -1 but no detector was provided")
        if self.detector_preproc is not None:
            dev = self._upload_chunk(images)
            hw = tuple(images[0].shape[:2])
            results = self.detect_model(self.detector_preproc(dev[selected]))
            results = [[dict(d, coordinates=self.detector_preproc.unletterbox(d["coordinates"], hw)[0]) for d in dets]
                       for dets in results]
        else:
            results = self.detect_model(selected)
        for i, dets in enumerate(results):
            dets = list(dets)
            if not self.special_classes_detection:
                self.special_classes_count = 0
            n_special = sum(1 for d in dets if int(np.asarray(d["class"]).reshape(-1)[0]) == self.special_classes)
            if n_special > self.special_classes_count:
                self.special_classes_detection = [d["coordinates"] for d in dets
                                                  if int(np.asarray(d["class"]).reshape(-1)[0]) == self.special_classes]
                self.special_classes_count = n_special
            detection_results[f"frame_{absolute[i]}"] = dets
        return detection_results

    def _upload_chunk(self, images):
        """The chunk's uint8 RGB frames as ONE device tensor [n, H, W, 3] (pinned staging, one asynchronous copy), shared by
        the detector pre-processing and the SAM 2 ingest of the same chunk."""
        if self._chunk_dev_u8 is None:
            dev = self.predictor.device
            n, (H, W) = len(images), images[0].shape[:2]
            stage = self.__dict__.get("_u8_stage")
            if stage is None or stage.numel() < n * H * W * 3:
                stage = self._u8_stage = torch.empty(n * H * W * 3, dtype=torch.uint8).pin_memory()
            view = stage[: n * H * W * 3].view(n, H, W, 3)
            host = view.numpy()
            for k, im in enumerate(images):
                np.copyto(host[k], im)
            self._chunk_dev_u8 = view.to(dev, non_blocking=True)
        return self._chunk_dev_u8

    # ---- prompts (det_sam2_RT.py:271-316) ----------------------------------------------------------
    def Detect_2_SAM2_Prompt(self, detection_results_json):
        if not detection_results_json:
            return self.inference_state
        for key, detections in detection_results_json.items():
            ann_frame_idx = int(key.replace("frame_", ""))
            boxes, repeats = {}, []
            for det in detections:
                obj_class = int(np.asarray(det["class"]).reshape(-1)[0])
                if obj_class in self.skip_classes:
                    continue
                box = np.array(det["coordinates"], dtype=np.float32)
                if obj_class in boxes:
                    # a second box of the same class on one frame is NOT a replacement in the reference: its
                    # add_new_points_or_box finds the first call's temporary output and feeds those (clamped) mask
                    # logits back as a dense prompt (det_sam2_RT.py:288-302 -> svp:470-482), so the repeats are
                    # replayed one by one, in detection order, after the first occurrences
                    repeats.append((obj_class, box))
                else:
                    boxes[obj_class] = box
            if not boxes:
                continue
            if hasattr(self.predictor, "add_new_boxes"):
                # all first-occurrence boxes of the frame in one B-wide decoder call
                self.predictor.add_new_boxes(self.inference_state, ann_frame_idx, boxes)
            else:
                for obj_class, box in boxes.items():
                    self.predictor.add_new_points_or_box(inference_state=self.inference_state, frame_idx=ann_frame_idx,
                                                         obj_id=obj_class, box=box)
            for obj_class, box in repeats:
                self.predictor.add_new_points_or_box(inference_state=self.inference_state, frame_idx=ann_frame_idx,
                                                     obj_id=obj_class, box=box)
        return self.inference_state

    # ---- one chunk (det_sam2_RT.py:342-411) --------------------------------------------------------
    def Detect_and_SAM2_inference(self, frame_idx):
        tm = self.timings
        t0 = time.perf_counter()
        past_num_frames = self.inference_state["num_frames"] if self.inference_state else 0
        detection_results_json = self.detect_predict(self.frame_buffer, past_num_frames)
        t1 = time.perf_counter()
        # frames for the predictor: the device copy the detector pre-processing already uploaded, if there is one
        source = self._chunk_dev_u8 if self._chunk_dev_u8 is not None else self.frame_buffer
        self._chunk_dev_u8 = None
        if self.inference_state is None:
            self.inference_state = self.predictor.init_state(video_path=source,
                                                             offload_video_to_cpu=not self.frames_on_device)
        else:
            self.inference_state = self.predictor.update_state(video_path=source,
                                                               inference_state=self.inference_state)
        t2 = time.perf_counter()
        try:
            self.inference_state = self.Detect_2_SAM2_Prompt(detection_results_json)
        except RuntimeError as e:
            if "reset_state" in str(e):
                self.predictor.reset_state(self.inference_state)
                self.inference_state = self.Detect_2_SAM2_Prompt(detection_results_json)
            else:
                raise
        t3 = time.perf_counter()
        pending = []
        for out_frame_idx, out_obj_ids, out_mask_logits in self.predictor.propagate_in_video(
                self.inference_state, start_frame_idx=frame_idx,
                max_frame_num_to_track=self.max_frame_num_to_track, reverse=True):
            if out_frame_idx >= self.pre_frames:
                if out_mask_logits.is_cuda:
                    # threshold on the device, copy into pinned memory WITHOUT synchronising: a blocking copy
                    # per frame (det_sam2_RT.py:396-399) idles the GPU while the host prepares the next step
                    stats = self._stats_to_pinned(out_mask_logits, len(pending)) if self.object_stats else None
                    pending.append((out_frame_idx, list(out_obj_ids), self._masks_to_pinned(out_mask_logits, len(pending)),
                                    stats))
                else:
                    self.video_segments[out_frame_idx] = self._masks_to_host(out_obj_ids, out_mask_logits)
        t4 = t5 = time.perf_counter()
        if pending:
            torch.cuda.current_stream().synchronize()
            t5 = time.perf_counter()
            # the pinned buffers are re-used by the next chunk, so every frame's [B,1,Hv,Wv] boolean block is copied
            # out — 1 GB per chunk at 16 objects x 1024^2, first-touch page faults included.  numpy releases the GIL
            # for these copies: a few worker threads bring the 0.18 s per chunk of a single thread down ~5x
            copies = list(self._copy_pool().map(lambda p: p[2].numpy().copy(), pending))
            for (out_frame_idx, ids, host, stats), m in zip(pending, copies):
                self.video_segments[out_frame_idx] = {oid: m[i] for i, oid in enumerate(ids)}
                if stats is not None:
                    st_ = stats.numpy()
                    self.video_stats[out_frame_idx] = {
                        oid: (None if st_[i, 0] == 0 else (int(st_[i, 0]), st_[i, 1] / st_[i, 0], st_[i, 2] / st_[i, 0]))
                        for i, oid in enumerate(ids)}
        t6 = time.perf_counter()
        if self.max_inference_state_frames != -1:
            self.predictor.release_old_frames(self.inference_state, frame_idx, self.max_inference_state_frames,
                                              self.pre_frames, release_images=self.vis_frame_stride == -1)
        t7 = time.perf_counter()
        # host wall time per phase of the chunk (the device runs asynchronously: "gpu_wait" is how long the host
        # idled at the end of the propagate loop for queued device work, i.e. the chunk was device-bound by that much)
        for k, v in (("detect", t1 - t0), ("frames", t2 - t1), ("prompt", t3 - t2), ("propagate_enqueue", t4 - t3),
                     ("gpu_wait", t5 - t4), ("hand_off", t6 - t5), ("release", t7 - t6)):
            tm[k] = tm.get(k, 0.0) + v

    def _copy_pool(self):
        pool = self.__dict__.get("_pool")
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._pool = ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1)),
                                                   thread_name_prefix="ds2-handoff")
        return pool

    def _masks_to_pinned(self, mask_logits, slot):
        """(logits > 0) -> pinned host buffer `slot` of a per-chunk ring, asynchronously on the current stream."""
        ring = self.__dict__.setdefault("_pinned_ring", [])
        shape = tuple(mask_logits.shape)
        while len(ring) <= slot:
            ring.append(None)
        if ring[slot] is None or tuple(ring[slot].shape) != shape:
            ring[slot] = torch.empty(shape, dtype=torch.bool).pin_memory()
        ring[slot].copy_(mask_logits > 0.0, non_blocking=True)
        return ring[slot]

    def _stats_to_pinned(self, mask_logits, slot):
        from . import ops
        ring = self.__dict__.setdefault("_stats_ring", [])
        B = mask_logits.shape[0]
        while len(ring) <= slot:
            ring.append(None)
        if ring[slot] is None or ring[slot][0].shape[0] != B:
            ring[slot] = (torch.empty((B, 3), dtype=torch.int64, device=mask_logits.device),
                          torch.empty((B, 3), dtype=torch.int64).pin_memory())
        dev, host = ring[slot]
        ops.mask_pack_stats(mask_logits, bits=False, stats=dev)
        host.copy_(dev, non_blocking=True)
        return host

    @staticmethod
    def _masks_to_host(obj_ids, mask_logits):
        """det_sam2_RT.py:396-399 does B x ``(logits[i] > 0).cpu().numpy()``; here the threshold runs
        once on the device and ONE copy brings all objects back."""
        m = (mask_logits > 0.0).cpu().numpy()
        return {oid: m[i] for i, oid in enumerate(obj_ids)}

    def process_frame(self, frame_idx, frame):
        """det_sam2_RT.py:421-435."""
        self.frame_buffer.append(frame)
        if len(self.frame_buffer) >= self.frame_buffer_size:
            self.Detect_and_SAM2_inference(frame_idx)
            self.frame_buffer.clear()
        return self.inference_state

    def clear(self):
        """det_sam2_RT.py:189-199."""
        self.frame_buffer = []
        self.pre_frames = 0
        self.special_classes_detection = []
        self.video_segments = {}
        self.video_stats = {}
        self.inference_state = None

    # ---- preload bank (det_sam2_RT.py:489-503) -----------------------------------------------------
    def save_inference_state(self, save_path):
        directory = os.path.dirname(save_path)
        if directory and not os.path.exists(directory):
            os.makedirs(directory)
        if str(save_path).endswith(".ds2bank"):
            # compact, versioned format (bank_format.py); any other suffix keeps the reference's pickle so that
            # banks stay exchangeable with Det-SAM2
            from .bank_format import save_bank
            save_bank(self.inference_state, save_path)
            return
        # the pickle must stay loadable by Det-SAM2 itself: `cached_features` holds engine handles (FrameFeats: a class
        # the reference cannot import, ~18 MB of transient backbone features per cached frame) where the reference
        # expects (image, backbone_out) tuples — both sides recompute features on demand, so write it empty
        st = dict(self.inference_state)
        st["cached_features"] = {}
        with open(save_path, "wb") as f:
            pickle.dump(st, f)

    def load_inference_state(self, load_path):
        if str(load_path).endswith(".ds2bank"):
            from .bank_format import load_bank
            return load_bank(load_path)
        with open(load_path, "rb") as f:
            return pickle.load(f)

    def load_frames_from_folder(self, folder_path):
        import cv2
        frames = []
        for name in sorted(f for f in os.listdir(folder_path) if f.endswith((".png", ".jpg", ".jpeg"))):
            frame = cv2.imread(os.path.join(folder_path, name))
            if frame is None:
                continue
            frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
        return frames

    # ---- whole stream (det_sam2_RT.py:526-640) -----------------------------------------------------
    def run(self, video_path=None, frame_dir=None, output_video_segments_pkl_path=None,
            output_special_classes_detection_pkl_path=None, frames=None):
        """``frames`` (an iterable of RGB uint8 ndarrays) is an addition for in-memory / synthetic
        streams; ``video_path`` and ``frame_dir`` behave as in the reference."""
        if self.load_inference_state_path is not None:
            st = self.load_inference_state(self.load_inference_state_path)
            st["preloading_memory_cond_frame_idx"] = list(st["output_dict"]["cond_frame_outputs"].keys())
            st["preloading_memory_non_cond_frames_idx"] = list(st["output_dict"]["non_cond_frame_outputs"].keys())
            self.pre_frames = st["num_frames"]
            self.inference_state = st
            self.predictor.init_preloading_state(st, offload_video_to_cpu=not self.frames_on_device,
                                                 offload_state_to_cpu=self.offload_state_to_cpu)

        def stream():
            if frames is not None:
                yield from frames
            elif video_path is not None:
                import cv2
                cap = cv2.VideoCapture(video_path)
                if not cap.isOpened():
                    raise IOError(f"cannot open video {video_path}")
                while True:
                    ret, frame = cap.read()
                    if not ret:
                        break
                    yield cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
                cap.release()
            elif frame_dir is not None:
                yield from self.load_frames_from_folder(frame_dir)
            else:
                raise ValueError("one of video_path, frame_dir or frames is required")

        frame_idx = 0
        for frame_rgb in stream():
            self.inference_state = self.process_frame(self.pre_frames + frame_idx, frame_rgb)
            frame_idx += 1
        if self.frame_buffer:  # tail shorter than frame_buffer_size
            self.Detect_and_SAM2_inference(frame_idx=self.pre_frames + frame_idx - 1)
            self.frame_buffer.clear()
        # results are re-based so that they do not count the preload frames (det_sam2_RT.py:612)
        self.video_segments = {idx - self.pre_frames: seg for idx, seg in self.video_segments.items()
                               if idx >= self.pre_frames}
        self.video_stats = {idx - self.pre_frames: v for idx, v in self.video_stats.items() if idx >= self.pre_frames}
        if output_video_segments_pkl_path:
            with open(output_video_segments_pkl_path, "wb") as f:
                pickle.dump(self.video_segments, f)
        if output_special_classes_detection_pkl_path:
            with open(output_special_classes_detection_pkl_path, "wb") as f:
                pickle.dump(self.special_classes_detection, f)
        if self.save_inference_state_path is not None:
            self.save_inference_state(self.save_inference_state_path)
        return self.video_segments
