"""Host-side mirror of the reference's ``SAM2ImagePredictor`` (/root/reference/sam2/sam2_image_predictor.py:20-466): embed
an image once (``set_image``), then decode masks for point / box / mask prompts (``predict``) — same method names,
arguments, return values and error behaviour — over an *engine* that owns all arithmetic (``CudaEngine`` in the product,
the fp32 oracle in tests).  SURVEY.md 8f rank 4: it reuses the image encoder, prompt encoder and mask decoder kernels of
the video path; nothing here computes on the CPU on behalf of a missing CUDA library.

Engine seams used:  encode_image(fp16 [3,S,S]), condition_on_memory(..., is_init_cond_frame=True) (adds ``no_mem_embed``,
sam2_image_predictor.py:119-121), decode_masks(...) (prompt encoder + mask decoder WITHOUT the tracking epilogue: the
image predictor calls ``sam_mask_decoder`` directly, :405-420, so masks are not gated by the object score),
resize_masks, connected components for the optional hole / sprinkle post-processing (utils/transforms.py:80-126).
"""
import numpy as np
import torch
import torch.nn.functional as F

IMG_MEAN = (0.485, 0.456, 0.406)
IMG_STD = (0.229, 0.224, 0.225)


class SAM2ImagePredictor:
    def __init__(self, engine, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0, **kwargs):
        self.engine = engine
        self.cfg = engine.cfg
        self.mask_threshold = mask_threshold
        self.max_hole_area = max_hole_area
        self.max_sprinkle_area = max_sprinkle_area
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False

    @classmethod
    def from_pretrained(cls, model_id, **kwargs):
        """sam2_image_predictor.py:69-83."""
        from .build_sam import build_sam2_hf
        return cls(build_sam2_hf(model_id, **kwargs).engine, **kwargs)

    @property
    def device(self):
        return self.engine.device

    # ---- image side ------------------------------------------------------------------------------
    def _transform(self, image):
        """utils/transforms.py:27-41: ToTensor -> Resize((S, S)) (bilinear, antialiased, on the float tensor) ->
        Normalize.  The arithmetic is torchvision's (third party); restated with F.interpolate.  The engine takes the
        frame as fp16 (the video path's frame format); its patch-embedding operand is bf16 either way."""
        if isinstance(image, np.ndarray):
            if image.ndim != 3 or image.shape[2] != 3:
                raise NotImplementedError("Image format not supported")
            x = torch.from_numpy(np.ascontiguousarray(image))
            h, w = image.shape[:2]
        else:
            try:
                from PIL.Image import Image
            except ImportError:  # pragma: no cover
                Image = ()
            if not isinstance(image, Image):
                raise NotImplementedError("Image format not supported")
            w, h = image.size
            x = torch.from_numpy(np.asarray(image.convert("RGB")).copy())
        x = x.permute(2, 0, 1).to(torch.float32)
        if x.max() > 1.0 or image.dtype == np.uint8 if isinstance(image, np.ndarray) else True:
            x = x / 255.0                                           # ToTensor scales uint8 images only
        S = self.cfg.image_size
        x = F.interpolate(x[None], size=(S, S), mode="bilinear", align_corners=False, antialias=True)[0]
        mean = torch.tensor(IMG_MEAN, dtype=torch.float32)[:, None, None]
        std = torch.tensor(IMG_STD, dtype=torch.float32)[:, None, None]
        return ((x - mean) / std), (h, w)

    @torch.inference_mode()
    def set_image(self, image):
        """sam2_image_predictor.py:86-129."""
        self.reset_predictor()
        x, hw = self._transform(image)
        self._orig_hw = [hw]
        self._features = [self._embed(x)]
        self._is_image_set = True

    @torch.inference_mode()
    def set_image_batch(self, image_list):
        """sam2_image_predictor.py:132-172."""
        self.reset_predictor()
        assert isinstance(image_list, list)
        self._orig_hw, self._features = [], []
        for image in image_list:
            assert isinstance(image, np.ndarray), "Images are expected to be an np.ndarray in RGB format, and of shape  HWC"
            x, hw = self._transform(image)
            self._orig_hw.append(hw)
            self._features.append(self._embed(x))
        self._is_image_set = True
        self._is_batch = True

    def _embed(self, x):
        # the CUDA engine takes frames in the video path's fp16 format (its patch-embedding operand is bf16 anyway); an
        # engine that can take the fp32 image as the reference does says so
        dtype = getattr(self.engine, "image_dtype", torch.float16)
        feats = self.engine.encode_image(x.to(dtype).to(self.device))
        # "Add no_mem_embed, which is added to the lowest res feat. map during training on videos" (:119-121)
        pix = self.engine.condition_on_memory(feats, 1, 0, True, None, 1, False, None)
        return {"feats": feats, "pix": pix.clone() if hasattr(pix, "clone") else pix}

    def get_image_embedding(self):
        """sam2_image_predictor.py:440-452: [1, C, h, w] embedding of the current image."""
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        assert self._features is not None, "Features must exist if an image has been set."
        pix = self._features[-1]["pix"]
        fs = self.cfg.feat_size
        if pix.dim() == 4:
            return pix
        return pix.reshape(1, fs, fs, -1).permute(0, 3, 1, 2)

    def reset_predictor(self):
        """sam2_image_predictor.py:459-466."""
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False

    # ---- prompts ---------------------------------------------------------------------------------
    def _transform_coords(self, coords, normalize, orig_hw):
        """utils/transforms.py:51-69."""
        if normalize:
            assert orig_hw is not None
            h, w = orig_hw
            coords = coords.clone()
            coords[..., 0] = coords[..., 0] / w
            coords[..., 1] = coords[..., 1] / h
        return coords * self.cfg.image_size

    def _prep_prompts(self, point_coords, point_labels, box, mask_logits, normalize_coords, img_idx=-1):
        """sam2_image_predictor.py:305-335."""
        unnorm_coords = labels = unnorm_box = mask_input = None
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied if point_coords is supplied."
            pc = torch.as_tensor(point_coords, dtype=torch.float)
            unnorm_coords = self._transform_coords(pc, normalize_coords, self._orig_hw[img_idx])
            labels = torch.as_tensor(point_labels, dtype=torch.int)
            if unnorm_coords.dim() == 2:
                unnorm_coords, labels = unnorm_coords[None, ...], labels[None, ...]
        if box is not None:
            b = torch.as_tensor(box, dtype=torch.float)
            unnorm_box = self._transform_coords(b.reshape(-1, 2, 2), normalize_coords, self._orig_hw[img_idx])
        if mask_logits is not None:
            mask_input = torch.as_tensor(mask_logits, dtype=torch.float)
            if mask_input.dim() == 3:
                mask_input = mask_input[None, :, :, :]
        return mask_input, unnorm_coords, labels, unnorm_box

    @torch.inference_mode()
    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False, normalize_coords=True):
        """sam2_image_predictor.py:237-303 -> (masks [C,H,W], iou_predictions [C], low_res_masks [C,256,256]) numpy."""
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        mask_input, coords, labels, boxes = self._prep_prompts(point_coords, point_labels, box, mask_input, normalize_coords)
        masks, ious, low = self._predict(coords, labels, boxes, mask_input, multimask_output, return_logits=return_logits)
        # the reference returns float32 arrays in both modes: thresholded masks come back as 0.0 / 1.0 (:299-302)
        return (masks.squeeze(0).float().cpu().numpy(), ious.squeeze(0).float().cpu().numpy(),
                low.squeeze(0).float().cpu().numpy())

    @torch.inference_mode()
    def predict_batch(self, point_coords_batch=None, point_labels_batch=None, box_batch=None, mask_input_batch=None,
                      multimask_output=True, return_logits=False, normalize_coords=True):
        """sam2_image_predictor.py:175-235."""
        assert self._is_batch, "This function should only be used when in batched mode"
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image_batch(...) before mask prediction.")
        all_masks, all_ious, all_low = [], [], []
        for i in range(len(self._features)):
            pc = point_coords_batch[i] if point_coords_batch is not None else None
            pl = point_labels_batch[i] if point_labels_batch is not None else None
            bx = box_batch[i] if box_batch is not None else None
            mi = mask_input_batch[i] if mask_input_batch is not None else None
            mask_input, coords, labels, boxes = self._prep_prompts(pc, pl, bx, mi, normalize_coords, img_idx=i)
            masks, ious, low = self._predict(coords, labels, boxes, mask_input, multimask_output,
                                             return_logits=return_logits, img_idx=i)
            all_masks.append(masks.squeeze(0).float().cpu().numpy())
            all_ious.append(ious.squeeze(0).float().cpu().numpy())
            all_low.append(low.squeeze(0).float().cpu().numpy())
        return all_masks, all_ious, all_low

    def _predict(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                 return_logits=False, img_idx=-1):
        """sam2_image_predictor.py:337-438 on batched torch prompts already in model coordinates."""
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        concat = (point_coords, point_labels) if point_coords is not None else None
        if boxes is not None:
            box_coords = boxes.reshape(-1, 2, 2)
            box_labels = torch.tensor([[2, 3]], dtype=torch.int).repeat(boxes.size(0), 1)
            if concat is not None:      # boxes go first (:387-392)
                concat = (torch.cat([box_coords, concat[0]], dim=1), torch.cat([box_labels, concat[1]], dim=1))
            else:
                concat = (box_coords, box_labels)
        f = self._features[img_idx]
        B = concat[0].shape[0] if concat is not None else (mask_input.shape[0] if mask_input is not None else 1)
        pix = f["pix"]
        if B > 1:                        # repeat_image (:403-404, mask_decoder.py:192-197)
            pix = pix.expand(B, *pix.shape[1:]).contiguous()
        coords, labels = (concat if concat is not None else (None, None))
        low_res_masks, ious = self.engine.decode_masks(pix, f["feats"], B, coords, labels, mask_input, bool(multimask_output))
        masks = self._postprocess_masks(low_res_masks, self._orig_hw[img_idx])
        low_res_masks = torch.clamp(low_res_masks, -32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks, ious, low_res_masks

    def _postprocess_masks(self, masks, orig_hw):
        """utils/transforms.py:80-126: optional hole / sprinkle removal on the low-resolution logits, then bilinear to
        the original size.  As in the reference, a failure of the connected-components step only skips that step."""
        masks = masks.float()
        if self.max_hole_area > 0 or self.max_sprinkle_area > 0:
            try:
                cc = self.engine.connected_components
                flat = masks.flatten(0, 1).unsqueeze(1)
                if self.max_hole_area > 0:
                    labels, areas = cc((flat <= self.mask_threshold).to(torch.uint8))
                    hole = ((labels > 0) & (areas <= self.max_hole_area)).reshape_as(masks)
                    masks = torch.where(hole, torch.full_like(masks, self.mask_threshold + 10.0), masks)
                if self.max_sprinkle_area > 0:
                    labels, areas = cc((flat > self.mask_threshold).to(torch.uint8))
                    spr = ((labels > 0) & (areas <= self.max_sprinkle_area)).reshape_as(masks)
                    masks = torch.where(spr, torch.full_like(masks, self.mask_threshold - 10.0), masks)
            except Exception as e:  # utils/transforms.py:110-120
                import warnings
                warnings.warn(f"{e}\n\nSkipping the post-processing step due to the error above.", category=UserWarning,
                              stacklevel=2)
        B, C = masks.shape[:2]
        out = self.engine.resize_masks(masks.reshape(B * C, 1, *masks.shape[-2:]).contiguous(), orig_hw[0], orig_hw[1])
        return out.reshape(B, C, orig_hw[0], orig_hw[1])
