"""Factory with the reference's name and signature (/root/reference/sam2/build_sam.py:111-146).

``config_file`` is one of the reference's ``configs/sam2.1/sam2.1_hiera_{t,s,b+,l}.yaml`` names (hydra
is not used: the YAML values are restated in detsam2_b200.config).  ``ckpt_path`` is a reference
checkpoint (``torch.load(path)["model"]``, strict key match, build_sam.py:166-178); when it is None
the model is initialised with seeded synthetic weights (there are no checkpoints offline).
Eval-time overrides of build_sam.py:126-135 are part of the config defaults: dynamic multimask via
stability (delta 0.05, thresh 0.98), binarize_mask_from_pts_for_mem_enc, fill_hole_area = 8.
"""
import torch

from .config import get_config
from .predictor import SAM2VideoPredictor
from .weights import param_shapes, synthetic_state_dict


# build_sam.py:33-66, the SAM 2.1 entries (the SAM 2.0 configurations are not restated in detsam2_b200.config)
HF_MODEL_ID_TO_FILENAMES = {
    "facebook/sam2.1-hiera-tiny": ("configs/sam2.1/sam2.1_hiera_t.yaml", "sam2.1_hiera_tiny.pt"),
    "facebook/sam2.1-hiera-small": ("configs/sam2.1/sam2.1_hiera_s.yaml", "sam2.1_hiera_small.pt"),
    "facebook/sam2.1-hiera-base-plus": ("configs/sam2.1/sam2.1_hiera_b+.yaml", "sam2.1_hiera_base_plus.pt"),
    "facebook/sam2.1-hiera-large": ("configs/sam2.1/sam2.1_hiera_l.yaml", "sam2.1_hiera_large.pt"),
}


def load_state_dict(cfg, ckpt_path=None, seed=0):
    if ckpt_path is None:
        return synthetic_state_dict(cfg, seed)
    sd = torch.load(ckpt_path, map_location="cpu", weights_only=True)["model"]
    want = param_shapes(cfg)
    missing = [k for k in want if k not in sd]
    unexpected = [k for k in sd if k not in want]
    if missing or unexpected:
        raise RuntimeError(f"checkpoint does not match {cfg.name}: missing {missing[:5]} unexpected {unexpected[:5]}")
    for k, shape in want.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise RuntimeError(f"checkpoint tensor {k} has shape {tuple(sd[k].shape)}, expected {tuple(shape)}")
    return sd


# hydra override keys the reference's callers pass (build_sam.py:111-146 and its docs), mapped onto ModelConfig fields
# ("cfg") or SAM2VideoPredictor keyword arguments ("pred").  hydra itself is not used.
_OVERRIDE_KEYS = {
    "model.fill_hole_area": ("cfg", "fill_hole_area", int),
    "model.non_overlap_masks": ("cfg", "non_overlap_masks", bool),
    "model.binarize_mask_from_pts_for_mem_enc": ("cfg", "binarize_mask_from_pts_for_mem_enc", bool),
    "model.sam_mask_decoder_extra_args.dynamic_multimask_via_stability": ("cfg", "dynamic_multimask_via_stability", bool),
    "model.sam_mask_decoder_extra_args.dynamic_multimask_stability_delta": ("cfg", "dynamic_multimask_stability_delta", float),
    "model.sam_mask_decoder_extra_args.dynamic_multimask_stability_thresh": ("cfg", "dynamic_multimask_stability_thresh", float),
    "model.image_size": ("cfg", "image_size", int),
    "model.num_maskmem": ("cfg", "num_maskmem", int),
    "model.max_cond_frames_in_attn": ("cfg", "max_cond_frames_in_attn", int),
    "model.max_obj_ptrs_in_encoder": ("cfg", "max_obj_ptrs_in_encoder", int),
    "model.multimask_min_pt_num": ("cfg", "multimask_min_pt_num", int),
    "model.multimask_max_pt_num": ("cfg", "multimask_max_pt_num", int),
    "model.clear_non_cond_mem_around_input": ("pred", "clear_non_cond_mem_around_input", bool),
    "model.clear_non_cond_mem_for_multi_obj": ("pred", "clear_non_cond_mem_for_multi_obj", bool),
    "model.add_all_frames_to_correct_as_cond": ("pred", "add_all_frames_to_correct_as_cond", bool),
}


def parse_hydra_overrides(overrides):
    """``["++model.fill_hole_area=0", ...]`` -> (ModelConfig overrides, predictor kwargs).  The reference appends the
    caller's ``hydra_overrides_extra`` to its own eval overrides (build_sam.py:123-137); the keys understood here are
    the model / predictor settings that exist on this path, anything else is refused by name."""
    cfg_over, pred_kw = {}, {}
    for item in overrides or ():
        if not isinstance(item, str) or "=" not in item:
            raise ValueError(f"malformed override {item!r}: expected '++model.<key>=<value>'")
        key, val = item.lstrip("+~").split("=", 1)
        if key not in _OVERRIDE_KEYS:
            raise ValueError(f"unsupported override {key!r}; supported: {sorted(_OVERRIDE_KEYS)}")
        where, name, typ = _OVERRIDE_KEYS[key]
        v = val.strip().strip("'\"")
        parsed = (v.lower() in ("true", "1", "yes")) if typ is bool else typ(float(v)) if typ is int else typ(v)
        (cfg_over if where == "cfg" else pred_kw)[name] = parsed
    return cfg_over, pred_kw


def build_sam2_video_predictor(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=(),
                               apply_postprocessing=True, state_dict=None, seed=0, engine=None, **kwargs):
    """Returns a detsam2_b200.predictor.SAM2VideoPredictor backed by the sm_100a CUDA engine.

    ``device`` must be a CUDA device: the product has no CPU compute path.  ``engine`` lets tests
    inject a checker implementation of the engine seams; production callers never pass it."""
    if mode != "eval":
        raise ValueError("only mode='eval' is supported (training is out of scope)")
    cfg_over, pred_kw = parse_hydra_overrides(hydra_overrides_extra)
    cfg_over.update({k: kwargs.pop(k) for k in list(kwargs) if k in get_config("tiny").__dataclass_fields__})
    for k, v in pred_kw.items():
        kwargs.setdefault(k, v)
    cfg = get_config(config_file, **cfg_over)
    if engine is None:
        if torch.device(device).type != "cuda":
            raise RuntimeError("detsam2_b200 runs the hot path on sm_100a CUDA kernels only; device must be 'cuda'")
        from .engine import CudaEngine
        sd = state_dict if state_dict is not None else load_state_dict(cfg, ckpt_path, seed)
        engine = CudaEngine(cfg, sd, device=device)
    fill = cfg.fill_hole_area if apply_postprocessing else 0
    return SAM2VideoPredictor(engine, fill_hole_area=fill, non_overlap_masks=cfg.non_overlap_masks, **kwargs)


class _ImageModel:
    """What build_sam2 returns: the handle SAM2ImagePredictor wraps (the reference returns the SAM2Base module)."""

    def __init__(self, engine):
        self.engine, self.cfg = engine, engine.cfg
        self.image_size, self.device = engine.cfg.image_size, engine.device


def build_sam2(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=(), apply_postprocessing=True,
               state_dict=None, seed=0, engine=None, **kwargs):
    """build_sam.py:69-108: the image model.  ``SAM2ImagePredictor(build_sam2(...).engine)`` or
    ``SAM2ImagePredictor.from_pretrained`` wrap it; same factory arguments as the video predictor."""
    pred = build_sam2_video_predictor(config_file, ckpt_path, device, mode, hydra_overrides_extra, apply_postprocessing,
                                      state_dict, seed, engine, **kwargs)
    return _ImageModel(pred.engine)


def build_sam2_hf(model_id, **kwargs):
    """build_sam.py:155-157."""
    config_name, ckpt_path = _hf_download(model_id)
    return build_sam2(config_file=config_name, ckpt_path=ckpt_path, **kwargs)


def _hf_download(model_id):
    """build_sam.py:148-153."""
    if model_id not in HF_MODEL_ID_TO_FILENAMES:
        raise KeyError(f"unknown model id {model_id!r}; supported: {sorted(HF_MODEL_ID_TO_FILENAMES)}")
    from huggingface_hub import hf_hub_download
    config_name, checkpoint_name = HF_MODEL_ID_TO_FILENAMES[model_id]
    return config_name, hf_hub_download(repo_id=model_id, filename=checkpoint_name)


def build_sam2_video_predictor_hf(model_id, **kwargs):
    """build_sam.py:159-163: checkpoint from the Hugging Face hub (or its local cache), then the factory above."""
    config_name, ckpt_path = _hf_download(model_id)
    return build_sam2_video_predictor(config_file=config_name, ckpt_path=ckpt_path, **kwargs)
