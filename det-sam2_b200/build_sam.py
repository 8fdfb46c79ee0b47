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


def build_sam2_video_predictor(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=(),
                               apply_postprocessing=True, state_dict=None, seed=0, engine=None, **kwargs):
    """Returns a detsam2_b200.predictor.SAM2VideoPredictor backed by the sm_100a CUDA engine.

    ``device`` must be a CUDA device: the product has no CPU compute path.  ``engine`` lets tests
    inject a checker implementation of the engine seams; production callers never pass it."""
    if mode != "eval":
        raise ValueError("only mode='eval' is supported (training is out of scope)")
    if hydra_overrides_extra:
        raise ValueError("hydra overrides are not supported; pass ModelConfig field overrides as keyword arguments")
    cfg_over = {k: kwargs.pop(k) for k in list(kwargs) if k in get_config("tiny").__dataclass_fields__}
    cfg = get_config(config_file, **cfg_over)
    if engine is None:
        if torch.device(device).type != "cuda":
            raise RuntimeError("detsam2_b200 runs the hot path on sm_100a CUDA kernels only; device must be 'cuda'")
        from .engine import CudaEngine
        sd = state_dict if state_dict is not None else load_state_dict(cfg, ckpt_path, seed)
        engine = CudaEngine(cfg, sd, device=device)
    fill = cfg.fill_hole_area if apply_postprocessing else 0
    return SAM2VideoPredictor(engine, fill_hole_area=fill, non_overlap_masks=cfg.non_overlap_masks, **kwargs)
