"""Model hyper-parameters of the SAM 2.1 family, restating the reference's hydra YAML
(/root/reference/sam2/configs/sam2.1/sam2.1_hiera_{t,s,b+,l}.yaml) and the SAM2Base defaults
(/root/reference/sam2/modeling/sam2_base.py:25-98) plus the eval-time overrides applied by
build_sam2_video_predictor (/root/reference/sam2/build_sam.py:121-141).
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class ModelConfig:
    name: str
    # Hiera trunk (backbones/hieradet.py:171-204)
    embed_dim: int
    num_heads: int
    stages: Tuple[int, ...]
    global_att_blocks: Tuple[int, ...]
    window_spec: Tuple[int, ...]
    window_pos_embed_bkg_spatial_size: Tuple[int, int]
    backbone_channel_list: Tuple[int, ...]
    q_pool: int = 3
    # common
    image_size: int = 1024
    backbone_stride: int = 16
    hidden_dim: int = 256
    mem_dim: int = 64
    num_maskmem: int = 7
    max_cond_frames_in_attn: int = 20
    max_obj_ptrs_in_encoder: int = 16
    sigmoid_scale_for_mem_enc: float = 20.0
    sigmoid_bias_for_mem_enc: float = -10.0
    memattn_layers: int = 4
    memattn_ffn: int = 2048
    rope_theta: float = 10000.0
    decoder_depth: int = 2
    decoder_heads: int = 8
    decoder_mlp: int = 2048
    num_multimask_outputs: int = 3
    # eval overrides (build_sam.py:126-135)
    dynamic_multimask_via_stability: bool = True
    dynamic_multimask_stability_delta: float = 0.05
    dynamic_multimask_stability_thresh: float = 0.98
    binarize_mask_from_pts_for_mem_enc: bool = True
    fill_hole_area: int = 8
    non_overlap_masks: bool = False
    multimask_min_pt_num: int = 0
    multimask_max_pt_num: int = 1

    @property
    def feat_size(self):
        return self.image_size // self.backbone_stride

    @property
    def depth(self):
        return sum(self.stages)

    @property
    def stage_ends(self):
        return [sum(self.stages[:i]) - 1 for i in range(1, len(self.stages) + 1)]

    @property
    def q_pool_blocks(self):
        return [x + 1 for x in self.stage_ends[:-1]][: self.q_pool]

    def block_specs(self):
        """Per-block (dim, dim_out, heads, window, q_pool?) following hieradet.py:238-263."""
        specs = []
        embed_dim, heads, cur_stage = self.embed_dim, self.num_heads, 1
        for i in range(self.depth):
            dim_out = embed_dim
            window = self.window_spec[cur_stage - 1]
            if i in self.global_att_blocks:
                window = 0
            if i - 1 in self.stage_ends:
                dim_out = embed_dim * 2
                heads = heads * 2
                cur_stage += 1
            specs.append(dict(dim=embed_dim, dim_out=dim_out, heads=heads, window=window,
                              q_pool=i in self.q_pool_blocks))
            embed_dim = dim_out
        return specs


_CONFIGS = {
    "tiny": ModelConfig(
        name="tiny", embed_dim=96, num_heads=1, stages=(1, 2, 7, 2), global_att_blocks=(5, 7, 9),
        window_spec=(8, 4, 14, 7), window_pos_embed_bkg_spatial_size=(7, 7),
        backbone_channel_list=(768, 384, 192, 96)),
    "small": ModelConfig(
        name="small", embed_dim=96, num_heads=1, stages=(1, 2, 11, 2), global_att_blocks=(7, 10, 13),
        window_spec=(8, 4, 14, 7), window_pos_embed_bkg_spatial_size=(7, 7),
        backbone_channel_list=(768, 384, 192, 96)),
    "base_plus": ModelConfig(
        name="base_plus", embed_dim=112, num_heads=2, stages=(2, 3, 16, 3), global_att_blocks=(12, 16, 20),
        window_spec=(8, 4, 14, 7), window_pos_embed_bkg_spatial_size=(14, 14),
        backbone_channel_list=(896, 448, 224, 112)),
    "large": ModelConfig(
        name="large", embed_dim=144, num_heads=2, stages=(2, 6, 36, 4), global_att_blocks=(23, 33, 43),
        window_spec=(8, 4, 16, 8), window_pos_embed_bkg_spatial_size=(7, 7),
        backbone_channel_list=(1152, 576, 288, 144)),
}

_YAML_ALIASES = {
    "sam2.1_hiera_t": "tiny", "sam2.1_hiera_s": "small", "sam2.1_hiera_b+": "base_plus",
    "sam2.1_hiera_l": "large",
}


def get_config(name_or_yaml: str, **overrides) -> ModelConfig:
    """Accepts 'tiny'/'small'/'base_plus'/'large' or a reference config path such as
    'configs/sam2.1/sam2.1_hiera_l.yaml' (build_sam.py:111-121)."""
    key = name_or_yaml
    if key not in _CONFIGS:
        import os
        stem = os.path.splitext(os.path.basename(name_or_yaml))[0]
        if stem not in _YAML_ALIASES:
            raise ValueError(f"unknown SAM 2.1 config {name_or_yaml!r}")
        key = _YAML_ALIASES[stem]
    cfg = _CONFIGS[key]
    if overrides:
        from dataclasses import replace
        cfg = replace(cfg, **overrides)
    return cfg
