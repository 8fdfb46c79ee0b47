"""Parameter inventory of SAM 2.1 (names/shapes identical to the reference checkpoint layout,
``torch.load(ckpt)["model"]`` in /root/reference/sam2/build_sam.py:166-178) and a deterministic
synthetic initialiser — there are no checkpoints offline, so parity and benchmarks run on seeded
random weights that load *strictly* into both the reference modules and this engine.
"""
import math
from collections import OrderedDict

import torch

from .config import ModelConfig


def param_shapes(cfg: ModelConfig) -> "OrderedDict[str, tuple]":
    """Every state-dict key of the reference SAM2VideoPredictor for ``cfg`` with its shape."""
    s = OrderedDict()
    D, M = cfg.hidden_dim, cfg.mem_dim
    s["maskmem_tpos_enc"] = (cfg.num_maskmem, 1, 1, M)
    s["no_mem_embed"] = (1, 1, D)
    s["no_mem_pos_enc"] = (1, 1, D)
    s["no_obj_ptr"] = (1, D)
    s["no_obj_embed_spatial"] = (1, M)
    # ---- Hiera trunk (backbones/hieradet.py) ----
    t = "image_encoder.trunk."
    E = cfg.embed_dim
    s[t + "pos_embed"] = (1, E) + tuple(cfg.window_pos_embed_bkg_spatial_size)
    s[t + "pos_embed_window"] = (1, E, cfg.window_spec[0], cfg.window_spec[0])
    s[t + "patch_embed.proj.weight"] = (E, 3, 7, 7)
    s[t + "patch_embed.proj.bias"] = (E,)
    for i, b in enumerate(cfg.block_specs()):
        p = f"{t}blocks.{i}."
        d, do = b["dim"], b["dim_out"]
        s[p + "norm1.weight"] = (d,)
        s[p + "norm1.bias"] = (d,)
        s[p + "attn.qkv.weight"] = (3 * do, d)
        s[p + "attn.qkv.bias"] = (3 * do,)
        s[p + "attn.proj.weight"] = (do, do)
        s[p + "attn.proj.bias"] = (do,)
        s[p + "norm2.weight"] = (do,)
        s[p + "norm2.bias"] = (do,)
        s[p + "mlp.layers.0.weight"] = (4 * do, do)
        s[p + "mlp.layers.0.bias"] = (4 * do,)
        s[p + "mlp.layers.1.weight"] = (do, 4 * do)
        s[p + "mlp.layers.1.bias"] = (do,)
        if d != do:
            s[p + "proj.weight"] = (do, d)
            s[p + "proj.bias"] = (do,)
    # ---- FPN neck (backbones/image_encoder.py) ----
    for i, c in enumerate(cfg.backbone_channel_list):
        s[f"image_encoder.neck.convs.{i}.conv.weight"] = (D, c, 1, 1)
        s[f"image_encoder.neck.convs.{i}.conv.bias"] = (D,)
    s["mask_downsample.weight"] = (1, 1, 4, 4)
    s["mask_downsample.bias"] = (1,)
    # ---- memory attention ----
    for l in range(cfg.memattn_layers):
        p = f"memory_attention.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (D, D)
            s[p + f"self_attn.{n}.bias"] = (D,)
        s[p + "cross_attn_image.q_proj.weight"] = (D, D)
        s[p + "cross_attn_image.q_proj.bias"] = (D,)
        s[p + "cross_attn_image.k_proj.weight"] = (D, M)
        s[p + "cross_attn_image.k_proj.bias"] = (D,)
        s[p + "cross_attn_image.v_proj.weight"] = (D, M)
        s[p + "cross_attn_image.v_proj.bias"] = (D,)
        s[p + "cross_attn_image.out_proj.weight"] = (D, D)
        s[p + "cross_attn_image.out_proj.bias"] = (D,)
        s[p + "linear1.weight"] = (cfg.memattn_ffn, D)
        s[p + "linear1.bias"] = (cfg.memattn_ffn,)
        s[p + "linear2.weight"] = (D, cfg.memattn_ffn)
        s[p + "linear2.bias"] = (D,)
        for n in ("norm1", "norm2", "norm3"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
    s["memory_attention.norm.weight"] = (D,)
    s["memory_attention.norm.bias"] = (D,)
    # ---- memory encoder ----
    p = "memory_encoder.mask_downsampler.encoder."
    cin, idx = 1, 0
    for _ in range(4):
        cout = cin * 4
        s[p + f"{idx}.weight"] = (cout, cin, 3, 3)
        s[p + f"{idx}.bias"] = (cout,)
        s[p + f"{idx + 1}.weight"] = (cout,)
        s[p + f"{idx + 1}.bias"] = (cout,)
        cin, idx = cout, idx + 3
    s[p + f"{idx}.weight"] = (D, cin, 1, 1)
    s[p + f"{idx}.bias"] = (D,)
    s["memory_encoder.pix_feat_proj.weight"] = (D, D, 1, 1)
    s["memory_encoder.pix_feat_proj.bias"] = (D,)
    for l in range(2):
        p = f"memory_encoder.fuser.layers.{l}."
        s[p + "gamma"] = (D,)
        s[p + "dwconv.weight"] = (D, 1, 7, 7)
        s[p + "dwconv.bias"] = (D,)
        s[p + "norm.weight"] = (D,)
        s[p + "norm.bias"] = (D,)
        s[p + "pwconv1.weight"] = (4 * D, D)
        s[p + "pwconv1.bias"] = (4 * D,)
        s[p + "pwconv2.weight"] = (D, 4 * D)
        s[p + "pwconv2.bias"] = (D,)
    s["memory_encoder.out_proj.weight"] = (M, D, 1, 1)
    s["memory_encoder.out_proj.bias"] = (M,)
    # ---- prompt encoder ----
    p = "sam_prompt_encoder."
    s[p + "pe_layer.positional_encoding_gaussian_matrix"] = (2, D // 2)
    for i in range(4):
        s[p + f"point_embeddings.{i}.weight"] = (1, D)
    s[p + "not_a_point_embed.weight"] = (1, D)
    s[p + "mask_downscaling.0.weight"] = (4, 1, 2, 2)
    s[p + "mask_downscaling.0.bias"] = (4,)
    s[p + "mask_downscaling.1.weight"] = (4,)
    s[p + "mask_downscaling.1.bias"] = (4,)
    s[p + "mask_downscaling.3.weight"] = (16, 4, 2, 2)
    s[p + "mask_downscaling.3.bias"] = (16,)
    s[p + "mask_downscaling.4.weight"] = (16,)
    s[p + "mask_downscaling.4.bias"] = (16,)
    s[p + "mask_downscaling.6.weight"] = (D, 16, 1, 1)
    s[p + "mask_downscaling.6.bias"] = (D,)
    s[p + "no_mask_embed.weight"] = (1, D)
    # ---- mask decoder ----
    p = "sam_mask_decoder."
    Dh = D // 2  # attention_downsample_rate = 2

    def attn(prefix, inner):
        for n in ("q_proj", "k_proj", "v_proj"):
            s[prefix + n + ".weight"] = (inner, D)
            s[prefix + n + ".bias"] = (inner,)
        s[prefix + "out_proj.weight"] = (D, inner)
        s[prefix + "out_proj.bias"] = (D,)

    for l in range(cfg.decoder_depth):
        q = f"{p}transformer.layers.{l}."
        attn(q + "self_attn.", D)
        s[q + "norm1.weight"] = (D,)
        s[q + "norm1.bias"] = (D,)
        attn(q + "cross_attn_token_to_image.", Dh)
        s[q + "norm2.weight"] = (D,)
        s[q + "norm2.bias"] = (D,)
        s[q + "mlp.layers.0.weight"] = (cfg.decoder_mlp, D)
        s[q + "mlp.layers.0.bias"] = (cfg.decoder_mlp,)
        s[q + "mlp.layers.1.weight"] = (D, cfg.decoder_mlp)
        s[q + "mlp.layers.1.bias"] = (D,)
        s[q + "norm3.weight"] = (D,)
        s[q + "norm3.bias"] = (D,)
        s[q + "norm4.weight"] = (D,)
        s[q + "norm4.bias"] = (D,)
        attn(q + "cross_attn_image_to_token.", Dh)
    attn(p + "transformer.final_attn_token_to_image.", Dh)
    s[p + "transformer.norm_final_attn.weight"] = (D,)
    s[p + "transformer.norm_final_attn.bias"] = (D,)
    s[p + "iou_token.weight"] = (1, D)
    s[p + "mask_tokens.weight"] = (cfg.num_multimask_outputs + 1, D)
    s[p + "obj_score_token.weight"] = (1, D)
    s[p + "output_upscaling.0.weight"] = (D, D // 4, 2, 2)
    s[p + "output_upscaling.0.bias"] = (D // 4,)
    s[p + "output_upscaling.1.weight"] = (D // 4,)
    s[p + "output_upscaling.1.bias"] = (D // 4,)
    s[p + "output_upscaling.3.weight"] = (D // 4, D // 8, 2, 2)
    s[p + "output_upscaling.3.bias"] = (D // 8,)
    s[p + "conv_s0.weight"] = (D // 8, D, 1, 1)
    s[p + "conv_s0.bias"] = (D // 8,)
    s[p + "conv_s1.weight"] = (D // 4, D, 1, 1)
    s[p + "conv_s1.bias"] = (D // 4,)

    def mlp3(prefix, dout):
        s[prefix + "layers.0.weight"] = (D, D)
        s[prefix + "layers.0.bias"] = (D,)
        s[prefix + "layers.1.weight"] = (D, D)
        s[prefix + "layers.1.bias"] = (D,)
        s[prefix + "layers.2.weight"] = (dout, D)
        s[prefix + "layers.2.bias"] = (dout,)

    for i in range(cfg.num_multimask_outputs + 1):
        mlp3(f"{p}output_hypernetworks_mlps.{i}.", D // 8)
    mlp3(p + "iou_prediction_head.", cfg.num_multimask_outputs + 1)
    mlp3(p + "pred_obj_score_head.", 1)
    mlp3("obj_ptr_proj.", D)
    s["obj_ptr_tpos_proj.weight"] = (M, D)
    s["obj_ptr_tpos_proj.bias"] = (M,)
    return s


def synthetic_state_dict(cfg: ModelConfig, seed: int = 0, obj_score_bias: float = 4.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random weights (fp32, CPU).  Scales are chosen so that activations stay O(1) through
    the 48-block trunk and mask logits are O(1..10): weights ~ N(0, 1/fan_in), norm scales ~ 1,
    biases / embeddings small, layer-scale gamma O(0.1) so the ConvNeXt branch matters."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = (".norm" in name and "norm_final" not in name and len(shape) == 1) or \
            "norm_final_attn" in name or \
            (len(shape) == 1 and ("mask_downsampler.encoder" in name or "mask_downscaling" in name
                                  or "output_upscaling.1" in name)
             and _is_ln2d(name))
        if name.endswith("gamma"):
            t = 0.05 + 0.25 * torch.rand(shape, generator=g)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("positional_encoding_gaussian_matrix"):
            t = torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) >= 2 and "embed" not in name and "token" not in name:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if "output_upscaling" in name and len(shape) == 4:  # ConvTranspose2d: [in, out, kh, kw]
                fan_in = shape[0]
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:  # embeddings, tokens, positional tables
            t = 0.5 * torch.randn(shape, generator=g) if ("token" in name or "point_embeddings" in name) \
                else 0.1 * torch.randn(shape, generator=g)
        sd[name] = t.float().contiguous()
    # random heads put the object score near 0, where the `score > 0` gate (sam2_base.py:342-350)
    # flips on rounding noise; a trained checkpoint is decisively positive on visible objects
    sd["sam_mask_decoder.pred_obj_score_head.layers.2.bias"] += obj_score_bias
    return sd


def _is_ln2d(name):
    # LayerNorm2d members inside nn.Sequential containers sit at indices 1, 4, 7, 10
    idx = name.rsplit(".", 2)[-2]
    return idx in ("1", "4", "7", "10")
