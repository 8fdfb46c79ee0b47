"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch, fp32) of the reference's SAM 2.1 video hot
path.  Nothing in the product imports this; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may.

It restates, function by function, what /root/reference does between a normalised frame and the
per-object outputs of one tracking step, operating directly on a reference-layout state dict
(``torch.load(ckpt)["model"]`` keys).  Every function cites the reference file:line it follows.

Parity pinning: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through
oracle/ref_shim.py: tests/test_oracle_vs_reference.py (live, skipped where /root/reference is
absent) and the committed fixtures under tests/golden/ produced by oracle/gen_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

NO_OBJ_SCORE = -1024.0  # sam2/modeling/sam2_base.py:17


# ------------------------------------------------------------------------------------------------
# small building blocks
# ------------------------------------------------------------------------------------------------
def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def layer_norm(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm_2d(sd, p, x, eps=1e-6):
    """sam2/modeling/sam2_utils.py:150-162 (biased variance over C of NCHW)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][:, None, None] * x + sd[p + ".bias"][:, None, None]


def mlp(sd, p, x, n, act=F.relu, sigmoid_out=False):
    """sam2/modeling/sam2_utils.py:121-145."""
    for i in range(n):
        x = linear(sd, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = act(x)
    return torch.sigmoid(x) if sigmoid_out else x


# False: attention written out (scores materialised, fp32) — the checker.  True: F.scaled_dot_product_attention, what the
# reference calls (sam/transformer.py:268,347; flash backend under CUDA bf16 autocast) — used by the GPU library baseline.
FUSED_SDPA = False


def sdpa(q, k, v):
    """F.scaled_dot_product_attention with no mask / dropout, written out (fp32) unless FUSED_SDPA."""
    if FUSED_SDPA:
        return F.scaled_dot_product_attention(q, k, v)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(q.shape[-1]))
    return torch.softmax(s, dim=-1) @ v


# ------------------------------------------------------------------------------------------------
# frame ingest — sam2/utils/misc.py:236-363 (ndarray-list branch) + sam2_video_predictor.py:1184-1186
# ------------------------------------------------------------------------------------------------
IMG_MEAN = (0.485, 0.456, 0.406)
IMG_STD = (0.229, 0.224, 0.225)


def load_frames(frames_rgb_u8, image_size):
    """list of HxWx3 uint8 RGB -> fp16 [N,3,S,S] normalised exactly as misc.py:327-359 does
    (cv2.resize, /255 in float64, store fp16, then in-place fp16 -= mean, /= std)."""
    import cv2
    imgs = torch.zeros(len(frames_rgb_u8), 3, image_size, image_size, dtype=torch.float16)
    for n, fr in enumerate(frames_rgb_u8):
        a = cv2.resize(fr, (image_size, image_size)) / 255.0
        imgs[n] = torch.from_numpy(a).permute(2, 0, 1)
    imgs -= torch.tensor(IMG_MEAN, dtype=torch.float32)[:, None, None]
    imgs /= torch.tensor(IMG_STD, dtype=torch.float32)[:, None, None]
    return imgs, frames_rgb_u8[0].shape[0], frames_rgb_u8[0].shape[1]


# ------------------------------------------------------------------------------------------------
# positional encodings — sam2/modeling/position_encoding.py
# ------------------------------------------------------------------------------------------------
def sine_pe_2d(num_pos_feats_total, h, w, temperature=10000.0):
    """PositionEmbeddingSine.forward (:78-112), normalize=True, scale=2*pi -> [C,h,w]."""
    npf = num_pos_feats_total // 2
    y = torch.arange(1, h + 1, dtype=torch.float32).view(-1, 1).repeat(1, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, -1).repeat(h, 1)
    eps, scale = 1e-6, 2 * math.pi
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1).contiguous()


def sine_pe_1d(pos, dim, temperature=10000.0):
    """get_1d_sine_pe, sam2/modeling/sam2_utils.py:69-79."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    e = pos.unsqueeze(-1) / dim_t
    return torch.cat([e.sin(), e.cos()], dim=-1)


def axial_rope_table(dim, end_x, end_y, theta=10000.0):
    """compute_axial_cis (:173-182) as (cos, sin) [end_x*end_y, dim/2]."""
    fr = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx = (t % end_x).float()
    ty = torch.div(t, end_x, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(tx, fr), torch.outer(ty, fr)], dim=-1)
    return ang.cos(), ang.sin()


def apply_rope(x, cos, sin):
    """apply_rotary_enc (:193-220) on real pairs (2i, 2i+1); x [..., L, D], tables [L, D/2]."""
    xr = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xr[..., 0], xr[..., 1]
    out = torch.stack([a * cos - b * sin, a * sin + b * cos], dim=-1)
    return out.flatten(-2).type_as(x)


# ------------------------------------------------------------------------------------------------
# Hiera trunk + FPN neck — sam2/modeling/backbones/{hieradet,image_encoder,utils}.py
# ------------------------------------------------------------------------------------------------
def window_partition(x, ws):
    """backbones/utils.py:16-38."""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C), (Hp, Wp)


def window_unpartition(win, ws, pad_hw, hw):
    """backbones/utils.py:41-63."""
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // ws // ws)
    x = win.reshape(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def maxpool2x2_nhwc(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), kernel_size=2, stride=2).permute(0, 2, 3, 1)


def hiera_block(sd, p, x, spec):
    """MultiScaleBlock.forward (hieradet.py:136-168) + MultiScaleAttention.forward (:57-82)."""
    dim, dim_out, heads, ws, pool = spec["dim"], spec["dim_out"], spec["heads"], spec["window"], spec["q_pool"]
    shortcut = x
    x = layer_norm(sd, p + "norm1", x, 1e-6)
    if dim != dim_out:
        shortcut = linear(sd, p + "proj", x)
        if pool:
            shortcut = maxpool2x2_nhwc(shortcut)
    H, W = x.shape[1], x.shape[2]
    pad_hw = (H, W)
    if ws > 0:
        x, pad_hw = window_partition(x, ws)
    Bw, h, w, _ = x.shape
    qkv = linear(sd, p + "attn.qkv", x).reshape(Bw, h * w, 3, heads, -1)
    q, k, v = torch.unbind(qkv, 2)
    if pool:
        q = maxpool2x2_nhwc(q.reshape(Bw, h, w, -1))
        h, w = q.shape[1:3]
        q = q.reshape(Bw, h * w, heads, -1)
    o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(Bw, h, w, -1)
    x = linear(sd, p + "attn.proj", o)
    if pool:
        ws = ws // 2
        H, W = shortcut.shape[1:3]
        pad_hw = (H + (ws - H % ws) % ws, W + (ws - W % ws) % ws) if ws > 0 else (H, W)
    if spec["window"] > 0:
        x = window_unpartition(x, ws, pad_hw, (H, W))
    x = shortcut + x
    y = layer_norm(sd, p + "norm2", x, 1e-6)
    y = linear(sd, p + "mlp.layers.1", F.gelu(linear(sd, p + "mlp.layers.0", y)))
    return x + y


def hiera_pos_embed(sd, h, w):
    """Hiera._get_pos_embed (hieradet.py:273-281) -> [1,h,w,E] (input independent)."""
    t = "image_encoder.trunk."
    pe = F.interpolate(sd[t + "pos_embed"], size=(h, w), mode="bicubic")
    we = sd[t + "pos_embed_window"]
    pe = pe + we.tile([x // y for x, y in zip(pe.shape, we.shape)])
    return pe.permute(0, 2, 3, 1)


def hiera_trunk(sd, cfg, img):
    """Hiera.forward (hieradet.py:283-299): img [1,3,S,S] fp32 -> 4 NCHW stage outputs."""
    t = "image_encoder.trunk."
    x = F.conv2d(img, sd[t + "patch_embed.proj.weight"], sd[t + "patch_embed.proj.bias"], stride=4, padding=3)
    x = x.permute(0, 2, 3, 1)
    x = x + hiera_pos_embed(sd, x.shape[1], x.shape[2])
    outs = []
    ends = cfg.stage_ends
    for i, spec in enumerate(cfg.block_specs()):
        x = hiera_block(sd, f"{t}blocks.{i}.", x, spec)
        if i in ends:
            outs.append(x.permute(0, 3, 1, 2))
    return outs


def forward_image(sd, cfg, img):
    """SAM2Base.forward_image (sam2_base.py:450-461) = ImageEncoder.forward (image_encoder.py:30-43)
    + FpnNeck.forward (:101-134, top-down only into level 2, nearest) + conv_s0 / conv_s1.
    Returns dict(backbone_fpn=[f256(32ch), f128(64ch), f64(256ch)], vision_pos_enc=[...3])."""
    xs = hiera_trunk(sd, cfg, img.float())
    n = len(xs) - 1
    out, prev = [None] * len(xs), None
    for i in range(n, -1, -1):
        c = f"image_encoder.neck.convs.{n - i}.conv"
        lat = F.conv2d(xs[i], sd[c + ".weight"], sd[c + ".bias"])
        if i in (2, 3) and prev is not None:
            prev = lat + F.interpolate(prev.float(), scale_factor=2.0, mode="nearest")
        else:
            prev = lat
        out[i] = prev
    feats = out[:-1]  # scalp = 1
    pos = [sine_pe_2d(256, f.shape[-2], f.shape[-1])[None] for f in feats]
    d = "sam_mask_decoder."
    feats[0] = F.conv2d(feats[0], sd[d + "conv_s0.weight"], sd[d + "conv_s0.bias"])
    feats[1] = F.conv2d(feats[1], sd[d + "conv_s1.weight"], sd[d + "conv_s1.bias"])
    return {"backbone_fpn": feats, "vision_pos_enc": pos}


# ------------------------------------------------------------------------------------------------
# memory attention — sam2/modeling/memory_attention.py + sam/transformer.py:286-363
# ------------------------------------------------------------------------------------------------
def rope_attention(sd, p, q, k, v, rope, num_k_exclude_rope=0, repeat_k=False):
    """RoPEAttention.forward (transformer.py:311-363), one head; q [B,Lq,256], k/v [B,Lk,kv_dim]."""
    q, k, v = linear(sd, p + ".q_proj", q), linear(sd, p + ".k_proj", k), linear(sd, p + ".v_proj", v)
    cos, sin = rope
    q = apply_rope(q, cos, sin)
    nk = k.shape[1] - num_k_exclude_rope
    if nk > 0:
        r = nk // q.shape[1] if repeat_k else 1
        kc, ks = (cos.repeat(r, 1), sin.repeat(r, 1)) if repeat_k else (cos, sin)
        k = torch.cat([apply_rope(k[:, :nk], kc, ks), k[:, nk:]], dim=1)
    return linear(sd, p + ".out_proj", sdpa(q, k, v))


def memory_attention(sd, cfg, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens):
    """MemoryAttention.forward (memory_attention.py:119-176) + MemoryAttentionLayer.forward (:83-99).
    curr, curr_pos [T,B,256]; memory, memory_pos [N,B,64] (sequence first) -> [T,B,256]."""
    x = (curr + 0.1 * curr_pos).transpose(0, 1)
    mem, mpos = memory.transpose(0, 1), memory_pos.transpose(0, 1)
    side = int(math.isqrt(x.shape[1]))
    rope = axial_rope_table(256, side, side, cfg.rope_theta)
    for l in range(cfg.memattn_layers):
        p = f"memory_attention.layers.{l}."
        t = layer_norm(sd, p + "norm1", x)
        x = x + rope_attention(sd, p + "self_attn", t, t, t, rope)
        t = layer_norm(sd, p + "norm2", x)
        x = x + rope_attention(sd, p + "cross_attn_image", t, mem + mpos, mem, rope,
                               num_k_exclude_rope=num_obj_ptr_tokens, repeat_k=True)
        t = layer_norm(sd, p + "norm3", x)
        x = x + linear(sd, p + "linear2", F.relu(linear(sd, p + "linear1", t)))
    return layer_norm(sd, "memory_attention.norm", x).transpose(0, 1)


def select_closest_cond_frames(frame_idx, cond, max_num, preload_idx=None):
    """sam2/modeling/sam2_utils.py:19-66 (with Det-SAM2's preload forcing, :56-60)."""
    if max_num == -1 or len(cond) <= max_num:
        return cond, {}
    sel = {}
    before = max((t for t in cond if t < frame_idx), default=None)
    if before is not None:
        sel[before] = cond[before]
    after = min((t for t in cond if t >= frame_idx), default=None)
    if after is not None:
        sel[after] = cond[after]
    remain = sorted((t for t in cond if t not in sel), key=lambda t: abs(t - frame_idx))[: max_num - len(sel)]
    sel.update((t, cond[t]) for t in remain)
    if preload_idx is not None:
        for t in preload_idx:
            if t not in sel:
                sel[t] = cond[t]
    return sel, {t: v for t, v in cond.items() if t not in sel}


def prepare_memory_conditioned_features(sd, cfg, frame_idx, is_init_cond_frame, vision_feat, vision_pos,
                                        output_dict, num_frames, reverse=False, preload_idx=None):
    """SAM2Base._prepare_memory_conditioned_features (sam2_base.py:479-690).
    vision_feat / vision_pos [T,B,256]; entries of output_dict hold NCHW maskmem_features (bf16),
    maskmem_pos_enc [.., [B,64,h,w]], obj_ptr [B,256].  Returns [B,256,h,w]."""
    T, B, C = vision_feat.shape
    side = int(math.isqrt(T))
    if is_init_cond_frame:
        return (vision_feat + sd["no_mem_embed"]).permute(1, 2, 0).reshape(B, C, side, side)
    nm = cfg.num_maskmem
    sel, unsel = select_closest_cond_frames(frame_idx, output_dict["cond_frame_outputs"],
                                            cfg.max_cond_frames_in_attn, preload_idx)
    t_pos_and_prevs = [(0, o) for o in sel.values()]
    for t_pos in range(1, nm):
        t_rel = nm - t_pos
        prev_idx = frame_idx + t_rel if reverse else frame_idx - t_rel  # stride 1 (eval default)
        o = output_dict["non_cond_frame_outputs"].get(prev_idx, None)
        if o is None:
            o = unsel.get(prev_idx, None)
        t_pos_and_prevs.append((t_pos, o))
    mems, poss = [], []
    for t_pos, prev in t_pos_and_prevs:
        if prev is None:
            continue
        f = prev["maskmem_features"].to(torch.float32)
        mems.append(f.flatten(2).permute(2, 0, 1))
        pe = prev["maskmem_pos_enc"][-1].float().flatten(2).permute(2, 0, 1)
        poss.append(pe + sd["maskmem_tpos_enc"][nm - t_pos - 1])
    max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
    ptr_cond = {t: o for t, o in sel.items() if (t >= frame_idx if reverse else t <= frame_idx)}
    sign = -1 if reverse else 1
    pos_and_ptrs = [((frame_idx - t) * sign, o["obj_ptr"]) for t, o in ptr_cond.items()]
    for t_diff in range(1, max_ptrs):
        t = frame_idx + t_diff if reverse else frame_idx - t_diff
        if t < 0 or (num_frames is not None and t >= num_frames):
            break
        o = output_dict["non_cond_frame_outputs"].get(t, unsel.get(t, None))
        if o is not None:
            pos_and_ptrs.append((t_diff, o["obj_ptr"]))
    n_ptr_tokens = 0
    if pos_and_ptrs:
        pos_list, ptrs = zip(*pos_and_ptrs)
        obj_ptrs = torch.stack([p.float() for p in ptrs], dim=0)
        obj_pos = sine_pe_1d(torch.tensor(pos_list, dtype=torch.float32) / (max_ptrs - 1), C)
        obj_pos = linear(sd, "obj_ptr_tpos_proj", obj_pos).unsqueeze(1).expand(-1, B, cfg.mem_dim)
        r = C // cfg.mem_dim
        obj_ptrs = obj_ptrs.reshape(-1, B, r, cfg.mem_dim).permute(0, 2, 1, 3).flatten(0, 1)
        obj_pos = obj_pos.repeat_interleave(r, dim=0)
        mems.append(obj_ptrs)
        poss.append(obj_pos)
        n_ptr_tokens = obj_ptrs.shape[0]
    memory, memory_pos = torch.cat(mems, 0), torch.cat(poss, 0)
    out = memory_attention(sd, cfg, vision_feat, vision_pos, memory, memory_pos, n_ptr_tokens)
    return out.permute(1, 2, 0).reshape(B, C, side, side)


# ------------------------------------------------------------------------------------------------
# prompt encoder + mask decoder — sam2/modeling/sam/{prompt_encoder,mask_decoder,transformer}.py
# ------------------------------------------------------------------------------------------------
def random_pe(sd, coords01):
    """PositionEmbeddingRandom._pe_encoding (position_encoding.py:131-136)."""
    g = sd["sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    c = (2 * coords01 - 1) @ g
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd, size):
    """PromptEncoder.get_dense_pe (prompt_encoder.py:64-71) -> [1,256,h,w]."""
    h = w = size
    g = torch.ones(h, w, dtype=torch.float32)
    y = (g.cumsum(0) - 0.5) / h
    x = (g.cumsum(1) - 0.5) / w
    return random_pe(sd, torch.stack([x, y], dim=-1)).permute(2, 0, 1)[None]


def prompt_encoder(sd, cfg, coords, labels, mask_prompt=None):
    """PromptEncoder.forward (prompt_encoder.py:134-171), boxes=None (pad point appended)."""
    p = "sam_prompt_encoder."
    B = coords.shape[0]
    pts = coords.float() + 0.5
    pts = torch.cat([pts, torch.zeros(B, 1, 2)], dim=1)
    lab = torch.cat([labels, -torch.ones(B, 1, dtype=labels.dtype)], dim=1)
    e = random_pe(sd, pts / cfg.image_size)
    e = torch.where((lab == -1)[..., None], torch.zeros_like(e), e)
    e = e + (lab == -1)[..., None] * sd[p + "not_a_point_embed.weight"]
    for i in range(4):
        e = e + (lab == i)[..., None] * sd[p + f"point_embeddings.{i}.weight"]
    fs = cfg.feat_size
    if mask_prompt is None:
        dense = sd[p + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, fs, fs)
    else:
        m = p + "mask_downscaling."
        x = F.conv2d(mask_prompt.float(), sd[m + "0.weight"], sd[m + "0.bias"], stride=2)
        x = F.gelu(layer_norm_2d(sd, m + "1", x))
        x = F.conv2d(x, sd[m + "3.weight"], sd[m + "3.bias"], stride=2)
        x = F.gelu(layer_norm_2d(sd, m + "4", x))
        dense = F.conv2d(x, sd[m + "6.weight"], sd[m + "6.bias"])
    return e, dense


def attention(sd, p, q, k, v, heads):
    """Attention.forward (transformer.py:253-284)."""
    q, k, v = linear(sd, p + ".q_proj", q), linear(sd, p + ".k_proj", k), linear(sd, p + ".v_proj", v)

    def split(x):
        b, n, c = x.shape
        return x.reshape(b, n, heads, c // heads).transpose(1, 2)

    o = sdpa(split(q), split(k), split(v)).transpose(1, 2)
    o = o.reshape(o.shape[0], o.shape[1], -1)
    return linear(sd, p + ".out_proj", o)


def two_way_transformer(sd, cfg, src, pos_src, tokens):
    """TwoWayTransformer.forward (transformer.py:90-133) + TwoWayAttentionBlock.forward (:178-211)."""
    t = "sam_mask_decoder.transformer."
    B, C, H, W = src.shape
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos_src.flatten(2).permute(0, 2, 1)
    queries, qpe, hds = tokens, tokens, cfg.decoder_heads
    for l in range(cfg.decoder_depth):
        p = f"{t}layers.{l}."
        if l == 0:
            queries = attention(sd, p + "self_attn", queries, queries, queries, hds)
        else:
            q = queries + qpe
            queries = queries + attention(sd, p + "self_attn", q, q, queries, hds)
        queries = layer_norm(sd, p + "norm1", queries)
        q, k = queries + qpe, keys + kpe
        queries = layer_norm(sd, p + "norm2", queries + attention(sd, p + "cross_attn_token_to_image", q, k, keys, hds))
        m = linear(sd, p + "mlp.layers.1", F.relu(linear(sd, p + "mlp.layers.0", queries)))
        queries = layer_norm(sd, p + "norm3", queries + m)
        q, k = queries + qpe, keys + kpe
        keys = layer_norm(sd, p + "norm4", keys + attention(sd, p + "cross_attn_image_to_token", k, q, queries, hds))
    q, k = queries + qpe, keys + kpe
    queries = layer_norm(sd, t + "norm_final_attn",
                         queries + attention(sd, t + "final_attn_token_to_image", q, k, keys, hds))
    return queries, keys


# Tests set this to a list to record the two data-dependent DISCRETE choices of the prompted decode
# (stability fallback, mask_decoder.py:261-296, and the argmax-IoU pick): with seeded random weights
# they can be near ties, and a bf16 implementation may then legitimately choose the other mask.
DECISION_LOG = None


def mask_decoder(sd, cfg, image_embeddings, image_pe, sparse, dense, multimask_output, high_res_features):
    """MaskDecoder.forward / predict_masks (mask_decoder.py:105-247) + stability fallback (:249-296)."""
    d = "sam_mask_decoder."
    B = sparse.shape[0]
    out_tokens = torch.cat([sd[d + "obj_score_token.weight"], sd[d + "iou_token.weight"], sd[d + "mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tokens[None].expand(B, -1, -1), sparse], dim=1)
    src = image_embeddings + dense
    pos_src = image_pe.expand(B, -1, -1, -1)
    b, c, h, w = src.shape
    hs, src2 = two_way_transformer(sd, cfg, src, pos_src, tokens)
    iou_token_out = hs[:, 1]
    nmt = cfg.num_multimask_outputs + 1
    mask_tokens_out = hs[:, 2:2 + nmt]
    src2 = src2.transpose(1, 2).reshape(b, c, h, w)
    feat_s0, feat_s1 = high_res_features
    u = d + "output_upscaling."
    x = F.conv_transpose2d(src2, sd[u + "0.weight"], sd[u + "0.bias"], stride=2) + feat_s1
    x = F.gelu(layer_norm_2d(sd, u + "1", x))
    x = F.gelu(F.conv_transpose2d(x, sd[u + "3.weight"], sd[u + "3.bias"], stride=2) + feat_s0)
    hyper = torch.stack([mlp(sd, f"{d}output_hypernetworks_mlps.{i}", mask_tokens_out[:, i], 3) for i in range(nmt)], 1)
    bb, cc, hh, ww = x.shape
    masks = (hyper @ x.reshape(bb, cc, hh * ww)).reshape(bb, -1, hh, ww)
    iou_pred = mlp(sd, d + "iou_prediction_head", iou_token_out, 3, sigmoid_out=True)
    obj_score = mlp(sd, d + "pred_obj_score_head", hs[:, 0], 3)
    if multimask_output:
        masks, iou_pred = masks[:, 1:], iou_pred[:, 1:]
        sam_tokens = mask_tokens_out[:, 1:]
    else:
        if cfg.dynamic_multimask_via_stability:
            mm, mi = masks[:, 1:], iou_pred[:, 1:]
            best = torch.argmax(mi, dim=-1)
            bi = torch.arange(B)
            sm = masks[:, 0:1].flatten(-2)
            dl = cfg.dynamic_multimask_stability_delta
            ai = (sm > dl).sum(-1).float()
            au = (sm > -dl).sum(-1).float()
            stab = torch.where(au > 0, ai / au, torch.ones_like(au))
            stable = stab >= cfg.dynamic_multimask_stability_thresh
            if DECISION_LOG is not None:
                top2 = torch.topk(mi, 2, dim=-1).values
                DECISION_LOG.append({"stability": stab.reshape(-1).tolist(), "stable": stable.reshape(-1).tolist(),
                                     "iou_top2_margin": (top2[:, 0] - top2[:, 1]).tolist()})
            masks = torch.where(stable[..., None, None], masks[:, 0:1], mm[bi, best].unsqueeze(1))
            iou_pred = torch.where(stable, iou_pred[:, 0:1], mi[bi, best].unsqueeze(1))
        else:
            masks, iou_pred = masks[:, 0:1], iou_pred[:, 0:1]
        sam_tokens = mask_tokens_out[:, 0:1]
    return masks, iou_pred, sam_tokens, obj_score


def forward_sam_heads(sd, cfg, backbone_features, point_coords=None, point_labels=None, mask_inputs=None,
                      high_res_features=None, multimask_output=False):
    """SAM2Base._forward_sam_heads (sam2_base.py:254-397)."""
    B = backbone_features.shape[0]
    if point_coords is None:
        point_coords = torch.zeros(B, 1, 2)
        point_labels = -torch.ones(B, 1, dtype=torch.int32)
    mask_prompt = None
    if mask_inputs is not None:
        ms = 4 * cfg.feat_size
        mask_prompt = mask_inputs.float()
        if mask_prompt.shape[-2:] != (ms, ms):
            mask_prompt = F.interpolate(mask_prompt, size=(ms, ms), align_corners=False, mode="bilinear", antialias=True)
    sparse, dense = prompt_encoder(sd, cfg, point_coords, point_labels, mask_prompt)
    low_multi, ious, tokens, obj_score = mask_decoder(sd, cfg, backbone_features, dense_pe(sd, cfg.feat_size), sparse,
                                                      dense, multimask_output, high_res_features)
    appearing = obj_score > 0
    low_multi = torch.where(appearing[:, None, None], low_multi, torch.full_like(low_multi, NO_OBJ_SCORE)).float()
    high_multi = F.interpolate(low_multi, size=(cfg.image_size, cfg.image_size), mode="bilinear", align_corners=False)
    tok = tokens[:, 0]
    if multimask_output:
        best = torch.argmax(ious, dim=-1)
        if DECISION_LOG is not None:
            top2 = torch.topk(ious, 2, dim=-1).values
            DECISION_LOG.append({"multimask_top2_margin": (top2[:, 0] - top2[:, 1]).tolist(), "best": best.tolist()})
        bi = torch.arange(B)
        low, high = low_multi[bi, best].unsqueeze(1), high_multi[bi, best].unsqueeze(1)
        if tokens.shape[1] > 1:
            tok = tokens[bi, best]
    else:
        low, high = low_multi, high_multi
    obj_ptr = mlp(sd, "obj_ptr_proj", tok, 3)
    lam = appearing.float()
    obj_ptr = lam * obj_ptr + (1 - lam) * sd["no_obj_ptr"]
    return dict(low_res_multimasks=low_multi, high_res_multimasks=high_multi, ious=ious, low_res_masks=low,
                high_res_masks=high, obj_ptr=obj_ptr, object_score_logits=obj_score)


def use_mask_as_output(sd, cfg, backbone_features, high_res_features, mask_inputs):
    """SAM2Base._use_mask_as_output (sam2_base.py:399-448)."""
    mf = mask_inputs.float()
    high = mf * 20.0 - 10.0
    low = F.interpolate(high, size=(high.shape[-2] // 4, high.shape[-1] // 4), align_corners=False, mode="bilinear",
                        antialias=True)
    ds = F.conv2d(mf, sd["mask_downsample.weight"], sd["mask_downsample.bias"], stride=4)
    obj_ptr = forward_sam_heads(sd, cfg, backbone_features, mask_inputs=ds, high_res_features=high_res_features)["obj_ptr"]
    appearing = torch.any(mf.flatten(1) > 0.0, dim=1)[..., None]
    lam = appearing.float()
    obj_ptr = lam * obj_ptr + (1 - lam) * sd["no_obj_ptr"]
    return dict(low_res_masks=low, high_res_masks=high, ious=torch.ones(mf.shape[0], 1), obj_ptr=obj_ptr,
                object_score_logits=20.0 * lam - 10.0)


# ------------------------------------------------------------------------------------------------
# memory encoder — sam2/modeling/memory_encoder.py + sam2_base.py:692-743
# ------------------------------------------------------------------------------------------------
def encode_new_memory(sd, cfg, pix_feat, pred_masks_high_res, object_score_logits, is_mask_from_pts):
    """SAM2Base._encode_new_memory (sam2_base.py:692-743) -> (maskmem_features fp32 [B,64,h,w], pos [B,64,h,w])."""
    if cfg.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts:
        m = (pred_masks_high_res > 0).float()
    else:
        m = torch.sigmoid(pred_masks_high_res)
    m = m * cfg.sigmoid_scale_for_mem_enc + cfg.sigmoid_bias_for_mem_enc
    e = "memory_encoder.mask_downsampler.encoder."
    for i in range(4):
        m = F.conv2d(m, sd[f"{e}{3 * i}.weight"], sd[f"{e}{3 * i}.bias"], stride=2, padding=1)
        m = F.gelu(layer_norm_2d(sd, f"{e}{3 * i + 1}", m))
    m = F.conv2d(m, sd[e + "12.weight"], sd[e + "12.bias"])
    x = F.conv2d(pix_feat, sd["memory_encoder.pix_feat_proj.weight"], sd["memory_encoder.pix_feat_proj.bias"]) + m
    for l in range(2):
        p = f"memory_encoder.fuser.layers.{l}."
        y = F.conv2d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=x.shape[1])
        y = layer_norm_2d(sd, p + "norm", y).permute(0, 2, 3, 1)
        y = linear(sd, p + "pwconv2", F.gelu(linear(sd, p + "pwconv1", y)))
        x = x + (sd[p + "gamma"] * y).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd["memory_encoder.out_proj.weight"], sd["memory_encoder.out_proj.bias"])
    pos = sine_pe_2d(cfg.mem_dim, x.shape[-2], x.shape[-1])[None].expand(x.shape[0], -1, -1, -1)
    appearing = (object_score_logits > 0).float()
    x = x + (1 - appearing[..., None, None]) * sd["no_obj_embed_spatial"][..., None, None].expand(*x.shape)
    return x, pos


# ------------------------------------------------------------------------------------------------
# hole filling — sam2/utils/misc.py:365-393 over csrc/connected_components.cu (integer oracle in C)
# ------------------------------------------------------------------------------------------------
def fill_holes_in_mask_scores(mask, max_area):
    from . import cc_oracle
    out = cc_oracle.fill_holes(mask.detach().float().cpu().numpy(), max_area)
    return torch.from_numpy(out).reshape(mask.shape).to(mask.device)


# ------------------------------------------------------------------------------------------------
# engine-shaped wrapper: the five compute seams the predictor calls (same interface as the CUDA
# engine in detsam2_b200/engine.py), so tests can drive the SAME host state machine with either.
# ------------------------------------------------------------------------------------------------
class OracleEngine:
    """Plain-PyTorch implementation of the engine seams, used ONLY by tests / smoke / bench.py's baseline legs.

    device="cpu" (default): the fp32 CPU checker that is pinned against the reference fixtures.
    device="cuda": the SAME code on torch's own CUDA kernels in fp32 with TF32 off (cuBLAS / cuDNN fp32) — an
        independent fp32 reference that finishes the full-size cases (large, 16 objects, steady-state bank) in seconds;
        tests/test_oracle_devices.py checks it against the CPU path.
    autocast=True (CUDA only): bf16 autocast + fused SDPA, i.e. the arithmetic the reference runs in production
        (det_sam2_RT.py:101-107, sam/transformer.py:28-41) — bench.py's `gpu_library_baseline`, never a checker."""

    name = "oracle-torch-fp32"
    image_dtype = torch.float32     # SAM2ImagePredictor feeds the fp32 normalised image (sam2_image_predictor.py:108-115)

    def __init__(self, cfg, state_dict, fill_holes=True, device="cpu", autocast=False, fused_sdpa=None):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type == "cuda":
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
            # a checker must not silently run TF32 convolutions / matmuls
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        if autocast and self.device.type != "cuda":
            raise ValueError("autocast=True is the CUDA bf16 library baseline")
        self.autocast = bool(autocast)
        # fused_sdpa=True: attention through F.scaled_dot_product_attention, the call the reference makes
        # (sam/transformer.py:268,347) — same mathematics, torch's blocked kernel instead of a materialised score matrix;
        # the timing legs of bench.py use it (the checker keeps the written-out form)
        self.fused_sdpa = self.autocast if fused_sdpa is None else bool(fused_sdpa)
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        self.fill_holes_enabled = fill_holes
        self._pos = {}

    def _ctx(self):
        """Factory calls inside the restated functions (torch.arange, torch.zeros ...) land on this engine's device."""
        import contextlib
        stack = contextlib.ExitStack()
        stack.enter_context(torch.device(self.device))
        if self.autocast:
            stack.enter_context(torch.autocast("cuda", dtype=torch.bfloat16))
        if self.fused_sdpa:
            stack.enter_context(_fused_sdpa())
        return stack

    # seam 1: sam2_video_predictor.py:1174-1212 + sam2_base.py:450-477
    def encode_image(self, image_f16):
        with self._ctx():
            bo = forward_image(self.sd, self.cfg, image_f16.to(self.device).float()[None])
        return {"fpn": bo["backbone_fpn"], "pos": bo["vision_pos_enc"]}

    def encode_images(self, images_f16):
        """Same seam for several frames (the product's engine batches them through one encoder pass; the backbone is
        per-frame, so the oracle simply encodes one after the other).  Lets the CPU tests drive the predictor's
        encode-ahead bookkeeping."""
        self.encode_images_calls = getattr(self, "encode_images_calls", 0) + 1
        return [self.encode_image(im) for im in images_f16]

    # seam 2: sam2_base.py:479-690
    def condition_on_memory(self, feats, B, frame_idx, is_init_cond_frame, output_dict, num_frames, reverse,
                            preload_idx):
        with self._ctx():
            f = feats["fpn"][-1].expand(B, -1, -1, -1).flatten(2).permute(2, 0, 1)
            p = feats["pos"][-1].expand(B, -1, -1, -1).flatten(2).permute(2, 0, 1)
            return prepare_memory_conditioned_features(self.sd, self.cfg, frame_idx, is_init_cond_frame, f, p,
                                                       output_dict, num_frames, reverse, preload_idx)

    # seam 3: sam2_base.py:254-397
    def sam_heads(self, pix_feat, feats, B, point_coords, point_labels, mask_inputs, multimask_output):
        with self._ctx():
            hr = [x.expand(B, -1, -1, -1) for x in feats["fpn"][:-1]]
            dev = lambda t: None if t is None else t.to(self.device)  # noqa: E731
            o = forward_sam_heads(self.sd, self.cfg, pix_feat, dev(point_coords), dev(point_labels), dev(mask_inputs),
                                  hr, multimask_output)
        return {"pred_masks": o["low_res_masks"].float(), "ious": o["ious"].float(), "obj_ptr": o["obj_ptr"].float(),
                "object_score_logits": o["object_score_logits"].float(), "_multimasks": o["low_res_multimasks"]}

    def decode_masks(self, pix_feat, feats, B, point_coords, point_labels, mask_inputs, multimask_output):
        """sam_prompt_encoder + sam_mask_decoder as SAM2ImagePredictor._predict calls them (sam2_image_predictor.py:395-420)."""
        with self._ctx():
            dev = lambda t: None if t is None else t.to(self.device)  # noqa: E731
            hr = [x.expand(B, -1, -1, -1) for x in feats["fpn"][:-1]]
            pc, pl = dev(point_coords), dev(point_labels)
            if pc is None:   # PromptEncoder.forward with points=None: no sparse tokens at all (prompt_encoder.py:150-160)
                sparse = torch.zeros(B, 0, self.cfg.hidden_dim)
                _, dense = prompt_encoder(self.sd, self.cfg, torch.zeros(B, 1, 2), -torch.ones(B, 1, dtype=torch.int32),
                                          dev(mask_inputs))
            else:
                sparse, dense = prompt_encoder(self.sd, self.cfg, pc, pl.to(torch.int32), dev(mask_inputs))
            masks, ious, _, _ = mask_decoder(self.sd, self.cfg, pix_feat.to(self.device), dense_pe(self.sd, self.cfg.feat_size),
                                             sparse, dense, multimask_output, hr)
        return masks.float(), ious.float()

    def connected_components(self, mask_u8):
        from . import cc_oracle
        lab, cnt = cc_oracle.connected_components(mask_u8.cpu().numpy())
        return torch.from_numpy(lab).to(mask_u8.device), torch.from_numpy(cnt).to(mask_u8.device)

    def mask_as_output(self, feats, mask_inputs):
        with self._ctx():
            B = mask_inputs.shape[0]
            pix = feats["fpn"][-1].expand(B, -1, -1, -1)
            hr = [x.expand(B, -1, -1, -1) for x in feats["fpn"][:-1]]
            o = use_mask_as_output(self.sd, self.cfg, pix, hr, mask_inputs.to(self.device))
        return {"pred_masks": o["low_res_masks"].float(), "ious": o["ious"].float(), "obj_ptr": o["obj_ptr"].float(),
                "object_score_logits": o["object_score_logits"].float()}

    # seam 4: sam2_base.py:692-743 (+ the x4 bilinear of sam2_base.py:355-360 / svp:736-741)
    def encode_memory(self, feats, B, pred_masks_low_res, object_score_logits, is_mask_from_pts):
        with self._ctx():
            S = self.cfg.image_size
            high = F.interpolate(pred_masks_low_res.to(self.device).float(), size=(S, S), mode="bilinear",
                                 align_corners=False)
            pix = feats["fpn"][-1].expand(B, -1, -1, -1)
            mf, pos = encode_new_memory(self.sd, self.cfg, pix, high, object_score_logits.to(self.device),
                                        is_mask_from_pts)
        return mf.to(torch.bfloat16), [pos]

    # seam 5: svp:1341-1348 and svp:618-642
    def fill_holes(self, pred_masks, max_area):
        if not self.fill_holes_enabled:
            return pred_masks
        return fill_holes_in_mask_scores(pred_masks, max_area)

    def resize_masks(self, masks, H, W):
        if masks.shape[-2:] == (H, W):
            return masks
        return F.interpolate(masks.float(), size=(H, W), mode="bilinear", align_corners=False)


class _fused_sdpa:
    """Context manager: route `sdpa` through F.scaled_dot_product_attention (library baseline only)."""

    def __enter__(self):
        global FUSED_SDPA
        self.prev, FUSED_SDPA = FUSED_SDPA, True

    def __exit__(self, *exc):
        global FUSED_SDPA
        FUSED_SDPA = self.prev
