"""Thin torch-tensor front-end over the C ABI (capi.py).  torch is plumbing only: device memory,
streams.  Every function launches hand-written sm_100a kernels on torch's current stream and raises
``capi.Ds2Error`` on a bad argument or a failed launch — there is no eager / CPU fallback.
"""
import ctypes as C

import torch

from . import capi

BF16 = torch.bfloat16
F32 = torch.float32


def _lib():
    return capi.load()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(rc, what):
    if rc != 0:
        capi.check(rc, what)


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise capi.Ds2Error(f"{name}: tensor must live on a CUDA device (no CPU path exists)")
    if t.dtype != dtype:
        raise capi.Ds2Error(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a, w, *, bias=None, act=0, gamma=None, residual=None, res_row_mod=0, out_f32=None, out_bf16=None,
         rope=None, impl=0):
    """out = epilogue(a @ w.T).  a [M,K] bf16 (row-strided ok), w [N,K] bf16.
    rope = (cs, col0, col1, rows_per_batch, row_limit) with cs f32, either the full pair-major table [128,P,2] or its
    axial form [64,side,2] (P = side^2; staged in shared memory by the kernel — the product path)."""
    _req(a, BF16, "gemm.a"); _req(w, BF16, "gemm.w"); _req(bias, F32, "gemm.bias")
    _req(gamma, F32, "gemm.gamma"); _req(residual, F32, "gemm.residual")
    _req(out_f32, F32, "gemm.out_f32"); _req(out_bf16, BF16, "gemm.out_bf16")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    g = capi.GemmArgs()
    g.A, g.W, g.lda, g.ldw = a.data_ptr(), w.data_ptr(), a.stride(0), w.stride(0)
    g.M, g.N, g.K = M, N, K
    g.bias, g.gamma = (bias.data_ptr() if bias is not None else None), (gamma.data_ptr() if gamma is not None else None)
    if residual is not None:
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    g.res_row_mod, g.act = res_row_mod, act
    if out_f32 is not None:
        assert out_f32.stride(1) == 1
        g.out_f32, g.ldc = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.stride(1) == 1
        g.out_bf16, g.ldc_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if rope is not None:
        cs, c0, c1, rpb, lim = rope
        _req(cs, F32, "gemm.rope_cs")
        g.rope_col0, g.rope_col1 = c0, c1
        assert cs.shape[0] in (64, 128) and cs.is_contiguous()
        if cs.shape[0] == 64:
            g.rope_axial, g.rope_side, g.rope_period = cs.data_ptr(), cs.shape[1], cs.shape[1] * cs.shape[1]
        else:
            g.rope_cs, g.rope_period = cs.data_ptr(), cs.shape[1]
        g.rope_rows_per_batch, g.rope_row_limit = rpb, lim
    g.impl = impl
    _chk(_lib().ds2_gemm(C.byref(g), _stream()), "ds2_gemm")


_FLASH_IMPL = int(__import__("os").environ.get("DS2_FLASH_IMPL", "0"))  # kernel-variant A/B switch (tuning only)


def flash_workspace_bytes(B, Lq, DV):
    return int(_lib().ds2_flash_workspace_bytes(B, Lq, DV))


def flash_attn(q, k, v, out, scale, impl=None, workspace=None, flags=0):
    """q [B,Lq,256] k [B,Lk,256] v [B,Lk,DV] out [B,Lq,DV], bf16; last dim contiguous.
    workspace: zero-initialised uint8 device tensor of at least flash_workspace_bytes(B, Lq, DV) bytes (DV = 64 only):
    every item is then computed as two key halves and the partial last wave may run as two CTAs per item."""
    if impl is None:
        impl = _FLASH_IMPL
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, BF16, "flash." + n)
        assert t.stride(2) == 1
    B, Lq, _ = q.shape
    f = capi.FlashArgs()
    f.q, f.k, f.v, f.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    f.ldq, f.ldk, f.ldv, f.ldo = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    f.bsq, f.bsk, f.bsv, f.bso = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    f.B, f.Lq, f.Lk, f.DV = B, Lq, k.shape[1], v.shape[2]
    f.scale, f.impl = scale, impl
    if workspace is not None:
        assert workspace.dtype == torch.uint8 and workspace.is_cuda and workspace.is_contiguous()
        f.workspace, f.workspace_bytes = workspace.data_ptr(), workspace.numel()
    f.impl_flags = flags
    _chk(_lib().ds2_flash_attn(C.byref(f), _stream()), "ds2_flash_attn")


def mha(q, k, v, out, *, heads, head_dim, scale, B, Lq=0, Lk=0, strides, window=0, Hm=0, Wm=0, q_pool=0,
        Lk_valid=0, pad=None):
    """strides = (q_tok, k_tok, v_tok, o_tok, q_bs, k_bs, v_bs, o_bs) in elements; q/k/v/out are bf16
    tensors whose data_ptr() is the address of (batch 0, token 0, head 0)."""
    m = capi.MhaArgs()
    m.q, m.k, m.v, m.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    (m.q_tok_stride, m.k_tok_stride, m.v_tok_stride, m.o_tok_stride, m.q_bs, m.k_bs, m.v_bs, m.o_bs) = strides
    m.B, m.H, m.D = B, heads, head_dim
    m.Lq, m.Lk = Lq, Lk
    m.window, m.Hm, m.Wm, m.q_pool = window, Hm, Wm, q_pool
    m.Lk_valid, m.scale = Lk_valid, scale
    if pad is not None:
        m.pad_q, m.pad_k, m.pad_v = pad[0].data_ptr(), pad[1].data_ptr(), pad[2].data_ptr()
    _chk(_lib().ds2_mha(C.byref(m), _stream()), "ds2_mha")


def layernorm(x, w, b, eps, *, out_f32=None, out_bf16=None, act=0, pos=None, pos_row_mod=0, out2_bf16=None):
    """x [rows, C] f32 or bf16 (row-strided ok)."""
    rows, Cc = x.shape
    a = capi.LnArgs()
    if x.dtype == F32:
        a.x = x.data_ptr()
    else:
        _req(x, BF16, "layernorm.x")
        a.x_bf16 = x.data_ptr()
    a.ldx, a.rows, a.C = x.stride(0), rows, Cc
    a.w, a.b, a.eps, a.act = w.data_ptr(), b.data_ptr(), eps, act
    ldo = None
    for o in (out_f32, out_bf16, out2_bf16):
        if o is not None:
            assert o.stride(1) == 1 and (ldo is None or ldo == o.stride(0))
            ldo = o.stride(0)
    a.ldo = ldo
    a.out_f32, a.out_bf16, a.out2_bf16 = _p(out_f32), _p(out_bf16), _p(out2_bf16)
    if pos is not None:
        a.pos, a.pos_row_mod = pos.data_ptr(), pos_row_mod
    _chk(_lib().ds2_layernorm(C.byref(a), _stream()), "ds2_layernorm")


def axpby(a, b, alpha, beta, *, b_row_mod=0, out_f32=None, out_bf16=None):
    rows, Cc = a.shape
    _chk(_lib().ds2_axpby(_p(a), _p(b), rows, Cc, b_row_mod, alpha, beta, _p(out_f32), _p(out_bf16), _stream()),
         "ds2_axpby")


def cast_f32_bf16(x, y):
    _chk(_lib().ds2_cast_f32_bf16(_p(x), _p(y), x.numel(), _stream()), "ds2_cast_f32_bf16")


def cast_bf16_f32(x, y):
    _chk(_lib().ds2_cast_bf16_f32(_p(x), _p(y), x.numel(), _stream()), "ds2_cast_bf16_f32")


def maxpool2x2(x, y, B, Hm, Wm, Cc):
    _chk(_lib().ds2_maxpool2x2(_p(x), _p(y), B, Hm, Wm, Cc, _stream()), "ds2_maxpool2x2")


def upsample2x_add(top, lat, y, B, Hm, Wm, Cc):
    _chk(_lib().ds2_upsample2x_add(_p(top), _p(lat), _p(y), B, Hm, Wm, Cc, _stream()), "ds2_upsample2x_add")


def im2col_patch(frame_f16, out_bf16, S, Kpad):
    _req(frame_f16, torch.float16, "im2col_patch.frame")
    _chk(_lib().ds2_im2col_patch(_p(frame_f16), _p(out_bf16), S, Kpad, _stream()), "ds2_im2col_patch")


def im2col_k3s2(x, out, B, Hi, Wi, Cc):
    _chk(_lib().ds2_im2col_k3s2(_p(x), _p(out), B, Hi, Wi, Cc, _stream()), "ds2_im2col_k3s2")


def dwconv7(x, w, bias, y, B, Hm, Wm, Cc):
    _chk(_lib().ds2_dwconv7(_p(x), _p(w), _p(bias), _p(y), B, Hm, Wm, Cc, _stream()), "ds2_dwconv7")


def maskds_stage1(lowres, B, Sl, binarize, scale, bias, w, b, ln_w, ln_b, out):
    _chk(_lib().ds2_maskds_stage1(_p(lowres), B, Sl, int(binarize), scale, bias, _p(w), _p(b), _p(ln_w), _p(ln_b),
                                  _p(out), _stream()), "ds2_maskds_stage1")


def maskds_conv(x, B, Hi, Wi, Cin, Cout, w, b, ln_w, ln_b, out):
    _chk(_lib().ds2_maskds_conv(_p(x), B, Hi, Wi, Cin, Cout, _p(w), _p(b), _p(ln_w), _p(ln_b), _p(out), _stream()),
         "ds2_maskds_conv")


def upscale1(g, bias, skip, ln_w, ln_b, y, B, Hm, Wm, Cc):
    _chk(_lib().ds2_upscale1(_p(g), _p(bias), _p(skip), _p(ln_w), _p(ln_b), _p(y), B, Hm, Wm, Cc, _stream()),
         "ds2_upscale1")


def upscale2_masks(g, bias, skip, hyper, masks, B, Hm, Wm, Cc, M):
    _chk(_lib().ds2_upscale2_masks(_p(g), _p(bias), _p(skip), _p(hyper), _p(masks), B, Hm, Wm, Cc, M, _stream()),
         "ds2_upscale2_masks")


def mlp3(x, w1, b1, w2, b2, w3, b3, y, *, rows, nmlp, gather=None, sigmoid_out=False):
    a = capi.Mlp3Args()
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    a.gather = gather.data_ptr() if gather is not None else None
    a.rows, a.nmlp = rows, nmlp
    # weights are input-major: w1 [nmlp, din, dh], w2 [nmlp, dh, dh], w3 [nmlp, dh, dout]
    a.din, a.dh, a.dout = w1.shape[-2], w1.shape[-1], w3.shape[-1]
    assert w2.shape[-2] == a.dh and w2.shape[-1] == a.dh and w3.shape[-2] == a.dh
    a.w1, a.b1, a.w2, a.b2, a.w3, a.b3 = (t.data_ptr() for t in (w1, b1, w2, b2, w3, b3))
    a.sigmoid_out = int(sigmoid_out)
    a.y, a.ldy = y.data_ptr(), y.stride(0)
    _chk(_lib().ds2_mlp3(C.byref(a), _stream()), "ds2_mlp3")


def sam_select(all_masks, ious, obj_score, mask_tokens, B, S, Cc, multimask, delta, thresh, low_res, iou_out,
               best_idx, token_out):
    _chk(_lib().ds2_sam_select(_p(all_masks), _p(ious), _p(obj_score), _p(mask_tokens), B, S, Cc, int(multimask),
                               delta, thresh, _p(low_res), _p(iou_out), _p(best_idx), _p(token_out), _stream()),
         "ds2_sam_select")


def objptr_mix(ptr, obj_score, no_obj_ptr, B, Cc):
    _chk(_lib().ds2_objptr_mix(_p(ptr), _p(obj_score), _p(no_obj_ptr), B, Cc, _stream()), "ds2_objptr_mix")


def bank_gather(mem, pos, tpos, kin, val, B, T, Cc, dst_bs, row0):
    _chk(_lib().ds2_bank_gather(_p(mem), _p(pos), _p(tpos), _p(kin), _p(val), B, T, Cc, dst_bs, row0, _stream()),
         "ds2_bank_gather")


def bank_ptr(ptr, tpos, kin, val, B, dst_bs, row0):
    _chk(_lib().ds2_bank_ptr(_p(ptr), _p(tpos), _p(kin), _p(val), B, dst_bs, row0, _stream()), "ds2_bank_ptr")


def prompt_tokens(coords, labels, B, P, gauss, point_emb, not_a_point, out_tokens, image_size, tokens):
    _chk(_lib().ds2_prompt_tokens(_p(coords), _p(labels), B, P, _p(gauss), _p(point_emb), _p(not_a_point),
                                  _p(out_tokens), out_tokens.shape[0], float(image_size), _p(tokens), _stream()),
         "ds2_prompt_tokens")


def bank_ptr_pe(ptr, dist_norm, w, bias, kin, val, B, dst_bs, row0):
    _chk(_lib().ds2_bank_ptr_pe(_p(ptr), float(dist_norm), _p(w), _p(bias), _p(kin), _p(val), B, dst_bs, row0,
                                _stream()), "ds2_bank_ptr_pe")


def bank_assemble(frame_src, frame_tpos, nf, ptr_src, ptr_dist, np_, pos, tpos_table, ptr_w, ptr_bias, kin, val, B, T,
                  Cc):
    """frame_src / frame_tpos / ptr_src / ptr_dist are DEVICE addresses (ints) of the per-step tables."""
    _chk(_lib().ds2_bank_assemble(C.c_void_p(frame_src), C.c_void_p(frame_tpos), nf, C.c_void_p(ptr_src),
                                  C.c_void_p(ptr_dist), np_, _p(pos), _p(tpos_table), _p(ptr_w), _p(ptr_bias),
                                  _p(kin), _p(val), B, T, Cc, _stream()), "ds2_bank_assemble")


def memenc_finish(x, score, no_obj_embed, out, B, T, Cc):
    _chk(_lib().ds2_memenc_finish(_p(x), _p(score), _p(no_obj_embed), _p(out), B, T, Cc, _stream()),
         "ds2_memenc_finish")


def connected_components(mask_u8):
    """Drop-in for sam2._C.get_connected_componnets: uint8 [N,1,H,W] cuda -> (labels, counts) int32."""
    if not mask_u8.is_cuda or mask_u8.dtype != torch.uint8 or mask_u8.dim() != 4 or mask_u8.shape[1] != 1:
        raise capi.Ds2Error("connected_components: expected a uint8 CUDA tensor of shape [N,1,H,W]")
    N, _, H, W = mask_u8.shape
    m = mask_u8.contiguous()
    labels = torch.empty((N, 1, H, W), dtype=torch.int32, device=m.device)
    counts = torch.empty_like(labels)
    _chk(_lib().ds2_connected_components(_p(m), _p(labels), _p(counts), N, H, W, _stream()),
         "ds2_connected_components")
    return labels, counts


def fill_holes(scores, labels_ws, counts_ws, N, H, W, max_area):
    _chk(_lib().ds2_fill_holes(_p(scores), _p(labels_ws), _p(counts_ws), N, H, W, max_area, _stream()),
         "ds2_fill_holes")


def resize_bilinear(x, y, N, Hi, Wi, Ho, Wo):
    _chk(_lib().ds2_resize_bilinear(_p(x), _p(y), N, Hi, Wi, Ho, Wo, _stream()), "ds2_resize_bilinear")


def threshold_pack(x, bits):
    _chk(_lib().ds2_threshold_pack(_p(x), _p(bits), x.numel(), _stream()), "ds2_threshold_pack")


def downsample4_aa(x, y, scale, bias):
    """x f32 [B, S, S] -> y f32 [B, S/4, S/4]: antialiased bilinear of x * scale + bias."""
    _req(x, F32, "downsample4_aa.x"); _req(y, F32, "downsample4_aa.y")
    _chk(_lib().ds2_downsample4_aa(_p(x), _p(y), x.shape[0], x.shape[-1], float(scale), float(bias), _stream()),
         "ds2_downsample4_aa")


def mask_prompt_embed(mask, wds, bds, w0, b0, ln0_w, ln0_b, w3, b3, ln3_w, ln3_b, out_bf16):
    _req(mask, F32, "mask_prompt_embed.mask"); _req(out_bf16, BF16, "mask_prompt_embed.out")
    _chk(_lib().ds2_mask_prompt_embed(_p(mask), mask.shape[0], mask.shape[-1], _p(wds), _p(bds), _p(w0), _p(b0), _p(ln0_w),
                                      _p(ln0_b), _p(w3), _p(b3), _p(ln3_w), _p(ln3_b), _p(out_bf16), _stream()),
         "ds2_mask_prompt_embed")


def mask_pack_stats(masks, bits=None, stats=None):
    """masks f32 [N, 1, H, W] or [N, H, W] on CUDA -> (bits uint8 [N, H, ceil(W/8)], stats int64 [N, 3] = area, sum_x,
    sum_y); pass preallocated outputs to avoid allocations on the hot path."""
    _req(masks, F32, "mask_pack_stats.masks")
    m = masks.contiguous()
    N, H, W = m.shape[0], m.shape[-2], m.shape[-1]
    if bits is False:
        bits = None   # statistics only
    elif bits is None:
        bits = torch.empty((N, H, (W + 7) // 8), dtype=torch.uint8, device=m.device)
    if stats is None:
        stats = torch.empty((N, 3), dtype=torch.int64, device=m.device)
    _chk(_lib().ds2_mask_pack_stats(_p(m), _p(bits), _p(stats), N, H, W, _stream()), "ds2_mask_pack_stats")
    return bits, stats


def letterbox_frames(src_u8, lut, out, new_h, new_w, top, left, pad_value=114):
    """Detector pre-processing on the device: src_u8 uint8 [N, Hv, Wv, 3] RGB, lut [256] (out's dtype: v / 255),
    out fp32 / fp16 [N, 3, Hd, Wd] contiguous; the resized frame lands at rows [top, top+new_h), cols [left, left+new_w)."""
    _req(src_u8, torch.uint8, "letterbox_frames.src_u8")
    if src_u8.dim() != 4 or src_u8.shape[3] != 3 or src_u8.stride(3) != 1 or src_u8.stride(2) != 3:
        raise capi.Ds2Error(f"letterbox_frames: src must be [N, Hv, Wv, 3] with packed RGB pixels, got {tuple(src_u8.shape)}")
    if out.dtype not in (torch.float32, torch.float16) or lut.dtype != out.dtype or not out.is_cuda or not lut.is_cuda:
        raise capi.Ds2Error("letterbox_frames: out / lut must be CUDA fp32 or fp16 tensors of one dtype")
    N, Hv, Wv, _ = src_u8.shape
    if out.dim() != 4 or out.shape[0] != N or out.shape[1] != 3 or not out.is_contiguous() or lut.numel() != 256:
        raise capi.Ds2Error("letterbox_frames: out must be contiguous [N, 3, Hd, Wd] and lut must hold 256 entries")
    _chk(_lib().ds2_letterbox_frames(_p(src_u8), N, Hv, Wv, src_u8.stride(1),
                                     src_u8.stride(0) if N > 1 else max(src_u8.stride(0), src_u8.stride(1) * Hv),
                                     _p(lut), _p(out), 1 if out.dtype == torch.float32 else 0, out.shape[2], out.shape[3],
                                     int(new_h), int(new_w), int(top), int(left), int(pad_value), _stream()),
         "ds2_letterbox_frames")
    return out


def ingest_frames(src_u8, lut, out):
    """Frame ingest (misc.py:336-359 on the device): src_u8 uint8 [N, Hv, Wv, 3] RGB (rows / frames may be strided),
    lut int16 [3, 256] = fp16 bit patterns of the normalised byte values, out fp16 [N, 3, S, S] contiguous."""
    for t, dt, nm in ((src_u8, torch.uint8, "src_u8"), (lut, torch.int16, "lut"), (out, torch.float16, "out")):
        _req(t, dt, f"ingest_frames.{nm}")
    if src_u8.dim() != 4 or src_u8.shape[3] != 3 or src_u8.stride(3) != 1 or src_u8.stride(2) != 3:
        raise capi.Ds2Error(f"ingest_frames: src must be [N, Hv, Wv, 3] with packed RGB pixels, got {tuple(src_u8.shape)}")
    N, Hv, Wv, _ = src_u8.shape
    S = out.shape[-1]
    if tuple(out.shape) != (N, 3, S, S) or not out.is_contiguous() or tuple(lut.shape) != (3, 256) or not lut.is_contiguous():
        raise capi.Ds2Error("ingest_frames: out must be contiguous [N, 3, S, S] and lut contiguous [3, 256]")
    _chk(_lib().ds2_ingest_frames(_p(src_u8), N, Hv, Wv, src_u8.stride(1), src_u8.stride(0) if N > 1 else max(src_u8.stride(0), src_u8.stride(1) * Hv),
                                  _p(lut), _p(out), S, _stream()), "ds2_ingest_frames")
    return out


def launch_count():
    return int(_lib().ds2_launch_count())
