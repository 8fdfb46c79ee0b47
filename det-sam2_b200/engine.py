"""CudaEngine — the per-frame SAM 2.1 hot path as sequences of hand-written sm_100a kernels launched
through the C ABI (include/detsam2.h).  torch is used for device memory, streams and a handful of
layout views only; every FLOP on the path is in libdetsam2.so.  There is no CPU / eager fallback:
constructing the engine without the library or without a CUDA device raises.

Data layout in HBM (all activations token-major, "channels last"):
  residual streams  f32  [tokens, C]          (LayerNorm / residual adds read and write f32)
  GEMM operands     bf16 [tokens, C]          (written by the producing LayerNorm / GEMM epilogue)
  weights           bf16 [N, K] (nn.Linear layout, K padded to a multiple of 8), biases / LN f32
  memory bank       bf16 [B, 4096, 64] per stored frame (exposed to the predictor as an NCHW *view*
                    so the reference's dict schema / pickle format holds), pointers f32 [B, 256]
  masks             f32  [B, 256, 256] low-res logits; the 1024^2 masks are never materialised.

Seams (one per reference function the predictor calls):
  encode_image         sam2_base.py:450-461, image_encoder.py:30-43,101-134, hieradet.py:57-168,283-299
  condition_on_memory  sam2_base.py:479-690, memory_attention.py:58-176, transformer.py:311-363
  sam_heads            sam2_base.py:254-397, prompt_encoder.py:73-171, mask_decoder.py:105-296,
                       transformer.py:90-284
  encode_memory        sam2_base.py:692-743, memory_encoder.py:17-181
  fill_holes / resize_masks   misc.py:365-393 (+ csrc/connected_components.cu), svp:618-642
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .capi import Ds2Error
from .memory_bank import plan_memory

BF16, F32 = torch.bfloat16, torch.float32


# --------------------------------------------------------------------------------------------------
# input-independent tables (computed once at load time)
# --------------------------------------------------------------------------------------------------
def _sine_pe_2d(num_pos_feats_total, h, w, temperature=10000.0):
    """position_encoding.py:78-112 (normalize=True, scale=2*pi) -> token-major [h*w, C]."""
    npf = num_pos_feats_total // 2
    y = torch.arange(1, h + 1, dtype=F32).view(-1, 1).repeat(1, w)
    x = torch.arange(1, w + 1, dtype=F32).view(1, -1).repeat(h, 1)
    y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = temperature ** (2 * (torch.arange(npf, dtype=F32) // 2) / npf)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, -1).contiguous()


def _rope_table(dim, side, theta):
    """position_encoding.py:173-182 as interleaved (cos, sin), pair-major [dim/2, side*side, 2] (the GEMM
    epilogue's lanes walk consecutive positions, so position is the fast axis)."""
    fr = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(side * side, dtype=F32)
    tx, ty = (t % side).float(), torch.div(t, side, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(tx, fr), torch.outer(ty, fr)], dim=-1)
    return torch.stack([ang.cos(), ang.sin()], dim=-1).permute(1, 0, 2).contiguous()


def _rope_axial(dim, side, theta):
    """The same table in its axial form [dim/4, side, 2]: entry [j, c] = (cos, sin)(c * freq_j).  Pair j < dim/4 of a head
    rotates by the x coordinate of the position and pair j + dim/4 by its y coordinate with the SAME frequency
    (position_encoding.py:173-182), so `_rope_table(dim, side, theta)[j, y * side + x] == _rope_axial(...)[j % (dim/4), x or y]`
    bit for bit (tests/test_kernels_gpu.py::test_rope_axial_table_equals_full_table); 32 KB instead of 4 MB at side 64,
    small enough for the GEMM epilogue to keep in shared memory."""
    fr = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    ang = torch.outer(torch.arange(side, dtype=F32), fr)           # [side, dim/4]
    return torch.stack([ang.cos(), ang.sin()], dim=-1).permute(1, 0, 2).contiguous()


def _dense_pe(gauss, side):
    """prompt_encoder.py:64-71 / position_encoding.py:131-149 -> token-major [side*side, 256]."""
    g = torch.ones(side, side, dtype=F32)
    y = (g.cumsum(0) - 0.5) / side
    x = (g.cumsum(1) - 0.5) / side
    c = (2 * torch.stack([x, y], dim=-1) - 1) @ gauss
    c = 2 * math.pi * c
    return torch.cat([c.sin(), c.cos()], dim=-1).reshape(side * side, -1).contiguous()


class _SeamGraphs:
    """CUDA-graph cache for the engine seams.  A seam body is a fixed sequence of C-ABI kernel launches
    whose arguments depend only on the seam's *signature* (shapes, flags) once its tensor inputs live in
    stable buffers; after `min_hits` eager executions of a signature the body is captured once (inputs
    copied into static buffers first) and from then on replayed — ~500 ctypes launches per frame become
    four cudaGraphLaunch calls.  Per-step variation that is not a tensor input (which stored frames make
    up the memory bank, pointer distances) lives in a device-resident table the kernels read
    (ds2_bank_assemble), never in launch arguments."""

    def __init__(self, device, enabled=True, min_hits=2, max_graphs=96):
        self.device, self.enabled, self.min_hits, self.max_graphs = device, enabled, min_hits, max_graphs
        self.hits = {}
        self.graphs = {}   # key -> (CUDAGraph, static_inputs, outputs)
        self.static_cache = {}
        self.pool = None
        self.replays = 0
        self.captures = 0
        self.captured_launches = 0   # C-ABI launches recorded into graphs (they did not execute then)
        self.replayed_launches = 0   # kernel nodes executed by graph replays

    def run(self, key, inputs, body, clone):
        """inputs: {name: tensor}; body(inputs_dict) -> tuple of tensors; clone: tuple of bools (copy the
        output out of the graph's static storage because the caller keeps it beyond the next replay)."""
        if not self.enabled:
            return body({k: v.to(self.device, non_blocking=True) for k, v in inputs.items()})
        g = self.graphs.get(key)
        if g is None:
            n = self.hits.get(key, 0) + 1
            self.hits[key] = n
            if n <= self.min_hits or len(self.graphs) >= self.max_graphs:
                if len(self.hits) > 4096:
                    self.hits.clear()
                return body({k: v.to(self.device, non_blocking=True) for k, v in inputs.items()})
            # static inputs are shared by all graphs of one seam (same name/shape/dtype): graphs of a seam
            # never run concurrently, and the ~20 bank-size signatures of the memory-attention seam would
            # otherwise each pin their own copy (device memory must stay flat over an endless stream)
            static = {}
            for k, v in inputs.items():
                ck = (key[0], k, tuple(v.shape), v.dtype)
                if ck not in self.static_cache:
                    self.static_cache[ck] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                static[k] = self.static_cache[ck]
            for k, v in inputs.items():
                static[k].copy_(v, non_blocking=True)
            body(static)  # eager on the static buffers: workspaces, TMA descriptors, func attributes exist
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            l0 = ops.launch_count()
            with torch.cuda.graph(graph, pool=self.pool):
                outs = body(static)
            nodes = ops.launch_count() - l0
            self.captured_launches += nodes
            g = (graph, static, outs, nodes)
            self.graphs[key] = g
            self.captures += 1
        graph, static, outs, nodes = g
        for k, v in inputs.items():
            static[k].copy_(v, non_blocking=True)
        graph.replay()
        self.replays += 1
        self.replayed_launches += nodes
        return tuple(o.clone() if c else o for o, c in zip(outs, clone))


class _TableRing:
    """Per-step bank tables: a small ring of pinned host buffers (so the CPU can run ahead of the GPU
    without overwriting a table whose H2D copy has not executed yet) and ONE device buffer with a fixed
    address that the captured bank kernels read."""
    MAXF, MAXP = 128, 64
    OFF_FSRC, OFF_PSRC = 0, 8 * 128
    OFF_TPOS = OFF_PSRC + 8 * 64
    OFF_DIST = OFF_TPOS + 4 * 128
    NBYTES = OFF_DIST + 4 * 64

    def __init__(self, device, slots=16):
        self.host = [torch.empty(self.NBYTES, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0
        self.dev = torch.zeros(self.NBYTES, dtype=torch.uint8, device=device)
        base = self.dev.data_ptr()
        self.fsrc, self.psrc = base + self.OFF_FSRC, base + self.OFF_PSRC
        self.tpos, self.dist = base + self.OFF_TPOS, base + self.OFF_DIST

    def upload(self, frame_ptrs, frame_tpos, ptr_ptrs, ptr_dist):
        nf, npt = len(frame_ptrs), len(ptr_ptrs)
        if nf > self.MAXF or npt > self.MAXP:
            raise Ds2Error(f"memory bank of {nf} frames / {npt} pointers exceeds the table capacity")
        i = self.i
        self.i = (i + 1) % len(self.host)
        if self.events[i] is not None:
            self.events[i].synchronize()
        h = self.host[i].numpy()
        h[self.OFF_FSRC:self.OFF_FSRC + 8 * nf].view(np.int64)[:] = frame_ptrs
        h[self.OFF_PSRC:self.OFF_PSRC + 8 * npt].view(np.int64)[:] = ptr_ptrs
        h[self.OFF_TPOS:self.OFF_TPOS + 4 * nf].view(np.int32)[:] = frame_tpos
        h[self.OFF_DIST:self.OFF_DIST + 4 * npt].view(np.float32)[:] = ptr_dist
        self.dev.copy_(self.host[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class FrameFeats:
    """Backbone outputs of one frame (what the reference keeps in cached_features, svp:1190)."""
    __slots__ = ("vis_f32", "vis_bf16", "feat_s0", "feat_s1", "pix_proj")

    def __init__(self, vis_f32, vis_bf16, feat_s0, feat_s1):
        self.vis_f32, self.vis_bf16, self.feat_s0, self.feat_s1 = vis_f32, vis_bf16, feat_s0, feat_s1
        self.pix_proj = None


class PendingFeats:
    """Features of an encoder pass that is still running on the engine's encoder stream (encode_images_async)."""
    __slots__ = ("feats", "event", "_joined")

    def __init__(self, feats, event):
        self.feats, self.event, self._joined = feats, event, False

    def wait(self):
        """Orders the CALLER's current stream after the pass and returns its list of FrameFeats (no host sync)."""
        if not self._joined:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.event)
            for f in self.feats:
                for t in (f.vis_f32, f.vis_bf16, f.feat_s0, f.feat_s1):
                    t.record_stream(cur)   # allocated on the encoder stream, read (and eventually freed) on this one
            self._joined = True
        return self.feats


class CudaEngine:
    name = "cuda-sm100a"

    def __init__(self, cfg, state_dict, device="cuda", use_graphs=True):
        if not torch.cuda.is_available():
            raise Ds2Error("CudaEngine needs a CUDA device: the hot path has no CPU implementation")
        ops._lib()  # raises if libdetsam2.so cannot be loaded / built
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._ws = {}
        # bench.py sets this to a list to collect (tag, start_event, end_event, meta) of the dominant kernel
        # (CUDA events cannot bracket a node inside a replayed graph, so that seam then runs eagerly)
        self.kernel_timers = None
        # encoder passes launched ahead of their frames run on their own stream (encode_images_async); every encoder
        # pass, on whichever stream, is ordered after the previous one: they share the encoder's workspaces
        self._enc_stream = None
        self._enc_last = None
        with torch.cuda.device(self.device):
            self.graphs = _SeamGraphs(self.device, enabled=(use_graphs and os.environ.get("DS2_GRAPHS", "1") != "0"))
            self._tables = _TableRing(self.device)
        self._pack(state_dict)

    def launches_executed(self):
        """Kernels of libdetsam2.so that have executed on the device so far: eager C-ABI launches plus
        the kernel nodes of every graph replay (launches recorded during capture did not execute)."""
        return ops.launch_count() - self.graphs.captured_launches + self.graphs.replayed_launches

    # ---------------------------------------------------------------------------------------------
    # weights
    # ---------------------------------------------------------------------------------------------
    def _pack(self, sd):
        cfg, dev = self.cfg, self.device
        sd = {k: v.detach().to(F32) for k, v in sd.items()}
        self.p = p = {}

        def f32(name, t):
            p[name] = t.to(F32).contiguous().to(dev)

        def w16(name, t):
            t = t.to(F32)
            k = t.shape[1]
            if k % 8:
                t = F.pad(t, (0, 8 - k % 8))
            p[name] = t.to(BF16).contiguous().to(dev)

        def lin(dst, src):
            w16(dst + ".w", sd[src + ".weight"].reshape(sd[src + ".weight"].shape[0], -1))
            f32(dst + ".b", sd[src + ".bias"])

        def ln(dst, src):
            f32(dst + ".w", sd[src + ".weight"])
            f32(dst + ".b", sd[src + ".bias"])

        # ---- Hiera trunk ----
        t = "image_encoder.trunk."
        E = cfg.embed_dim
        S0 = cfg.image_size // 4
        lin("pe", t + "patch_embed.proj")
        pe = F.interpolate(sd[t + "pos_embed"], size=(S0, S0), mode="bicubic")
        we = sd[t + "pos_embed_window"]
        pe = pe + we.tile([x // y for x, y in zip(pe.shape, we.shape)])
        f32("pos_embed", pe.permute(0, 2, 3, 1).reshape(S0 * S0, E))
        self.blocks = cfg.block_specs()
        for i, b in enumerate(self.blocks):
            q = f"{t}blocks.{i}."
            ln(f"b{i}.n1", q + "norm1")
            ln(f"b{i}.n2", q + "norm2")
            lin(f"b{i}.qkv", q + "attn.qkv")
            lin(f"b{i}.proj", q + "attn.proj")
            lin(f"b{i}.fc1", q + "mlp.layers.0")
            lin(f"b{i}.fc2", q + "mlp.layers.1")
            if b["dim"] != b["dim_out"]:
                lin(f"b{i}.sc", q + "proj")
            # zero-padded window tokens are 0 after norm1, so their q/k/v equal the qkv bias
            p[f"b{i}.padrow"] = sd[q + "attn.qkv.bias"].to(BF16).contiguous().to(dev)
        # ---- FPN neck (+ conv_s0 / conv_s1 folded into the lateral 1x1 convs of levels 0 / 1) ----
        n = "image_encoder.neck.convs."
        d = "sam_mask_decoder."

        def conv1(name):
            return sd[name + ".weight"].reshape(sd[name + ".weight"].shape[0], -1), sd[name + ".bias"]

        w3, b3 = conv1(n + "0.conv")   # 8E -> 256 (32^2 level)
        w2, b2 = conv1(n + "1.conv")   # 4E -> 256 (64^2 level)
        w1, b1 = conv1(n + "2.conv")   # 2E -> 256 (128^2 level)
        w0, b0 = conv1(n + "3.conv")   # E  -> 256 (256^2 level)
        ws0, bs0 = conv1(d + "conv_s0")
        ws1, bs1 = conv1(d + "conv_s1")
        w16("neck3.w", w3); f32("neck3.b", b3)
        w16("neck2.w", w2); f32("neck2.b", b2)
        w16("s1.w", ws1 @ w1); f32("s1.b", ws1 @ b1 + bs1)
        w16("s0.w", ws0 @ w0); f32("s0.b", ws0 @ b0 + bs0)
        fs = cfg.feat_size
        f32("vision_pos", _sine_pe_2d(256, fs, fs))
        f32("maskmem_pos", _sine_pe_2d(cfg.mem_dim, fs, fs))
        f32("no_mem_embed", sd["no_mem_embed"].reshape(1, -1))
        f32("maskmem_tpos", sd["maskmem_tpos_enc"].reshape(cfg.num_maskmem, -1))
        f32("no_obj_ptr", sd["no_obj_ptr"].reshape(-1))
        f32("no_obj_embed_spatial", sd["no_obj_embed_spatial"].reshape(-1))
        f32("ptr_tpos.w", sd["obj_ptr_tpos_proj.weight"])
        f32("ptr_tpos.b", sd["obj_ptr_tpos_proj.bias"])
        # ---- memory attention ----
        f32("rope", _rope_axial(256, fs, cfg.rope_theta))
        for l in range(cfg.memattn_layers):
            q = f"memory_attention.layers.{l}."
            sa, ca = q + "self_attn.", q + "cross_attn_image."
            w16(f"ma{l}.sa_qkv.w", torch.cat([sd[sa + "q_proj.weight"], sd[sa + "k_proj.weight"], sd[sa + "v_proj.weight"]]))
            f32(f"ma{l}.sa_qkv.b", torch.cat([sd[sa + "q_proj.bias"], sd[sa + "k_proj.bias"], sd[sa + "v_proj.bias"]]))
            lin(f"ma{l}.sa_out", sa + "out_proj")
            lin(f"ma{l}.ca_q", ca + "q_proj")
            lin(f"ma{l}.ca_k", ca + "k_proj")
            # softmax rows sum to one, so the 64->256 value projection commutes with P.V and folds
            # into the output projection:  Wo (P (M Wv^T + bv)) + bo = (P M)(Wo Wv)^T + (Wo bv + bo)
            wo, bo = sd[ca + "out_proj.weight"], sd[ca + "out_proj.bias"]
            wv, bv = sd[ca + "v_proj.weight"], sd[ca + "v_proj.bias"]
            w16(f"ma{l}.ca_ov.w", wo @ wv)
            f32(f"ma{l}.ca_ov.b", wo @ bv + bo)
            lin(f"ma{l}.ff1", q + "linear1")
            lin(f"ma{l}.ff2", q + "linear2")
            for k in (1, 2, 3):
                ln(f"ma{l}.n{k}", q + f"norm{k}")
        ln("ma.norm", "memory_attention.norm")
        # ---- prompt encoder ----
        pe_ = "sam_prompt_encoder."
        gauss = sd[pe_ + "pe_layer.positional_encoding_gaussian_matrix"]
        f32("gauss", gauss)
        f32("point_emb", torch.cat([sd[pe_ + f"point_embeddings.{i}.weight"] for i in range(4)]))
        f32("not_a_point", sd[pe_ + "not_a_point_embed.weight"].reshape(-1))
        f32("no_mask_embed", sd[pe_ + "no_mask_embed.weight"].reshape(1, -1))
        # dense mask prompts (add_new_mask): mask_downsample (k4 s4) + PromptEncoder.mask_downscaling
        md = pe_ + "mask_downscaling."
        f32("mp.wds", sd["mask_downsample.weight"].reshape(-1))
        f32("mp.bds", sd["mask_downsample.bias"].reshape(-1))
        f32("mp.w0", sd[md + "0.weight"].reshape(4, 4))
        f32("mp.b0", sd[md + "0.bias"])
        ln("mp.ln0", md + "1")
        f32("mp.w3", sd[md + "3.weight"].reshape(16, 16))
        f32("mp.b3", sd[md + "3.bias"])
        ln("mp.ln3", md + "4")
        w16("mp.w6", sd[md + "6.weight"].reshape(sd[md + "6.weight"].shape[0], -1))
        f32("mp.b6", sd[md + "6.bias"])
        dpe = _dense_pe(gauss, fs)
        f32("dense_pe", dpe)
        # ---- mask decoder ----
        f32("out_tokens", torch.cat([sd[d + "obj_score_token.weight"], sd[d + "iou_token.weight"], sd[d + "mask_tokens.weight"]]))
        tr = d + "transformer."

        def attn_w(prefix, n_):
            return sd[prefix + n_ + ".weight"], sd[prefix + n_ + ".bias"]

        def t2i(dst, prefix):
            """token->image attention: q from tokens; K,V from image keys in one GEMM with the
            positional term (pe Wk^T + bk) folded into a row-periodic residual table."""
            lin(dst + ".q", prefix + "q_proj")
            wk, bk = attn_w(prefix, "k_proj")
            wv, bv = attn_w(prefix, "v_proj")
            w16(dst + ".kv.w", torch.cat([wk, wv]))
            f32(dst + ".kv.tab", torch.cat([dpe @ wk.t() + bk, bv.expand(dpe.shape[0], -1)], dim=1))
            lin(dst + ".out", prefix + "out_proj")

        for l in range(cfg.decoder_depth):
            q = f"{tr}layers.{l}."
            sa = q + "self_attn."
            if l == 0:
                w16(f"dec{l}.sa_qkv.w", torch.cat([sd[sa + "q_proj.weight"], sd[sa + "k_proj.weight"], sd[sa + "v_proj.weight"]]))
                f32(f"dec{l}.sa_qkv.b", torch.cat([sd[sa + "q_proj.bias"], sd[sa + "k_proj.bias"], sd[sa + "v_proj.bias"]]))
            else:
                w16(f"dec{l}.sa_qk.w", torch.cat([sd[sa + "q_proj.weight"], sd[sa + "k_proj.weight"]]))
                f32(f"dec{l}.sa_qk.b", torch.cat([sd[sa + "q_proj.bias"], sd[sa + "k_proj.bias"]]))
                lin(f"dec{l}.sa_v", sa + "v_proj")
            lin(f"dec{l}.sa_out", sa + "out_proj")
            t2i(f"dec{l}.t2i", q + "cross_attn_token_to_image.")
            lin(f"dec{l}.mlp1", q + "mlp.layers.0")
            lin(f"dec{l}.mlp2", q + "mlp.layers.1")
            i2t = q + "cross_attn_image_to_token."
            wq, bq = attn_w(i2t, "q_proj")
            w16(f"dec{l}.i2t.q.w", wq)
            f32(f"dec{l}.i2t.q.tab", dpe @ wq.t() + bq)
            lin(f"dec{l}.i2t.k", i2t + "k_proj")
            lin(f"dec{l}.i2t.v", i2t + "v_proj")
            lin(f"dec{l}.i2t.out", i2t + "out_proj")
            for k in (1, 2, 3, 4):
                ln(f"dec{l}.n{k}", q + f"norm{k}")
        t2i("decf.t2i", tr + "final_attn_token_to_image.")
        ln("decf.n", tr + "norm_final_attn")
        up = d + "output_upscaling."
        w16("up0.w", sd[up + "0.weight"].permute(2, 3, 1, 0).reshape(-1, sd[up + "0.weight"].shape[0]))
        f32("up0.b", sd[up + "0.bias"])
        ln("up0.ln", up + "1")
        w16("up1.w", sd[up + "3.weight"].permute(2, 3, 1, 0).reshape(-1, sd[up + "3.weight"].shape[0]))
        f32("up1.b", sd[up + "3.bias"])

        def mlp3(dst, prefixes):
            for j in range(3):
                # input-major [set, in, out]: consecutive threads of ds2_mlp3 = consecutive outputs
                f32(f"{dst}.w{j + 1}", torch.stack([sd[f"{q_}layers.{j}.weight"].t() for q_ in prefixes]))
                f32(f"{dst}.b{j + 1}", torch.stack([sd[f"{q_}layers.{j}.bias"] for q_ in prefixes]))

        nmt = cfg.num_multimask_outputs + 1
        mlp3("hyper", [f"{d}output_hypernetworks_mlps.{i}." for i in range(nmt)])
        mlp3("iou", [d + "iou_prediction_head."])
        mlp3("objscore", [d + "pred_obj_score_head."])
        mlp3("objptr", ["obj_ptr_proj."])
        # ---- memory encoder ----
        me = "memory_encoder."
        e = me + "mask_downsampler.encoder."
        f32("md0.w", sd[e + "0.weight"]); f32("md0.b", sd[e + "0.bias"]); ln("md0.ln", e + "1")
        f32("md1.w", sd[e + "3.weight"]); f32("md1.b", sd[e + "3.bias"]); ln("md1.ln", e + "4")
        for j, idx in ((2, 6), (3, 9)):
            w = sd[e + f"{idx}.weight"]
            w16(f"md{j}.w", w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))  # k = (ky*3+kx)*Cin + c
            f32(f"md{j}.b", sd[e + f"{idx}.bias"])
            ln(f"md{j}.ln", e + f"{idx + 1}")
        lin("md4", e + "12")
        lin("pixproj", me + "pix_feat_proj")
        for l in range(2):
            q = f"{me}fuser.layers.{l}."
            f32(f"cx{l}.dw.w", sd[q + "dwconv.weight"].reshape(-1, 49))
            f32(f"cx{l}.dw.b", sd[q + "dwconv.bias"])
            ln(f"cx{l}.ln", q + "norm")
            lin(f"cx{l}.pw1", q + "pwconv1")
            lin(f"cx{l}.pw2", q + "pwconv2")
            f32(f"cx{l}.gamma", sd[q + "gamma"])
        lin("memout", me + "out_proj")

    # ---------------------------------------------------------------------------------------------
    # workspaces (stable addresses: TMA descriptors are cached per pointer, and the launch sequence
    # of a step can be captured into a CUDA graph)
    # ---------------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    def _buf_rows(self, name, B, rows, cols, dtype, min_rows):
        """[B, rows, cols] view of a capacity-allocated workspace whose base address does not depend on
        `rows` (the number of memory tokens changes from step to step while the bank fills up; captured
        graphs of every bank size then share one buffer).  Buffers only ever grow by adding a new one —
        an address that a captured graph uses is never freed."""
        key = ("rows", name, B, cols, dtype)
        lst = self._ws.setdefault(key, [])
        for cap, t in lst:
            if cap >= rows:
                return t[: B * rows * cols].view(B, rows, cols)
        # grow geometrically: a bank that keeps growing row by row (preload banks with many conditioning frames) must not
        # leave one retired buffer per size behind
        cap = max(rows, min_rows, int(1.5 * lst[-1][0]) if lst else 0)
        t = torch.empty(B * cap * cols, dtype=dtype, device=self.device)
        lst.append((cap, t))
        lst.sort(key=lambda e: e[0])
        return t[: B * rows * cols].view(B, rows, cols)

    # ---------------------------------------------------------------------------------------------
    # seam 1: image encoder
    # ---------------------------------------------------------------------------------------------
    def encode_image(self, image_f16):
        cfg, p = self.cfg, self.p
        S = cfg.image_size
        if image_f16.dtype != torch.float16 or tuple(image_f16.shape) != (3, S, S):
            raise Ds2Error(f"encode_image expects an fp16 [3,{S},{S}] frame, got {image_f16.dtype} {tuple(image_f16.shape)}")
        self._enc_fence()
        vis, vis16, feat_s0, feat_s1 = self.graphs.run(("enc",), {"img": image_f16.contiguous()},
                                                       self._encode_image_body, (True, True, True, True))
        self._enc_mark()
        return FrameFeats(vis, vis16, feat_s0, feat_s1)

    def _enc_fence(self):
        """The encoder's workspaces and static graph inputs exist once: a pass waits for the previous pass, which may
        have been launched on another stream (a no-op event wait when it was this stream)."""
        if self._enc_last is not None:
            torch.cuda.current_stream().wait_event(self._enc_last)

    def _enc_mark(self):
        ev = torch.cuda.Event()
        ev.record()
        self._enc_last = ev
        return ev

    def encode_images_async(self, images_f16):
        """``encode_images`` on the engine's encoder stream: returns at once with a PendingFeats handle.

        The backbone of the frames a propagate call will reach next does not depend on the tracker, so the predictor
        launches it while the tracker works on the frames already encoded.  The tracker's step is a chain of ~250
        dependent launches, many of them far smaller than the GPU (decoder, memory encoder, hole filling) and its
        largest kernel ends on a partial wave (3.46 waves of 148 CTAs at 16 objects); the encoder's CTAs fill those
        holes.  Same kernels, same arithmetic, same bits per frame — only the placement in time changes.  The pass
        starts after everything the calling stream has enqueued so far (the frames it reads are ready by then)."""
        cur = torch.cuda.current_stream()
        if self._enc_stream is None:
            self._enc_stream = torch.cuda.Stream(device=self.device)
        s = self._enc_stream
        s.wait_stream(cur)
        if images_f16.is_cuda:
            images_f16.record_stream(s)
        with torch.cuda.stream(s):
            feats = self.encode_images(images_f16)
            ev = self._enc_last    # recorded on the encoder stream by encode_images / encode_image
        return PendingFeats(feats, ev)

    def encode_images(self, images_f16):
        """Several frames through ONE pass of the encoder: fp16 [n,3,S,S] -> list of n FrameFeats.

        The frames of a chunk (or of an offline video) are all known before they are tracked, and the backbone does
        not depend on the tracker, so the predictor encodes the next few frames of its processing order together
        (predictor._get_image_feature).  Every kernel of the encoder treats rows / windows / (frame, head) pairs
        independently, so each frame's features are bit-identical to ``encode_image`` on that frame alone
        (tests/test_engine_gpu.py::test_batched_encoder_is_bit_identical); what changes is the shape of the work:
        at one frame the Hiera stage-3 GEMMs have M = 4096 rows (32 row tiles for 148 SMs, ~20 us launches dominated by
        fill/drain), at four frames the same launches carry four times the rows."""
        cfg = self.cfg
        S = cfg.image_size
        n = images_f16.shape[0]
        if images_f16.dtype != torch.float16 or tuple(images_f16.shape[1:]) != (3, S, S):
            raise Ds2Error(f"encode_images expects fp16 [n,3,{S},{S}] frames, got {images_f16.dtype} {tuple(images_f16.shape)}")
        if n == 1:
            return [self.encode_image(images_f16[0])]
        body = lambda inp: self._encode_image_body(inp, n)  # noqa: E731
        self._enc_fence()
        vis, vis16, feat_s0, feat_s1 = self.graphs.run(("enc", n), {"img": images_f16.contiguous()}, body,
                                                       (True, True, True, True))
        self._enc_mark()
        return [FrameFeats(*(t.view(n, t.shape[0] // n, t.shape[1])[i] for t in (vis, vis16, feat_s0, feat_s1)))
                for i in range(n)]

    def _encode_image_body(self, inp, nb=1):
        """Token-major throughout: activations are [nb * tokens, C] with frame f in rows [f * tokens, (f + 1) * tokens)."""
        cfg, p = self.cfg, self.p
        S = cfg.image_size
        img = inp["img"].view(nb, 3, S, S)
        Hc = Wc = S // 4
        T = Hc * Wc
        E = cfg.embed_dim
        kpad = p["pe.w"].shape[1]
        cols = self._buf("im2col", (nb * T, kpad), BF16)
        x = self._buf("x0", (nb * T, E), F32)
        for f in range(nb):
            ops.im2col_patch(img[f], cols[f * T:(f + 1) * T], S, kpad)
            ops.gemm(cols[f * T:(f + 1) * T], p["pe.w"], bias=p["pe.b"], residual=p["pos_embed"],
                     out_f32=x[f * T:(f + 1) * T])
        stage_out = []
        ends = cfg.stage_ends
        for i, b in enumerate(self.blocks):
            dim, do, heads, ws, pool = b["dim"], b["dim_out"], b["heads"], b["window"], b["q_pool"]
            T = Hc * Wc
            xn = self._buf("xn", (nb * T, dim), BF16)
            ops.layernorm(x, p[f"b{i}.n1.w"], p[f"b{i}.n1.b"], 1e-6, out_bf16=xn)
            if dim != do:
                sp = self._buf("sc_full", (nb * T, do), F32)
                ops.gemm(xn, p[f"b{i}.sc.w"], bias=p[f"b{i}.sc.b"], out_f32=sp)
                if pool:
                    shortcut = self._buf(f"sc_pool{i}", (nb * T // 4, do), F32)
                    ops.maxpool2x2(sp, shortcut, nb, Hc, Wc, do)
                else:
                    shortcut = sp
            else:
                shortcut = x
            qkv = self._buf("qkv", (nb * T, 3 * do), BF16)
            ops.gemm(xn, p[f"b{i}.qkv.w"], bias=p[f"b{i}.qkv.b"], out_bf16=qkv)
            Ho, Wo = (Hc // 2, Wc // 2) if pool else (Hc, Wc)
            Tq = Ho * Wo
            att = self._buf("att", (nb * Tq, do), BF16)
            hd = do // heads
            pad = None
            if ws > 0 and (Hc % ws or Wc % ws):
                pr = p[f"b{i}.padrow"]
                pad = (pr[:do], pr[do:2 * do], pr[2 * do:])
            ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=hd, scale=1.0 / math.sqrt(hd), B=nb,
                    Lq=Tq if ws == 0 else 0, Lk=T if ws == 0 else 0,
                    strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, Tq * do),
                    window=ws, Hm=Hc, Wm=Wc, q_pool=1 if pool else 0, pad=pad)
            if shortcut is x:
                xo = x
            else:
                xo = self._buf(f"x{i + 1}", (nb * Tq, do), F32)
            ops.gemm(att, p[f"b{i}.proj.w"], bias=p[f"b{i}.proj.b"], residual=shortcut, out_f32=xo)
            x = xo
            Hc, Wc = Ho, Wo
            xn2 = self._buf("xn", (nb * Tq, do), BF16)
            ops.layernorm(x, p[f"b{i}.n2.w"], p[f"b{i}.n2.b"], 1e-6, out_bf16=xn2)
            h = self._buf("mlp_h", (nb * Tq, 4 * do), BF16)
            ops.gemm(xn2, p[f"b{i}.fc1.w"], bias=p[f"b{i}.fc1.b"], act=2, out_bf16=h)
            ops.gemm(h, p[f"b{i}.fc2.w"], bias=p[f"b{i}.fc2.b"], residual=x, out_f32=x)
            if i in ends:
                so = self._buf(f"stage{len(stage_out)}", (nb * Tq, do), BF16)
                ops.cast_f32_bf16(x, so)
                stage_out.append((so, Hc, Wc))
        (s1, h1, w1_), (s2, h2, w2_), (s3, h3, w3_), (s4, h4, w4_) = stage_out
        lat3 = self._buf("lat3", (nb * h4 * w4_, 256), F32)
        ops.gemm(s4, p["neck3.w"], bias=p["neck3.b"], out_f32=lat3)
        lat2 = self._buf("lat2", (nb * h3 * w3_, 256), F32)
        ops.gemm(s3, p["neck2.w"], bias=p["neck2.b"], out_f32=lat2)
        vis = torch.empty((nb * h3 * w3_, 256), dtype=F32, device=self.device)
        ops.upsample2x_add(lat3, lat2, vis, nb, h4, w4_, 256)
        vis16 = torch.empty((nb * h3 * w3_, 256), dtype=BF16, device=self.device)
        ops.cast_f32_bf16(vis, vis16)
        feat_s1 = torch.empty((nb * h2 * w2_, 64), dtype=F32, device=self.device)
        ops.gemm(s2, p["s1.w"], bias=p["s1.b"], out_f32=feat_s1)
        feat_s0 = torch.empty((nb * h1 * w1_, 32), dtype=F32, device=self.device)
        ops.gemm(s1, p["s0.w"], bias=p["s0.b"], out_f32=feat_s0)
        return vis, vis16, feat_s0, feat_s1

    # ---------------------------------------------------------------------------------------------
    # seam 2: memory attention over the bank
    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _token_major(mf, B, T, C):
        """maskmem_features as stored ([B,C,h,w] view of a token-major buffer, or a plain NCHW tensor
        from a reference pickle) -> contiguous bf16 [B,T,C]."""
        t = mf.permute(0, 2, 3, 1)
        if not t.is_contiguous():
            t = t.contiguous()
        return t.reshape(B, T, C)

    def condition_on_memory(self, feats, B, frame_idx, is_init_cond_frame, output_dict, num_frames, reverse,
                            preload_idx):
        cfg, p = self.cfg, self.p
        T = cfg.feat_size * cfg.feat_size
        D, M = cfg.hidden_dim, cfg.mem_dim
        if is_init_cond_frame:
            # directly_add_no_mem_embed (sam2_base.py:651-657); a workspace (stable address) so that the
            # decoder graph of the B sequential box prompts of a detection frame can be replayed
            out = self._buf("init_pix", (1, T, D), F32)
            ops.axpby(feats.vis_f32, p["no_mem_embed"], 1.0, 1.0, b_row_mod=1, out_f32=out.view(T, D))
            return out if B == 1 else out.expand(B, T, D).contiguous()
        plan = plan_memory(frame_idx, output_dict, num_frames, reverse, preload_idx, cfg.num_maskmem,
                           cfg.max_cond_frames_in_attn, cfg.max_obj_ptrs_in_encoder)
        keep = []  # temporaries whose addresses are in the table
        fptr, ftpos, pptr, pdist = [], [], [], []
        for tpos_idx, out in plan.frames:
            mf = out["maskmem_features"]
            if mf.shape[0] != B:
                raise RuntimeError(f"memory of a stored frame has batch {mf.shape[0]}, current step has {B} objects")
            mem = self._token_major(mf.to(self.device, non_blocking=True), B, T, M)
            if mem.dtype != BF16:
                mem = mem.to(BF16)
            keep.append(mem)
            fptr.append(mem.data_ptr())
            ftpos.append(tpos_idx)
        for dist, out in plan.ptrs:
            ptr = out["obj_ptr"]
            if ptr.shape[0] != B:
                raise RuntimeError(f"object pointer of a stored frame has batch {ptr.shape[0]}, expected {B}")
            ptr = ptr.to(self.device, dtype=F32).contiguous()
            keep.append(ptr)
            pptr.append(ptr.data_ptr())
            pdist.append(dist / plan.t_diff_max)
        self._tables.upload(fptr, ftpos, pptr, pdist)
        nf, npt = len(fptr), len(pptr)
        body = lambda inp: self._memory_attention_body(inp, B, nf, npt)  # noqa: E731
        if self.kernel_timers is not None:
            return body({"vis": feats.vis_f32})[0]
        return self.graphs.run(("ma", B, nf, npt), {"vis": feats.vis_f32}, body, (False,))[0]

    def _memory_attention_body(self, inp, B, nf, npt):
        cfg, p = self.cfg, self.p
        T = cfg.feat_size * cfg.feat_size
        D, M = cfg.hidden_dim, cfg.mem_dim
        n_ptr_tok = 4 * npt
        N = nf * T + n_ptr_tok
        cap = (cfg.num_maskmem + 1) * T + 4 * cfg.max_obj_ptrs_in_encoder
        kin = self._buf_rows("bank_kin", B, N, M, BF16, cap)
        val = self._buf_rows("bank_val", B, N, M, BF16, cap)
        tb = self._tables
        ops.bank_assemble(tb.fsrc, tb.tpos, nf, tb.psrc, tb.dist, npt, p["maskmem_pos"], p["maskmem_tpos"],
                          p["ptr_tpos.w"], p["ptr_tpos.b"], kin, val, B, T, M)
        # x = curr + 0.1 * curr_pos is identical for every object at the input (memory_attention.py:139-141),
        # so everything of layer 0 that precedes the first look at the per-object bank — LN1, self-attention,
        # its residual, LN2 and the cross-attention query projection — is computed ONCE and broadcast to the
        # B objects (same arithmetic per row, 1/B of the work)
        x1 = self._buf("ma_x1", (T, D), F32)
        ops.axpby(inp["vis"], p["vision_pos"], 1.0, 0.1, out_f32=x1)
        x = self._buf("ma_x", (B, T, D), F32)
        x2 = x.view(B * T, D)
        t16 = self._buf("ma_t", (B * T, D), BF16)
        qkv = self._buf("ma_qkv", (B * T, 3 * D), BF16)
        att = self._buf("ma_att", (B, T, D), BF16)
        q16 = self._buf("ma_q", (B, T, D), BF16)
        k16 = self._buf_rows("ma_k", B, N, D, BF16, cap)
        o64 = self._buf("ma_o64", (B, T, M), BF16)
        hff = self._buf("ma_ff", (B * T, cfg.memattn_ffn), BF16)
        rope = p["rope"]
        scale = 1.0 / math.sqrt(D)
        # scratch of the cross-attention kernel (two key halves per item, see ds2_flash_args.workspace): zeroed once —
        # the kernel leaves its arrival counters at zero — and private to this engine (= one stream)
        fws = self._ws.get(("flash_ws", B))
        if os.environ.get("DS2_FLASH_TWO_PHASE", "1") == "0":   # A/B switch (tuning only): one pass per item
            fws = self._ws[("flash_ws", B)] = torch.zeros(16, dtype=torch.uint8, device=self.device)
        if fws is None:
            fws = self._ws[("flash_ws", B)] = torch.zeros(max(ops.flash_workspace_bytes(B, T, M), 16), dtype=torch.uint8,
                                                          device=self.device)
        for l in range(cfg.memattn_layers):
            w = f"ma{l}."
            # rows of the shared prefix: one object's worth in layer 0, all objects afterwards
            Bs = 1 if l == 0 else B
            xs = x1 if l == 0 else x2
            R = Bs * T
            qkv3 = qkv[:R].view(Bs, T, 3 * D)
            ops.layernorm(xs, p[w + "n1.w"], p[w + "n1.b"], 1e-5, out_bf16=t16[:R])
            ops.gemm(t16[:R], p[w + "sa_qkv.w"], bias=p[w + "sa_qkv.b"], out_bf16=qkv[:R], rope=(rope, 0, 2 * D, T, T))
            ops.flash_attn(qkv3[:, :, :D], qkv3[:, :, D:2 * D], qkv3[:, :, 2 * D:], att[:Bs], scale)
            ops.gemm(att[:Bs].view(R, D), p[w + "sa_out.w"], bias=p[w + "sa_out.b"], residual=xs, out_f32=xs)
            ops.layernorm(xs, p[w + "n2.w"], p[w + "n2.b"], 1e-5, out_bf16=t16[:R])
            ops.gemm(t16[:R], p[w + "ca_q.w"], bias=p[w + "ca_q.b"], out_bf16=q16[:Bs].view(R, D), rope=(rope, 0, D, T, T))
            if l == 0:
                x.copy_(x1.unsqueeze(0).expand(B, T, D))
                if B > 1:
                    q16[1:].copy_(q16[:1].expand(B - 1, T, D))
            ops.gemm(kin.view(B * N, M), p[w + "ca_k.w"], bias=p[w + "ca_k.b"], out_bf16=k16.view(B * N, D),
                     rope=(rope, 0, D, N, N - n_ptr_tok))
            if self.kernel_timers is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                ops.flash_attn(q16, k16, val, o64, scale, workspace=fws)
                ev1.record()
                self.kernel_timers.append(("flash_cross", ev0, ev1, {"N": N, "B": B}))
            else:
                ops.flash_attn(q16, k16, val, o64, scale, workspace=fws)
            ops.gemm(o64.view(B * T, M), p[w + "ca_ov.w"], bias=p[w + "ca_ov.b"], residual=x2, out_f32=x2)
            ops.layernorm(x2, p[w + "n3.w"], p[w + "n3.b"], 1e-5, out_bf16=t16)
            ops.gemm(t16, p[w + "ff1.w"], bias=p[w + "ff1.b"], act=1, out_bf16=hff)
            ops.gemm(hff, p[w + "ff2.w"], bias=p[w + "ff2.b"], residual=x2, out_f32=x2)
        # a workspace, not a fresh tensor: the conditioned features are consumed by the mask decoder of the
        # same step, and a per-graph output buffer (B*T*D f32 = 268 MB at 16 objects) for each of the ~20
        # bank-size signatures would grow device memory while the window fills
        out = self._buf("ma_out", (B, T, D), F32)
        ops.layernorm(x2, p["ma.norm.w"], p["ma.norm.b"], 1e-5, out_f32=out.view(B * T, D))
        return (out,)

    # ---------------------------------------------------------------------------------------------
    # seam 3: prompt encoder + mask decoder + selection epilogue
    # ---------------------------------------------------------------------------------------------
    def _gather_idx(self, B, nt):
        key = ("gidx", B, nt)
        g = self._ws.get(key)
        if g is None:
            base = torch.arange(B, dtype=torch.int32) * nt
            hyper = (base[:, None] + 2 + torch.arange(4, dtype=torch.int32)[None]).reshape(-1)
            g = (hyper.to(self.device), (base + 1).to(self.device), base.to(self.device))
            self._ws[key] = g
        return g

    def sam_heads(self, pix_feat, feats, B, point_coords, point_labels, mask_inputs, multimask_output):
        cfg, p = self.cfg, self.p
        T = cfg.feat_size * cfg.feat_size
        D = cfg.hidden_dim
        pix = pix_feat.reshape(B * T, D)
        if not pix.is_contiguous():
            pix = pix.contiguous()
        dense = None
        if mask_inputs is not None:
            # refinement click on a frame that already has a mask for this object: the previous low-resolution logits
            # come back as a DENSE prompt (svp:463-480 -> sam2_base.py:306-329 -> PromptEncoder._embed_masks)
            S4 = 4 * cfg.feat_size
            if tuple(mask_inputs.shape) != (B, 1, S4, S4):
                raise NotImplementedError(
                    f"dense prompt of shape {tuple(mask_inputs.shape)}: only the prompt encoder's own input size "
                    f"[B,1,{S4},{S4}] (previous low-resolution logits) is implemented by the CUDA engine")
            m = mask_inputs.to(device=self.device, dtype=F32).contiguous().view(B, S4, S4)
            f16 = self._buf("mp_f16", (B * T, 16), BF16)
            ops.mask_prompt_embed(m, None, None, p["mp.w0"], p["mp.b0"], p["mp.ln0.w"], p["mp.ln0.b"],
                                  p["mp.w3"], p["mp.b3"], p["mp.ln3.w"], p["mp.ln3.b"], f16)
            dense = self._buf("mp_dense", (B * T, D), F32)
            ops.gemm(f16, p["mp.w6"], bias=p["mp.b6"], out_f32=dense)
        if point_coords is not None:
            P = point_coords.shape[1]
            coords = point_coords.to(dtype=F32)
            labels = point_labels.to(dtype=torch.int32)
            if coords.shape[0] != B:
                raise Ds2Error("point prompts must have one row per object")
        else:
            # no prompt: one padding point with label -1 (sam2_base.py:302-304) + the pad point
            P = 1
            coords, labels = self._noprompt(B)
        mm = bool(multimask_output)
        if dense is not None:
            # rare interactive path: launched eagerly (the dense embedding lives in a workspace, not a graph input)
            low, iou_out, obj_ptr, score = (t.clone() for t in self._sam_heads_body(
                {"coords": coords.to(self.device), "labels": labels.to(self.device), "s0": feats.feat_s0, "s1": feats.feat_s1},
                pix, B, P, mm, dense=dense))
            S4 = 4 * cfg.feat_size
            return {"pred_masks": low, "ious": iou_out, "obj_ptr": obj_ptr, "object_score_logits": score,
                    "_all_masks": self._buf("dec_masks", (B, 4, S4, S4), F32), "_all_ious": self._buf("dec_ious", (B, 4), F32),
                    "_best_idx": self._buf("dec_best", (B,), torch.int32)}
        body = lambda inp: self._sam_heads_body(inp, pix, B, P, mm)  # noqa: E731
        low, iou_out, obj_ptr, score = self.graphs.run(
            ("sam", B, P, mm, pix.data_ptr()),
            {"coords": coords, "labels": labels, "s0": feats.feat_s0, "s1": feats.feat_s1}, body,
            (True, True, True, True))
        nt = 6 + P + 1
        S4 = 4 * cfg.feat_size
        # "_"-prefixed entries are workspace views for tests / diagnostics (valid until the next call)
        return {"pred_masks": low, "ious": iou_out, "obj_ptr": obj_ptr, "object_score_logits": score,
                "_all_masks": self._buf("dec_masks", (B, 4, S4, S4), F32), "_all_ious": self._buf("dec_ious", (B, 4), F32),
                "_best_idx": self._buf("dec_best", (B,), torch.int32)}

    def decode_masks(self, pix_feat, feats, B, point_coords, point_labels, mask_inputs, multimask_output):
        """Prompt encoder + mask decoder WITHOUT the tracking epilogue — what SAM2ImagePredictor._predict calls
        (sam2_image_predictor.py:395-420): the three multimask candidates with their predicted IoUs, or the single mask
        after the stability fallback (mask_decoder.py:143-148, 261-296); never gated by the object score.
        Returns (low_res_masks [B,C,4h,4w] f32, iou_predictions [B,C] f32), C = 3 or 1."""
        if point_coords is None:
            # PromptEncoder.forward with points=None emits NO sparse tokens (prompt_encoder.py:150-160), unlike the
            # tracker's padded "no prompt" decode: a mask-only image prompt is not wired to the CUDA decoder
            raise NotImplementedError("decode_masks without point / box prompts (mask-only prompt) is not implemented")
        dev = lambda t: None if t is None else t.to(self.device)  # noqa: E731
        out = self.sam_heads(pix_feat.to(self.device), feats, B, dev(point_coords), dev(point_labels), dev(mask_inputs),
                             multimask_output)
        if "_all_masks" not in out:
            raise Ds2Error("decode_masks: the decoder did not expose its candidate masks")
        allm, alli, best = out["_all_masks"], out["_all_ious"], out["_best_idx"]
        if multimask_output:
            return allm[:, 1:4].clone(), alli[:, 1:4].clone()
        idx = best.long()
        bi = torch.arange(B, device=self.device)
        return allm[bi, idx].unsqueeze(1).clone(), alli[bi, idx].unsqueeze(1).clone()

    def connected_components(self, mask_u8):
        """uint8 [N,1,H,W] -> (labels, areas) int32: ds2_connected_components, the drop-in for the reference's only FFI."""
        return ops.connected_components(mask_u8.to(self.device).contiguous())

    def _noprompt(self, B):
        key = ("noprompt", B)
        v = self._ws.get(key)
        if v is None:
            v = (torch.zeros((B, 1, 2), dtype=F32, device=self.device),
                 torch.full((B, 1), -1, dtype=torch.int32, device=self.device))
            self._ws[key] = v
        return v

    def _sam_heads_body(self, inp, pix, B, P, multimask_output, dense=None):
        cfg, p = self.cfg, self.p
        T = cfg.feat_size * cfg.feat_size
        D = cfg.hidden_dim
        Dh = D // 2
        H = cfg.decoder_heads
        coords, labels = inp["coords"].contiguous(), inp["labels"].contiguous()
        feat_s0, feat_s1 = inp["s0"], inp["s1"]
        nt = 6 + P + 1
        R = B * nt
        qpe = self._buf("dec_qpe", (R, D), F32)
        ops.prompt_tokens(coords, labels, B, P, p["gauss"], p["point_emb"], p["not_a_point"], p["out_tokens"],
                          cfg.image_size, qpe)
        keys = self._buf("dec_keys", (B * T, D), F32)
        keys16 = self._buf("dec_keys16", (B * T, D), BF16)
        if dense is None:
            ops.axpby(pix, p["no_mask_embed"], 1.0, 1.0, b_row_mod=1, out_f32=keys, out_bf16=keys16)
        else:  # dense mask-prompt embedding [B*T, D] instead of the broadcast no_mask_embed (prompt_encoder.py:163-170)
            ops.axpby(pix, dense, 1.0, 1.0, out_f32=keys, out_bf16=keys16)
        qf = self._buf("dec_q", (R, D), F32)       # queries (f32 residual stream)
        q16 = self._buf("dec_q16", (R, D), BF16)   # bf16(queries)
        qp16 = self._buf("dec_qp16", (R, D), BF16)  # bf16(queries + query_pe)
        tmp = self._buf("dec_tmp", (R, D), F32)
        a16 = self._buf("dec_att", (R, D), BF16)
        a8 = self._buf("dec_att_h", (R, Dh), BF16)
        kv = self._buf("dec_kv", (B * T, D), BF16)
        qi = self._buf("dec_qi", (B * T, Dh), BF16)
        ai = self._buf("dec_ai", (B * T, Dh), BF16)
        tq = self._buf("dec_tq", (R, Dh), BF16)
        tk = self._buf("dec_tk", (R, Dh), BF16)
        tv = self._buf("dec_tv", (R, Dh), BF16)
        hm = self._buf("dec_mlp", (R, cfg.decoder_mlp), BF16)
        sa3 = self._buf("dec_sa3", (R, 3 * D), BF16)

        def token_to_image(w):
            """queries += Attn(q = queries + pe, k = keys + key_pe, v = keys)   (transformer.py:190-197)"""
            ops.axpby(qf, qpe, 1.0, 1.0, out_bf16=qp16)
            ops.gemm(qp16, p[w + ".q.w"], bias=p[w + ".q.b"], out_bf16=tq)
            ops.gemm(keys16, p[w + ".kv.w"], residual=p[w + ".kv.tab"], res_row_mod=T, out_bf16=kv)
            ops.mha(tq, kv, kv[:, Dh:], a8, heads=H, head_dim=Dh // H, scale=1.0 / math.sqrt(Dh // H), B=B, Lq=nt, Lk=T,
                    strides=(Dh, D, D, Dh, nt * Dh, T * D, T * D, nt * Dh))
            ops.gemm(a8, p[w + ".out.w"], bias=p[w + ".out.b"], residual=qf, out_f32=tmp)

        ops.cast_f32_bf16(qpe, q16)
        for l in range(cfg.decoder_depth):
            w = f"dec{l}"
            hd = D // H
            if l == 0:
                # skip_first_layer_pe: queries = self_attn(queries) (replaces, transformer.py:180-182)
                ops.gemm(q16, p[w + ".sa_qkv.w"], bias=p[w + ".sa_qkv.b"], out_bf16=sa3)
                ops.mha(sa3, sa3[:, D:], sa3[:, 2 * D:], a16, heads=H, head_dim=hd, scale=1.0 / math.sqrt(hd), B=B,
                        Lq=nt, Lk=nt, strides=(3 * D, 3 * D, 3 * D, D, nt * 3 * D, nt * 3 * D, nt * 3 * D, nt * D))
                ops.gemm(a16, p[w + ".sa_out.w"], bias=p[w + ".sa_out.b"], out_f32=tmp)
            else:
                ops.axpby(qf, qpe, 1.0, 1.0, out_bf16=qp16)
                ops.gemm(qp16, p[w + ".sa_qk.w"], bias=p[w + ".sa_qk.b"], out_bf16=sa3[:, :2 * D])
                ops.gemm(q16, p[w + ".sa_v.w"], bias=p[w + ".sa_v.b"], out_bf16=sa3[:, 2 * D:])
                ops.mha(sa3, sa3[:, D:], sa3[:, 2 * D:], a16, heads=H, head_dim=hd, scale=1.0 / math.sqrt(hd), B=B,
                        Lq=nt, Lk=nt, strides=(3 * D, 3 * D, 3 * D, D, nt * 3 * D, nt * 3 * D, nt * 3 * D, nt * D))
                ops.gemm(a16, p[w + ".sa_out.w"], bias=p[w + ".sa_out.b"], residual=qf, out_f32=tmp)
            ops.layernorm(tmp, p[w + ".n1.w"], p[w + ".n1.b"], 1e-5, out_f32=qf, out_bf16=q16)
            token_to_image(w + ".t2i")
            ops.layernorm(tmp, p[w + ".n2.w"], p[w + ".n2.b"], 1e-5, out_f32=qf, out_bf16=q16)
            ops.gemm(q16, p[w + ".mlp1.w"], bias=p[w + ".mlp1.b"], act=1, out_bf16=hm)
            ops.gemm(hm, p[w + ".mlp2.w"], bias=p[w + ".mlp2.b"], residual=qf, out_f32=tmp)
            ops.layernorm(tmp, p[w + ".n3.w"], p[w + ".n3.b"], 1e-5, out_f32=qf, out_bf16=q16)
            # image -> token: keys += Attn(q = keys + key_pe, k = queries + pe, v = queries)
            ops.axpby(qf, qpe, 1.0, 1.0, out_bf16=qp16)
            ops.gemm(keys16, p[w + ".i2t.q.w"], residual=p[w + ".i2t.q.tab"], res_row_mod=T, out_bf16=qi)
            ops.gemm(qp16, p[w + ".i2t.k.w"], bias=p[w + ".i2t.k.b"], out_bf16=tk)
            ops.gemm(q16, p[w + ".i2t.v.w"], bias=p[w + ".i2t.v.b"], out_bf16=tv)
            ops.mha(qi, tk, tv, ai, heads=H, head_dim=Dh // H, scale=1.0 / math.sqrt(Dh // H), B=B, Lq=T, Lk=nt,
                    strides=(Dh, Dh, Dh, Dh, T * Dh, nt * Dh, nt * Dh, T * Dh))
            ops.gemm(ai, p[w + ".i2t.out.w"], bias=p[w + ".i2t.out.b"], residual=keys, out_f32=keys)
            ops.layernorm(keys, p[w + ".n4.w"], p[w + ".n4.b"], 1e-5, out_f32=keys, out_bf16=keys16)
        token_to_image("decf.t2i")
        hs = self._buf("dec_hs", (R, D), F32)
        ops.layernorm(tmp, p["decf.n.w"], p["decf.n.b"], 1e-5, out_f32=hs)
        # ---- heads ----
        g_hyper, g_iou, g_obj = self._gather_idx(B, nt)
        hyper = self._buf("dec_hyper", (B * 4, 32), F32)
        ops.mlp3(hs, p["hyper.w1"], p["hyper.b1"], p["hyper.w2"], p["hyper.b2"], p["hyper.w3"], p["hyper.b3"], hyper,
                 rows=B * 4, nmlp=4, gather=g_hyper)
        ious = self._buf("dec_ious", (B, 4), F32)
        ops.mlp3(hs, p["iou.w1"], p["iou.b1"], p["iou.w2"], p["iou.b2"], p["iou.w3"], p["iou.b3"], ious, rows=B, nmlp=1,
                 gather=g_iou, sigmoid_out=True)
        score = torch.empty((B, 1), dtype=F32, device=self.device)
        ops.mlp3(hs, p["objscore.w1"], p["objscore.b1"], p["objscore.w2"], p["objscore.b2"], p["objscore.w3"],
                 p["objscore.b3"], score, rows=B, nmlp=1, gather=g_obj)
        # ---- upscaling + hypernetwork mask product (mask_decoder.py:216-238) ----
        fs = cfg.feat_size
        g1 = self._buf("dec_g1", (B * T, 4 * 64), F32)
        ops.gemm(keys16, p["up0.w"], out_f32=g1)
        y1 = self._buf("dec_y1", (B * 4 * T, 64), BF16)
        ops.upscale1(g1, p["up0.b"], feat_s1, p["up0.ln.w"], p["up0.ln.b"], y1, B, fs, fs, 64)
        g2 = self._buf("dec_g2", (B * 4 * T, 4 * 32), F32)
        ops.gemm(y1, p["up1.w"], out_f32=g2)
        S4 = 4 * fs
        masks = self._buf("dec_masks", (B, 4, S4, S4), F32)
        ops.upscale2_masks(g2, p["up1.b"], feat_s0, hyper, masks, B, 2 * fs, 2 * fs, 32, 4)
        mtok = hs.view(B, nt, D)[:, 2:6].contiguous()
        low = torch.empty((B, 1, S4, S4), dtype=F32, device=self.device)
        iou_out = torch.empty((B, 1), dtype=F32, device=self.device)
        best = self._buf("dec_best", (B,), torch.int32)
        tok = self._buf("dec_tok", (B, D), F32)
        ops.sam_select(masks, ious, score, mtok, B, S4, D, bool(multimask_output), cfg.dynamic_multimask_stability_delta,
                       cfg.dynamic_multimask_stability_thresh, low, iou_out, best, tok)
        obj_ptr = torch.empty((B, D), dtype=F32, device=self.device)
        ops.mlp3(tok, p["objptr.w1"], p["objptr.b1"], p["objptr.w2"], p["objptr.b2"], p["objptr.w3"], p["objptr.b3"],
                 obj_ptr, rows=B, nmlp=1)
        ops.objptr_mix(obj_ptr, score, p["no_obj_ptr"], B, D)
        return low, iou_out, obj_ptr, score

    def mask_as_output(self, feats, mask_inputs):
        """sam2_base.py:399-448 (_use_mask_as_output): the prompt mask itself becomes the output (logits +-10,
        antialiased x1/4 for the low-resolution copy), and the object pointer comes from a SAM decode that takes
        the mask as a DENSE prompt.  mask_inputs: [B, 1, S, S] 0/1 floats at the model resolution."""
        cfg, p = self.cfg, self.p
        B = mask_inputs.shape[0]
        S = cfg.image_size
        S4 = 4 * cfg.feat_size
        T = cfg.feat_size * cfg.feat_size
        D = cfg.hidden_dim
        if tuple(mask_inputs.shape) != (B, 1, S, S):
            raise Ds2Error(f"mask prompt must be [B,1,{S},{S}], got {tuple(mask_inputs.shape)}")
        m = mask_inputs.to(device=self.device, dtype=F32).contiguous().view(B, S, S)
        _, stats = ops.mask_pack_stats(m, bits=False)
        if not bool((stats[:, 0] > 0).any()):
            # all-empty prompt (_get_empty_mask_ptr, svp:769-804, for objects missing from a prompted frame): with
            # fixed_no_obj_ptr the pointer of an empty mask is exactly no_obj_ptr, whatever SAM computes
            return {
                "pred_masks": torch.full((B, 1, S4, S4), -10.0, dtype=F32, device=self.device),
                "ious": torch.ones((B, 1), dtype=F32, device=self.device),
                "obj_ptr": p["no_obj_ptr"].reshape(1, -1).expand(B, -1).contiguous(),
                "object_score_logits": torch.full((B, 1), -10.0, dtype=F32, device=self.device),
            }
        low = torch.empty((B, 1, S4, S4), dtype=F32, device=self.device)
        ops.downsample4_aa(m, low.view(B, S4, S4), 20.0, -10.0)
        f16 = self._buf("mp_f16", (B * T, 16), BF16)
        ops.mask_prompt_embed(m, p["mp.wds"], p["mp.bds"], p["mp.w0"], p["mp.b0"], p["mp.ln0.w"], p["mp.ln0.b"],
                              p["mp.w3"], p["mp.b3"], p["mp.ln3.w"], p["mp.ln3.b"], f16)
        dense = self._buf("mp_dense", (B * T, D), F32)
        ops.gemm(f16, p["mp.w6"], bias=p["mp.b6"], out_f32=dense)
        # SAM decode without point prompts (one padding point), single-mask output, raw backbone features
        pix = feats.vis_f32.reshape(1, T, D).expand(B, T, D).contiguous().view(B * T, D)
        coords, labels = self._noprompt(B)
        _, _, obj_ptr, _ = self._sam_heads_body({"coords": coords, "labels": labels, "s0": feats.feat_s0, "s1": feats.feat_s1},
                                                pix, B, 1, False, dense=dense)
        # lambda = mask non-empty (sam2_base.py:433-447): score = +-10, pointer mixed once more with no_obj_ptr
        score = ((stats[:, 0:1] > 0).to(F32) * 20.0 - 10.0).contiguous()
        obj_ptr = obj_ptr.clone()
        ops.objptr_mix(obj_ptr, score, p["no_obj_ptr"], B, D)
        return {"pred_masks": low, "ious": torch.ones((B, 1), dtype=F32, device=self.device), "obj_ptr": obj_ptr,
                "object_score_logits": score}

    # ---------------------------------------------------------------------------------------------
    # seam 4: memory encoder
    # ---------------------------------------------------------------------------------------------
    def encode_memory(self, feats, B, pred_masks_low_res, object_score_logits, is_mask_from_pts):
        cfg, p = self.cfg, self.p
        fs = cfg.feat_size
        T = fs * fs
        D, M = cfg.hidden_dim, cfg.mem_dim
        Sl = 4 * fs
        low = pred_masks_low_res.to(dtype=F32).reshape(B, Sl, Sl)
        score = object_score_logits.to(dtype=F32).reshape(B)
        binarize = bool(cfg.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts)
        body = lambda inp: self._encode_memory_body(inp, B, binarize)  # noqa: E731
        (mem,) = self.graphs.run(("me", B, binarize), {"vis16": feats.vis_bf16, "low": low, "score": score}, body,
                                 (True,))
        maskmem = mem.view(B, fs, fs, M).permute(0, 3, 1, 2)  # [B,64,h,w] view, channels-last storage
        pos = p["maskmem_pos"].view(1, fs, fs, M).permute(0, 3, 1, 2).expand(B, -1, -1, -1)
        return maskmem, [pos]

    def _encode_memory_body(self, inp, B, binarize):
        cfg, p = self.cfg, self.p
        fs = cfg.feat_size
        T = fs * fs
        D, M = cfg.hidden_dim, cfg.mem_dim
        Sl = 4 * fs
        low, score = inp["low"].contiguous(), inp["score"].contiguous()
        pix_proj = self._buf("me_pixproj", (T, D), F32)
        ops.gemm(inp["vis16"], p["pixproj.w"], bias=p["pixproj.b"], out_f32=pix_proj)
        m1 = self._buf("me_m1", (B, 2 * Sl, 2 * Sl, 4), BF16)
        ops.maskds_stage1(low, B, Sl, binarize, cfg.sigmoid_scale_for_mem_enc, cfg.sigmoid_bias_for_mem_enc,
                          p["md0.w"], p["md0.b"], p["md0.ln.w"], p["md0.ln.b"], m1)
        m2 = self._buf("me_m2", (B, Sl, Sl, 16), BF16)
        ops.maskds_conv(m1, B, 2 * Sl, 2 * Sl, 4, 16, p["md1.w"], p["md1.b"], p["md1.ln.w"], p["md1.ln.b"], m2)
        c3 = self._buf("me_c3", (B * (Sl // 2) ** 2, 144), BF16)
        ops.im2col_k3s2(m2, c3, B, Sl, Sl, 16)
        g3 = self._buf("me_g3", (B * (Sl // 2) ** 2, 64), F32)
        ops.gemm(c3, p["md2.w"], bias=p["md2.b"], out_f32=g3)
        m3 = self._buf("me_m3", (B * (Sl // 2) ** 2, 64), BF16)
        ops.layernorm(g3, p["md2.ln.w"], p["md2.ln.b"], 1e-6, act=2, out_bf16=m3)
        c4 = self._buf("me_c4", (B * T, 576), BF16)
        ops.im2col_k3s2(m3, c4, B, Sl // 2, Sl // 2, 64)
        g4 = self._buf("me_g4", (B * T, D), F32)
        ops.gemm(c4, p["md3.w"], bias=p["md3.b"], out_f32=g4)
        m4 = self._buf("me_m4", (B * T, D), BF16)
        ops.layernorm(g4, p["md3.ln.w"], p["md3.ln.b"], 1e-6, act=2, out_bf16=m4)
        x = self._buf("me_x", (B * T, D), F32)
        ops.gemm(m4, p["md4.w"], bias=p["md4.b"], residual=pix_proj, res_row_mod=T, out_f32=x)
        dw = self._buf("me_dw", (B * T, D), F32)
        t16 = self._buf("me_t16", (B * T, D), BF16)
        h = self._buf("me_h", (B * T, 4 * D), BF16)
        for l in range(2):
            w = f"cx{l}."
            ops.dwconv7(x, p[w + "dw.w"], p[w + "dw.b"], dw, B, fs, fs, D)
            ops.layernorm(dw, p[w + "ln.w"], p[w + "ln.b"], 1e-6, out_bf16=t16)
            ops.gemm(t16, p[w + "pw1.w"], bias=p[w + "pw1.b"], act=2, out_bf16=h)
            ops.gemm(h, p[w + "pw2.w"], bias=p[w + "pw2.b"], gamma=p[w + "gamma"], residual=x, out_f32=x)
        ops.cast_f32_bf16(x, t16)
        o = self._buf("me_o", (B * T, M), F32)
        ops.gemm(t16, p["memout.w"], bias=p["memout.b"], out_f32=o)
        mem = torch.empty((B, T, M), dtype=BF16, device=self.device)
        ops.memenc_finish(o, score, p["no_obj_embed_spatial"], mem, B, T, M)
        return (mem,)

    # ---------------------------------------------------------------------------------------------
    # seam 5: post-processing
    # ---------------------------------------------------------------------------------------------
    def fill_holes(self, pred_masks, max_area):
        B, _, H, W = pred_masks.shape
        out = pred_masks.to(self.device, dtype=F32).clone()
        lab = self._buf("fh_lab", (B, H, W), torch.int32)
        cnt = self._buf("fh_cnt", (B, H, W), torch.int32)
        ops.fill_holes(out, lab, cnt, B, H, W, int(max_area))
        return out

    def resize_masks(self, masks, H, W):
        if masks.shape[-2:] == (H, W):
            return masks
        N = masks.shape[0] * masks.shape[1]
        x = masks.to(self.device, dtype=F32).contiguous()
        y = torch.empty((masks.shape[0], masks.shape[1], H, W), dtype=F32, device=self.device)
        ops.resize_bilinear(x, y, N, masks.shape[-2], masks.shape[-1], H, W)
        return y
