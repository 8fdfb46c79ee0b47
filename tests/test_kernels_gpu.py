"""Per-kernel parity on the GPU: every entry point of the C ABI against a plain torch fp32
restatement of the same op on the same seeded inputs (floating point -> tolerance stated per test;
integer work -> bit exact)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from detsam2_b200 import ops as _ops
    _ops._lib()
    return _ops


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 144, 160), (16, 256, 256), (4096, 432, 144),
                                   (1000, 1152, 576), (4096, 32, 256), (333, 64, 2048)])
def test_gemm_plain(ops, M, N, K):
    torch.manual_seed(1)
    a = bf(torch.randn(M, K, device=DEV))
    w = bf(torch.randn(N, K, device=DEV) / math.sqrt(K))
    b = torch.randn(N, device=DEV)
    of = torch.empty(M, N, device=DEV)
    ob = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, bias=b, out_f32=of, out_bf16=ob)
    ref = a.float() @ w.float().t() + b
    # f32 accumulate of bf16 products: error bounded by summation order only
    assert (of - ref).abs().max().item() < 1e-3
    assert (ob.float() - ref).abs().max().item() < 4e-2


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 144, 160), (4096, 432, 144), (1000, 1152, 576),
                                   (4096, 32, 256), (333, 64, 2048), (65536, 256, 64), (300, 16, 64)])
@pytest.mark.parametrize("mode", ["f32", "bf16", "inplace", "direct"])
def test_gemm_store_paths(ops, M, N, K, mode):
    """TMA-store epilogues (f32, bf16, in-place reduce-add) against the row-per-thread fallback and torch."""
    torch.manual_seed(11)
    a = bf(torch.randn(M, K, device=DEV))
    w = bf(torch.randn(N, K, device=DEV) / math.sqrt(K))
    b = torch.randn(N, device=DEV)
    ref = a.float() @ w.float().t() + b
    if mode == "f32":
        o = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm(a, w, bias=b, out_f32=o)
        assert (o - ref).abs().max().item() < 1e-3
    elif mode == "bf16":
        o = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        ops.gemm(a, w, bias=b, act=2, out_bf16=o)
        assert (o.float() - F.gelu(ref)).abs().max().item() < 4e-2
    elif mode == "inplace":
        x = torch.randn(M, N, device=DEV)
        x0 = x.clone()
        ops.gemm(a, w, bias=b, residual=x, out_f32=x)
        assert (x - (x0 + ref)).abs().max().item() < 1e-3
    else:
        o = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm(a, w, bias=b, out_f32=o, impl=2)
        assert (o - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(512, 32, 64), (1000, 64, 144), (4096 + 128, 192, 576), (16384, 288, 288),
                                   (2048, 576, 2304), (4096, 2304, 576), (65536, 256, 64), (777, 96, 200)])
@pytest.mark.parametrize("mode", ["bf16_gelu", "f32", "inplace", "gamma_res", "two_out", "relu_bf16"])
def test_gemm_cta_pair_kernel_vs_single_cta_kernel(ops, M, N, K, mode):
    """gemm2 (tcgen05 cta_group::2, 256-row tiles on CTA pairs, 16 epilogue warps; impl 3) against the single-CTA kernel
    (impl 4) and torch, every epilogue / store path, ragged M / N / K."""
    torch.manual_seed(21)
    a = bf(torch.randn(M, K, device=DEV))
    w = bf(torch.randn(N, K, device=DEV) / math.sqrt(K))
    b = torch.randn(N, device=DEV)
    ref = a.float() @ w.float().t() + b
    g = torch.rand(N, device=DEV) + 0.5
    r = torch.randn(256, N, device=DEV)
    outs = []
    for impl in (3, 4):
        if mode == "bf16_gelu":
            o = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
            ops.gemm(a, w, bias=b, act=2, out_bf16=o, impl=impl)
            want, tol = F.gelu(ref), 4e-2
        elif mode == "relu_bf16":
            o = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
            ops.gemm(a, w, bias=b, act=1, out_bf16=o, impl=impl)
            want, tol = F.relu(ref), 4e-2
        elif mode == "f32":
            o = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(a, w, bias=b, out_f32=o, impl=impl)
            want, tol = ref, 1e-3
        elif mode == "inplace":
            torch.manual_seed(5)
            o = torch.randn(M, N, device=DEV)
            want, tol = o.clone() + ref, 1e-3
            ops.gemm(a, w, bias=b, residual=o, out_f32=o, impl=impl)
        elif mode == "gamma_res":
            o = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(a, w, bias=b, act=2, gamma=g, residual=r, res_row_mod=256, out_f32=o, impl=impl)
            want, tol = F.gelu(ref) * g + r.repeat((M + 255) // 256, 1)[:M], 1e-3
        else:
            o = torch.full((M, N), float("nan"), device=DEV)
            o2 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
            ops.gemm(a, w, bias=b, out_f32=o, out_bf16=o2, impl=impl)
            assert (o2.float() - ref).abs().max().item() < 4e-2
            want, tol = ref, 1e-3
        assert (o.float() - want).abs().max().item() < tol, (impl, mode)
        outs.append(o.float())
    # same products accumulated in the same k order: the two kernels agree to fp32 rounding (usually bit for bit)
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()) + (8e-3 if "bf16" in mode else 0)


def test_gemm_bf16_store_into_column_slice(ops):
    torch.manual_seed(12)
    a = bf(torch.randn(500, 256, device=DEV))
    w = bf(torch.randn(512, 256, device=DEV) / 16)
    buf = torch.zeros(500, 768, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, out_bf16=buf[:, :512])
    ref = a.float() @ w.float().t()
    assert (buf[:, :512].float() - ref).abs().max().item() < 4e-2
    assert buf[:, 512:].abs().max().item() == 0


def test_gemm_epilogue_all(ops):
    torch.manual_seed(2)
    M, N, K = 1024, 256, 256
    a = bf(torch.randn(M, K, device=DEV))
    w = bf(torch.randn(N, K, device=DEV) / 16)
    b = torch.randn(N, device=DEV)
    g = torch.rand(N, device=DEV) + 0.5
    r = torch.randn(256, N, device=DEV)
    of = torch.empty(M, N, device=DEV)
    ops.gemm(a, w, bias=b, act=2, gamma=g, residual=r, res_row_mod=256, out_f32=of)
    ref = F.gelu(a.float() @ w.float().t() + b) * g + r.repeat(4, 1)
    assert (of - ref).abs().max().item() < 1e-3


def test_gemm_rope(ops):
    torch.manual_seed(3)
    B, T, K = 2, 256, 256
    a = bf(torch.randn(B * T, K, device=DEV))
    w = bf(torch.randn(768, K, device=DEV) / 16)
    ang = torch.rand(64, 128, device=DEV) * 6.28
    cs = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
    of = torch.empty(B * T, 768, device=DEV)
    ops.gemm(a, w, out_f32=of, rope=(cs.permute(1, 0, 2).contiguous(), 0, 512, T, T - 4))
    ref = a.float() @ w.float().t()
    rows = torch.arange(B * T, device=DEV) % T
    pos = rows % 64
    for c0 in (0, 256):
        x = ref[:, c0:c0 + 256].reshape(-1, 128, 2)
        c, s = cs[pos][..., 0], cs[pos][..., 1]
        rot = torch.stack([x[..., 0] * c - x[..., 1] * s, x[..., 0] * s + x[..., 1] * c], -1).reshape(-1, 256)
        ref[:, c0:c0 + 256] = torch.where((rows < T - 4)[:, None], rot, ref[:, c0:c0 + 256])
    assert (of - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("side,B,nptr,out", [(64, 2, 8, "bf16"), (32, 3, 4, "f32"), (64, 1, 0, "bf16")])
def test_gemm_rope_axial_equals_full_table_path(ops, side, B, nptr, out):
    """The rotary epilogue with the axial table staged in shared memory (product path) against the same GEMM with the
    full pair-major table read from global memory: identical bits, on the memory-attention shapes (rows of several
    stored frames per object, the pointer tokens at the end of every object's rows left unrotated)."""
    from detsam2_b200.engine import _rope_axial, _rope_table
    torch.manual_seed(33)
    T = side * side
    N_rows = 2 * T + nptr
    a = bf(torch.randn(B * N_rows, 64, device=DEV))
    w = bf(torch.randn(256, 64, device=DEV) / 8)
    bias = torch.randn(256, device=DEV)
    full, ax = _rope_table(256, side, 10000.0).to(DEV), _rope_axial(256, side, 10000.0).to(DEV)
    kw = dict(out_bf16=None, out_f32=None)
    outs = []
    # full table (single-CTA kernel only), axial table on the single-CTA kernel (impl 4) and on the CTA-pair kernel (impl 3)
    for tab, impl in ((full, 4), (ax, 4), (ax, 3)):
        o = torch.empty(B * N_rows, 256, device=DEV, dtype=torch.bfloat16 if out == "bf16" else torch.float32)
        kw = {"out_bf16": o} if out == "bf16" else {"out_f32": o}
        ops.gemm(a, w, bias=bias, rope=(tab, 0, 256, N_rows, N_rows - nptr), impl=impl, **kw)
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    assert (outs[2].float() - outs[1].float()).abs().max().item() <= (8e-3 if out == "bf16" else 1e-5)
    # and the SIMT debug kernel reads the axial table the same way
    o1 = torch.empty(B * N_rows, 256, device=DEV)
    ops.gemm(a, w, bias=bias, rope=(ax, 0, 256, N_rows, N_rows - nptr), out_f32=o1, impl=1)
    assert (o1 - outs[1].float()).abs().max().item() < (4e-2 if out == "bf16" else 1e-3)


def test_gemm_matches_simt_debug_kernel(ops):
    torch.manual_seed(4)
    a = bf(torch.randn(300, 96, device=DEV))
    w = bf(torch.randn(80, 96, device=DEV))
    o0 = torch.empty(300, 80, device=DEV)
    o1 = torch.empty(300, 80, device=DEV)
    ops.gemm(a, w, out_f32=o0)
    ops.gemm(a, w, out_f32=o1, impl=1)
    assert (o0 - o1).abs().max().item() < 1e-3


def test_gemm_rejects_cpu_and_misaligned(ops):
    from detsam2_b200.capi import Ds2Error
    a = bf(torch.randn(8, 64))
    with pytest.raises(Ds2Error):
        ops.gemm(a, a, out_f32=torch.empty(8, 8))
    a = bf(torch.randn(8, 68, device=DEV))[:, :60]
    with pytest.raises(Ds2Error):
        ops.gemm(a, bf(torch.randn(8, 60, device=DEV)), out_f32=torch.empty(8, 8, device=DEV))


# ------------------------------------------------------------------------------------------ flash
@pytest.mark.parametrize("B,Lq,Lk,DV,impl,qmul", [
    (1, 128, 128, 64, 0, 1.0), (2, 300, 517, 64, 0, 1.0), (2, 256, 3000, 64, 0, 6.0),
    (2, 512, 1000, 256, 0, 1.0), (1, 256, 3000, 256, 0, 6.0),
    (2, 512, 1000, 64, 3, 1.0), (2, 256, 3000, 64, 3, 6.0), (3, 4096, 4 + 2 * 4096, 64, 0, 1.0),
    # impl 2 = Q as a shared-memory operand (impl 0 keeps Q in TMEM); odd key counts so the masked tail
    # falls into either column half of the split-row softmax (impl 9)
    (2, 300, 517, 64, 2, 1.0), (2, 512, 1000, 256, 2, 1.0), (2, 256, 3000, 64, 2, 6.0),
    (1, 200, 128 + 3, 64, 9, 1.0), (1, 200, 128 + 67, 64, 9, 3.0), (1, 130, 64 + 5, 256, 9, 1.0),
    (1, 130, 64 + 37, 256, 9, 3.0), (2, 300, 517, 64, 9, 1.0), (2, 512, 1000, 256, 9, 1.0),
    (2, 256, 3000, 64, 9, 6.0), (1, 200, 128 + 67, 64, 0, 3.0), (1, 130, 64 + 37, 256, 0, 3.0),
    ])
def test_flash(ops, B, Lq, Lk, DV, impl, qmul):
    torch.manual_seed(5)
    q = bf(qmul * torch.randn(B, Lq, 256, device=DEV))
    k = bf(torch.randn(B, Lk, 256, device=DEV))
    v = bf(torch.randn(B, Lk, DV, device=DEV))
    o = torch.empty(B, Lq, DV, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn(q, k, v, o, 1.0 / 16, impl=impl)
    ref = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
    # P is rounded to bf16 before P·V (as in any flash kernel) and the output is bf16
    assert not torch.isnan(o.float()).any()
    assert (o.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,Lq,Lk", [(2, 300, 2100), (3, 512, 16 * 128 + 5), (16, 4096, 4 + 7 * 4096), (1, 4096, 28736),
                                     (2, 256, 517), (5, 1024, 4100)])
def test_flash_two_key_halves(ops, B, Lq, Lk):
    """DV = 64 with a workspace: every item is two key halves combined in a fixed order.  Must agree with fp32 SDPA, and
    the bits must not depend on how the launch schedules an item — whole (one CTA runs both halves), split (two CTAs,
    the last to arrive combines), the launch's own choice (partial last wave split), or the batch size."""
    torch.manual_seed(15)
    q = bf(torch.randn(B, Lq, 256, device=DEV))
    k = bf(torch.randn(B, Lk, 256, device=DEV))
    v = bf(torch.randn(B, Lk, 64, device=DEV))
    ws = torch.zeros(ops.flash_workspace_bytes(B, Lq, 64), dtype=torch.uint8, device=DEV)
    outs = []
    for flags in (0, 1, 2):
        o = torch.full((B, Lq, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.flash_attn(q, k, v, o, 1.0 / 16, workspace=ws, flags=flags)
        outs.append(o)
    ref = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
    assert not torch.isnan(outs[0].float()).any()
    assert (outs[0].float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # the arrival counters are back at zero: the workspace can be reused by the next launch as it is
    n_items = B * ((Lq + 127) // 128)
    assert int(ws[:4 * n_items].view(torch.int32).abs().sum().item()) == 0
    # batch independence, bit for bit
    o1 = torch.empty(1, Lq, 64, device=DEV, dtype=torch.bfloat16)
    ws1 = torch.zeros(ops.flash_workspace_bytes(1, Lq, 64), dtype=torch.uint8, device=DEV)
    last = B - 1
    ops.flash_attn(q[last:].contiguous(), k[last:].contiguous(), v[last:].contiguous(), o1, 1.0 / 16, workspace=ws1)
    assert torch.equal(o1[0], outs[0][last])


def test_flash_strided_k(ops):
    torch.manual_seed(6)
    B, Lq, Lk = 2, 256, 700
    q = bf(torch.randn(B, Lq, 256, device=DEV))
    kall = bf(torch.randn(B, Lk, 1024, device=DEV))
    v = bf(torch.randn(B, Lk, 64, device=DEV))
    o = torch.empty(B, Lq, 64, device=DEV, dtype=torch.bfloat16)
    k = kall[:, :, 512:768]
    ops.flash_attn(q, k, v, o, 1.0 / 16)
    ref = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
    assert (o.float() - ref).abs().max().item() < 2e-2


# ------------------------------------------------------------------------------------------ mha
def _mha_ref(q, k, v, scale, Lk_valid=None):
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if Lk_valid is not None:
        s[..., Lk_valid:] = float("-inf")
    p = s.softmax(-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float())


@pytest.mark.parametrize("B,H,D,Lq,Lk,Lkv", [(2, 8, 16, 9, 4096, 0), (2, 8, 16, 4096, 16, 9), (3, 8, 32, 8, 16, 8),
                                              # the mask decoder's shapes on the cluster-split / thread-per-query kernels
                                              (16, 8, 16, 8, 4096, 0), (3, 8, 16, 16, 300, 290), (1, 8, 16, 1, 257, 0),
                                              (3, 8, 16, 4096, 9, 0), (2, 8, 16, 1000, 1, 0), (2, 8, 16, 300, 16, 0),
                                              (1, 8, 72, 1024, 1024, 0), (1, 2, 96, 100, 130, 0),
                                              # head_dim in (64, 80], Lq / Lk multiples of 128: tcgen05 flash loop
                                              (2, 4, 72, 256, 384, 0), (1, 2, 80, 128, 128, 0), (1, 8, 72, 4096, 4096, 0)])
def test_mha_dense(ops, B, H, D, Lq, Lk, Lkv):
    torch.manual_seed(7)
    q = bf(torch.randn(B, Lq, H, D, device=DEV))
    k = bf(torch.randn(B, Lk, H, D, device=DEV))
    v = bf(torch.randn(B, Lk, H, D, device=DEV))
    o = torch.zeros(B, Lq, H, D, device=DEV, dtype=torch.bfloat16)
    ops.mha(q, k, v, o, heads=H, head_dim=D, scale=D ** -0.5, B=B, Lq=Lq, Lk=Lk, Lk_valid=Lkv,
            strides=(H * D, H * D, H * D, H * D, Lq * H * D, Lk * H * D, Lk * H * D, Lq * H * D))
    ref = _mha_ref(q, k, v, D ** -0.5, Lkv or None)
    assert (o.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("B,H,D,T", [(2, 8, 72, 512), (1, 8, 72, 4096), (3, 2, 80, 256), (1, 5, 72, 128)])
def test_mha_global_hiera_layout(ops, monkeypatch, mode, B, H, D, T):
    """Hiera's global-attention blocks (hieradet.py:57-82, window_size 0) in the engine's layout: q, k, v are column
    sections of one token-major qkv matrix.  DS2_GLOB_FLASH = 1 (default) / 2: the flash kernel's multi-head variant (TMA boxes of
    64 columns that run past the head — and, for the last head, past the section — into data that must not matter) with
    one / two softmax threads per row; 0: the serial-chain kernel.  All against fp32 softmax attention; the two flash
    variants must also agree with each other to bf16 rounding."""
    monkeypatch.setenv("DS2_GLOB_FLASH", str(mode))
    torch.manual_seed(11)
    do = H * D
    qkv = bf(torch.randn(B * T, 3 * do, device=DEV))
    # poison everything a careless box could pick up beyond the q / k / v sections of interest
    out = torch.full((B * T, do), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], out, heads=H, head_dim=D, scale=D ** -0.5, B=B, Lq=T, Lk=T,
            strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, T * do))
    q, k, v = (qkv[:, i * do:(i + 1) * do].reshape(B, T, H, D) for i in range(3))
    ref = _mha_ref(q, k, v, D ** -0.5)
    got = out.view(B, T, H, D).float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 2e-2


def _window_partition(x, w):
    B, H, W, C = x.shape
    ph, pw = (w - H % w) % w, (w - W % w) % w
    x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // w, w, Wp // w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, w, w, C)
    return x, (Hp, Wp)


def _window_unpartition(win, w, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = win.shape[0] // (Hp * Wp // w // w)
    x = win.view(B, Hp // w, Wp // w, w, w, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W]


@pytest.mark.parametrize("Hm,w,heads,D,pool", [(32, 8, 2, 72, 0), (32, 8, 2, 72, 1), (16, 4, 4, 72, 0),
                                                (16, 4, 2, 72, 1), (32, 16, 4, 72, 0), (32, 14, 2, 56, 0),
                                                (32, 14, 2, 56, 1), (16, 7, 2, 96, 0),
                                                # 4x4 / 8x8 windows without padding: the whole-window-per-CTA kernel
                                                (16, 8, 1, 96, 0), (16, 4, 2, 56, 1), (32, 4, 8, 72, 1), (64, 8, 4, 72, 1),
                                                (16, 8, 2, 56, 0), (24, 4, 4, 72, 0),
                                                # 16x16 windows with head_dim in (64, 80]: the tcgen05 kernel
                                                (64, 16, 8, 72, 0), (32, 16, 3, 80, 0), (16, 16, 1, 72, 0)])
def test_mha_window(ops, Hm, w, heads, D, pool):
    _mha_window_case(ops, Hm, w, heads, D, pool)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("Hm,heads,D", [(64, 8, 72), (32, 3, 80), (16, 1, 72), (48, 4, 72)])
def test_mha_window16_on_flash_variant(ops, monkeypatch, mode, Hm, heads, D):
    """DS2_WIN_FLASH: the 16 x 16 windows as two 128-row items each of the flash kernel's multi-head variant (key tiles are
    TMA boxes {64 columns, 16 x, 8 y} of the token raster) — same reference as the window kernel."""
    monkeypatch.setenv("DS2_WIN_FLASH", str(mode))
    _mha_window_case(ops, Hm, 16, heads, D, 0)


def _mha_window_case(ops, Hm, w, heads, D, pool):
    """qkv token-major [B, Hm*Wm, 3*heads*D] exactly as MultiScaleAttention consumes it
    (hieradet.py:57-82), including zero-pad windows whose pad tokens carry the qkv bias."""
    torch.manual_seed(8)
    B, Wm = 2, Hm
    C = heads * D
    bias = bf(torch.randn(3 * C, device=DEV) * 0.3)
    qkv = bf(torch.randn(B, Hm * Wm, 3 * C, device=DEV))
    Ho = Hm // 2 if pool else Hm
    o = torch.zeros(B, Ho * Ho, C, device=DEV, dtype=torch.bfloat16)
    ops.mha(qkv, qkv[:, :, C:], qkv[:, :, 2 * C:], o, heads=heads, head_dim=D, scale=D ** -0.5, B=B,
            strides=(3 * C, 3 * C, 3 * C, C, Hm * Wm * 3 * C, Hm * Wm * 3 * C, Hm * Wm * 3 * C, Ho * Ho * C),
            window=w, Hm=Hm, Wm=Wm, q_pool=pool, pad=(bias, bias[C:], bias[2 * C:]))
    # reference: partition, pad tokens = bias, attention, unpartition
    x = qkv.float().view(B, Hm, Wm, 3 * C)
    ph = (w - Hm % w) % w
    xp = bias.float().view(1, 1, 1, -1).expand(B, Hm + ph, Wm + ph, 3 * C).clone()
    xp[:, :Hm, :Wm] = x
    win, pad_hw = _window_partition(xp, w)
    nW = win.shape[0]
    t = win.reshape(nW, w * w, 3, heads, D)
    q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    ww = w
    if pool:
        q = F.max_pool2d(q.reshape(nW, w, w, C).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        ww = w // 2
        q = q.reshape(nW, ww * ww, heads, D)
        pad_hw = (pad_hw[0] // 2, pad_hw[1] // 2)
    r = _mha_ref(q, k, v, D ** -0.5).reshape(nW, ww, ww, C)
    ref = _window_unpartition(r, ww, pad_hw, (Ho, Ho)).reshape(B, Ho * Ho, C)
    assert (o.float() - ref).abs().max().item() < 2e-2


# ------------------------------------------------------------------------------------------ LN etc.
@pytest.mark.parametrize("rows,C,eps", [(1000, 144, 1e-6), (77, 256, 1e-5), (513, 1152, 1e-6), (64, 64, 1e-6)])
def test_layernorm(ops, rows, C, eps):
    torch.manual_seed(9)
    x = torch.randn(rows, C, device=DEV) * 3 + 1
    w, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    pos = torch.randn(16, C, device=DEV)
    of = torch.empty(rows, C, device=DEV)
    ob = torch.empty(rows, C, device=DEV, dtype=torch.bfloat16)
    o2 = torch.empty(rows, C, device=DEV, dtype=torch.bfloat16)
    ops.layernorm(x, w, b, eps, out_f32=of, out_bf16=ob, pos=pos, pos_row_mod=16, out2_bf16=o2)
    ref = F.layer_norm(x, (C,), w, b, eps)
    assert (of - ref).abs().max().item() < 2e-5 * max(1, ref.abs().max().item())
    assert (ob.float() - ref).abs().max().item() < 2e-2 * max(1, ref.abs().max().item())
    ref2 = ref + pos.repeat((rows + 15) // 16, 1)[:rows]
    assert (o2.float() - ref2).abs().max().item() < 2e-2 * max(1, ref2.abs().max().item())
    ops.layernorm(x, w, b, eps, out_f32=of, act=2)
    assert (of - F.gelu(ref)).abs().max().item() < 2e-5 * max(1, ref.abs().max().item())
    xb = bf(x)
    ops.layernorm(xb, w, b, eps, out_f32=of)
    assert (of - F.layer_norm(xb.float(), (C,), w, b, eps)).abs().max().item() < 1e-4 * max(1, ref.abs().max().item())


def test_elementwise(ops):
    torch.manual_seed(10)
    a = torch.randn(512, 256, device=DEV)
    b = torch.randn(64, 256, device=DEV)
    of = torch.empty_like(a)
    ob = torch.empty(512, 256, device=DEV, dtype=torch.bfloat16)
    ops.axpby(a, b, 1.0, 0.1, b_row_mod=64, out_f32=of, out_bf16=ob)
    ref = a + 0.1 * b.repeat(8, 1)
    # the kernel contracts a*alpha + b*beta into FMAs: equal to torch up to one rounding of 0.1*b
    assert (of - ref).abs().max().item() < 1e-6
    assert (ob.float() - ref).abs().max().item() < 2e-2
    y = torch.empty(1001, device=DEV, dtype=torch.bfloat16)
    x = torch.randn(1001, device=DEV)
    ops.cast_f32_bf16(x, y)
    assert torch.equal(y, bf(x))
    z = torch.empty(1001, device=DEV)
    ops.cast_bf16_f32(y, z)
    assert torch.equal(z, y.float())
    x = torch.randn(2, 16, 12, 32, device=DEV)
    y = torch.empty(2, 8, 6, 32, device=DEV)
    ops.maxpool2x2(x, y, 2, 16, 12, 32)
    assert torch.equal(y, F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1))
    top, lat = torch.randn(1, 8, 8, 64, device=DEV), torch.randn(1, 16, 16, 64, device=DEV)
    out = torch.empty_like(lat)
    ops.upsample2x_add(top, lat, out, 1, 8, 8, 64)
    ref = lat + F.interpolate(top.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out, ref)


def test_bank_gather(ops):
    torch.manual_seed(11)
    B, T, C = 3, 256, 64
    mem = bf(torch.randn(B, T, C, device=DEV))
    pos, tpos = torch.randn(T, C, device=DEV), torch.randn(C, device=DEV)
    N = 2 * T + 8
    kin = torch.zeros(B, N, C, device=DEV, dtype=torch.bfloat16)
    val = torch.zeros_like(kin)
    ops.bank_gather(mem, pos, tpos, kin, val, B, T, C, N * C, T)
    assert torch.equal(val[:, T:2 * T], mem)
    assert torch.equal(kin[:, T:2 * T], bf(mem.float() + (pos + tpos)))
    ptr, tp = torch.randn(B, 256, device=DEV), torch.randn(64, device=DEV)
    ops.bank_ptr(ptr, tp, kin, val, B, N * C, 2 * T + 4)
    assert torch.equal(val[:, 2 * T + 4:], bf(ptr).view(B, 4, 64))
    assert torch.equal(kin[:, 2 * T + 4:], bf(ptr.view(B, 4, 64) + tp))


# ------------------------------------------------------------------------------------------ convs
def test_im2col_patch_matches_conv(ops):
    torch.manual_seed(12)
    S, E = 64, 48
    img = torch.randn(3, S, S, device=DEV).half()
    wt = torch.randn(E, 3, 7, 7, device=DEV) / 12
    cols = torch.empty((S // 4) ** 2, 160, device=DEV, dtype=torch.bfloat16)
    ops.im2col_patch(img, cols, S, 160)
    wp = torch.zeros(E, 160, device=DEV, dtype=torch.bfloat16)
    wp[:, :147] = bf(wt.reshape(E, 147))
    out = torch.empty((S // 4) ** 2, E, device=DEV)
    ops.gemm(cols, wp, out_f32=out)
    ref = F.conv2d(bf(img.float()).float()[None], bf(wt).float(), stride=4, padding=3)[0].permute(1, 2, 0).reshape(-1, E)
    assert (out - ref).abs().max().item() < 1e-3


def test_im2col_k3s2_matches_conv(ops):
    torch.manual_seed(13)
    B, Hi, C, Co = 2, 32, 16, 64
    x = bf(torch.randn(B, Hi, Hi, C, device=DEV))
    wt = torch.randn(Co, C, 3, 3, device=DEV) / 12
    cols = torch.empty(B * (Hi // 2) ** 2, 9 * C, device=DEV, dtype=torch.bfloat16)
    ops.im2col_k3s2(x, cols, B, Hi, Hi, C)
    wp = bf(wt.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous())
    out = torch.empty(cols.shape[0], Co, device=DEV)
    ops.gemm(cols, wp, out_f32=out)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), bf(wt).float(), stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("B,Hm,C", [(2, 16, 64), (1, 18, 256), (3, 7, 32), (2, 64, 256), (1, 9, 48), (1, 33, 96)])
def test_dwconv7(ops, B, Hm, C):
    torch.manual_seed(14)
    x = torch.randn(B, Hm, Hm, C, device=DEV)
    w, b = torch.randn(C, 1, 7, 7, device=DEV) / 7, torch.randn(C, device=DEV)
    y = torch.empty_like(x)
    ops.dwconv7(x, w.reshape(C, 49).contiguous(), b, y, B, Hm, Hm, C)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=3, groups=C).permute(0, 2, 3, 1)
    assert (y - ref).abs().max().item() < 1e-4


def _ln2d(x, w, b):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + 1e-6)) + b[:, None, None]


@pytest.mark.parametrize("binarize", [0, 1])
def test_maskds_stage1_and_conv(ops, binarize):
    torch.manual_seed(15)
    B, Sl = 2, 32
    low = torch.randn(B, Sl, Sl, device=DEV) * 4
    w1, b1 = torch.randn(4, 1, 3, 3, device=DEV) / 3, torch.randn(4, device=DEV) * 0.1
    g1, be1 = torch.rand(4, device=DEV) + 0.5, torch.randn(4, device=DEV) * 0.1
    out1 = torch.empty(B, 2 * Sl, 2 * Sl, 4, device=DEV, dtype=torch.bfloat16)
    ops.maskds_stage1(low, B, Sl, binarize, 20.0, -10.0, w1, b1, g1, be1, out1)
    hi = F.interpolate(low[:, None], size=(4 * Sl, 4 * Sl), mode="bilinear", align_corners=False)
    m = ((hi > 0).float() if binarize else torch.sigmoid(hi)) * 20.0 - 10.0
    ref1 = F.gelu(_ln2d(F.conv2d(m, w1, b1, stride=2, padding=1), g1, be1)).permute(0, 2, 3, 1)
    if binarize:
        # a hi-res pixel within rounding distance of 0 may binarise differently; allow a handful
        bad = ((out1.float() - ref1).abs() > 3e-2).float().mean().item()
        assert bad < 1e-3
    else:
        assert (out1.float() - ref1).abs().max().item() < 3e-2
    w2, b2 = torch.randn(16, 4, 3, 3, device=DEV) / 6, torch.randn(16, device=DEV) * 0.1
    g2, be2 = torch.rand(16, device=DEV) + 0.5, torch.randn(16, device=DEV) * 0.1
    out2 = torch.empty(B, Sl, Sl, 16, device=DEV, dtype=torch.bfloat16)
    ops.maskds_conv(out1, B, 2 * Sl, 2 * Sl, 4, 16, w2, b2, g2, be2, out2)
    ref2 = F.gelu(_ln2d(F.conv2d(out1.float().permute(0, 3, 1, 2), w2, b2, stride=2, padding=1), g2, be2)).permute(0, 2, 3, 1)
    assert (out2.float() - ref2).abs().max().item() < 3e-2


# ------------------------------------------------------------------------------------------ decoder tail
def test_upscale_chain(ops):
    torch.manual_seed(16)
    B, Hm = 2, 8
    src = torch.randn(B, 256, Hm, Hm, device=DEV)
    ct1 = torch.nn.ConvTranspose2d(256, 64, 2, 2).to(DEV)
    ct2 = torch.nn.ConvTranspose2d(64, 32, 2, 2).to(DEV)
    s1 = torch.randn(1, 64, 2 * Hm, 2 * Hm, device=DEV)
    s0 = torch.randn(1, 32, 4 * Hm, 4 * Hm, device=DEV)
    lw, lb = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV) * 0.1
    hyper = torch.randn(B, 4, 32, device=DEV)
    with torch.no_grad():
        u1 = F.gelu(_ln2d(ct1(src) + s1, lw, lb))
        u2 = F.gelu(ct2(u1) + s0)
        ref = (hyper @ u2.view(B, 32, -1)).view(B, 4, 4 * Hm, 4 * Hm)
        # ours: ConvT as GEMM with W[(dy*2+dx)*Cout + c][k]
        a = bf(src.permute(0, 2, 3, 1).reshape(-1, 256))
        w1 = bf(ct1.weight.permute(2, 3, 1, 0).reshape(4 * 64, 256).contiguous())
        g1 = torch.empty(a.shape[0], 256, device=DEV)
        ops.gemm(a, w1, out_f32=g1)
        y1 = torch.empty(B, 2 * Hm, 2 * Hm, 64, device=DEV, dtype=torch.bfloat16)
        ops.upscale1(g1, ct1.bias.detach(), s1[0].permute(1, 2, 0).reshape(-1, 64).contiguous(), lw, lb, y1, B, Hm, Hm, 64)
        assert (y1.float() - u1.permute(0, 2, 3, 1)).abs().max().item() < 5e-2
        w2 = bf(ct2.weight.permute(2, 3, 1, 0).reshape(4 * 32, 64).contiguous())
        g2 = torch.empty(B * 4 * Hm * Hm, 128, device=DEV)
        ops.gemm(y1.view(-1, 64), w2, out_f32=g2)
        masks = torch.empty(B, 4, 4 * Hm, 4 * Hm, device=DEV)
        ops.upscale2_masks(g2, ct2.bias.detach(), s0[0].permute(1, 2, 0).reshape(-1, 32).contiguous(), hyper, masks,
                           B, 2 * Hm, 2 * Hm, 32, 4)
    assert (masks - ref).abs().max().item() < 0.15  # bf16 operands through two transposed convs


def test_mlp3(ops):
    torch.manual_seed(17)
    B, n = 5, 4
    x = torch.randn(B * n, 256, device=DEV)
    w1, b1 = torch.randn(n, 256, 256, device=DEV) / 16, torch.randn(n, 256, device=DEV) * 0.1
    w2, b2 = torch.randn(n, 256, 256, device=DEV) / 16, torch.randn(n, 256, device=DEV) * 0.1
    w3, b3 = torch.randn(n, 32, 256, device=DEV) / 16, torch.randn(n, 32, device=DEV) * 0.1
    y = torch.empty(B * n, 32, device=DEV)
    tr = lambda w: w.transpose(-1, -2).contiguous()  # the kernel takes input-major weights [set, in, out]
    ops.mlp3(x, tr(w1), b1, tr(w2), b2, tr(w3), b3, y, rows=B * n, nmlp=n, sigmoid_out=True)
    xs = x.view(B, n, 256)
    h = torch.relu(torch.einsum("bnk,nok->bno", xs, w1) + b1)
    h = torch.relu(torch.einsum("bnk,nok->bno", h, w2) + b2)
    ref = torch.sigmoid(torch.einsum("bnk,nok->bno", h, w3) + b3).reshape(B * n, 32)
    assert (y - ref).abs().max().item() < 1e-4
    gather = torch.tensor([3, 3, 0, 7], device=DEV, dtype=torch.int32)
    y2 = torch.empty(4, 32, device=DEV)
    ops.mlp3(x, tr(w1[:1]), b1[:1], tr(w2[:1]), b2[:1], tr(w3[:1]), b3[:1], y2, rows=4, nmlp=1, gather=gather)
    # narrow heads (IoU: 4 outputs, object score: 1): the K range is sliced over the spare threads
    for dout in (4, 1):
        w3n, b3n = torch.randn(1, dout, 256, device=DEV) / 16, torch.randn(1, dout, device=DEV) * 0.1
        y3 = torch.empty(B * n, dout, device=DEV)
        ops.mlp3(x, tr(w1[:1]), b1[:1], tr(w2[:1]), b2[:1], tr(w3n), b3n, y3, rows=B * n, nmlp=1)
        hh = torch.relu(torch.relu(x @ w1[0].t() + b1[0]) @ w2[0].t() + b2[0])
        assert (y3 - (hh @ w3n[0].t() + b3n[0])).abs().max().item() < 1e-4
    # object pointer head: 256 outputs, 16 and 64 rows (one and four row blocks of the cluster kernel)
    for rows in (16, 64):
        xr = torch.randn(rows, 256, device=DEV)
        w3w, b3w = torch.randn(1, 256, 256, device=DEV) / 16, torch.randn(1, 256, device=DEV) * 0.1
        y4 = torch.empty(rows, 256, device=DEV)
        ops.mlp3(xr, tr(w1[:1]), b1[:1], tr(w2[:1]), b2[:1], tr(w3w), b3w, y4, rows=rows, nmlp=1)
        hh = torch.relu(torch.relu(xr @ w1[0].t() + b1[0]) @ w2[0].t() + b2[0])
        assert (y4 - (hh @ w3w[0].t() + b3w[0])).abs().max().item() < 1e-4
    xg = x[gather.long()]
    h = torch.relu(xg @ w1[0].t() + b1[0])
    h = torch.relu(h @ w2[0].t() + b2[0])
    assert (y2 - (h @ w3[0].t() + b3[0])).abs().max().item() < 1e-4


@pytest.mark.parametrize("multimask", [0, 1])
def test_sam_select(ops, multimask):
    torch.manual_seed(18)
    B, S, C = 6, 32, 256
    masks = torch.randn(B, 4, S, S, device=DEV)
    masks[0, 0] = masks[0, 0].sign() * 5        # stable single mask
    masks[1, 0] = masks[1, 0] * 0.01            # unstable -> fallback
    ious = torch.rand(B, 4, device=DEV)
    score = torch.tensor([1.0, 2.0, -1.0, 0.5, 3.0, -0.2], device=DEV)
    toks = torch.randn(B, 4, C, device=DEV)
    low = torch.empty(B, S, S, device=DEV)
    iou_o = torch.empty(B, device=DEV)
    idx = torch.empty(B, device=DEV, dtype=torch.int32)
    tok = torch.empty(B, C, device=DEV)
    ops.sam_select(masks, ious, score, toks, B, S, C, multimask, 0.05, 0.98, low, iou_o, idx, tok)
    best = 1 + ious[:, 1:].argmax(-1)
    if multimask:
        exp_idx = best
        exp_tok = toks[torch.arange(B), best]
    else:
        m0 = masks[:, 0].flatten(1)
        ai, au = (m0 > 0.05).sum(-1).float(), (m0 > -0.05).sum(-1).float()
        stab = torch.where(au > 0, ai / au, torch.ones_like(au))
        exp_idx = torch.where(stab >= 0.98, torch.zeros_like(best), best)
        exp_tok = toks[:, 0]
    assert torch.equal(idx.long(), exp_idx)
    exp_low = masks[torch.arange(B), exp_idx]
    exp_low = torch.where((score > 0)[:, None, None], exp_low, torch.full_like(exp_low, -1024.0))
    assert torch.equal(low, exp_low)
    assert torch.equal(tok, exp_tok)
    assert torch.equal(iou_o, ious[torch.arange(B), exp_idx])
    ptr = torch.randn(B, C, device=DEV)
    nop = torch.randn(C, device=DEV)
    exp = torch.where((score > 0)[:, None], ptr, nop[None].expand(B, C))
    ops.objptr_mix(ptr, score, nop, B, C)
    assert torch.equal(ptr, exp)


# ------------------------------------------------------------------------------------------ post
def test_resize_bilinear(ops):
    torch.manual_seed(19)
    x = torch.randn(3, 64, 64, device=DEV)
    for Ho, Wo in ((256, 256), (90, 160), (64, 64), (33, 47)):
        y = torch.empty(3, Ho, Wo, device=DEV)
        ops.resize_bilinear(x, y, 3, 64, 64, Ho, Wo)
        ref = F.interpolate(x[:, None], size=(Ho, Wo), mode="bilinear", align_corners=False)[:, 0]
        assert (y - ref).abs().max().item() < 1e-5


def test_threshold_pack(ops):
    x = torch.randn(1003, device=DEV)
    bits = torch.zeros((1003 + 7) // 8, device=DEV, dtype=torch.uint8)
    ops.threshold_pack(x, bits)
    import numpy as np
    exp = np.packbits((x > 0).cpu().numpy(), bitorder="little")
    assert (bits.cpu().numpy() == exp).all()


def test_downsample4_aa_matches_torch_antialias(ops):
    """F.interpolate(x*20-10, 1/4, bilinear, antialias=True) (sam2_base.py:407-413), borders included."""
    torch.manual_seed(22)
    x = (torch.rand(3, 64, 64, device=DEV) > 0.6).float()
    y = torch.empty(3, 16, 16, device=DEV)
    ops.downsample4_aa(x, y, 20.0, -10.0)
    ref = F.interpolate((x * 20.0 - 10.0)[:, None], size=(16, 16), mode="bilinear", align_corners=False, antialias=True)[:, 0]
    assert (y - ref).abs().max().item() < 1e-4


def test_mask_prompt_embed_matches_conv_stack(ops):
    """mask_downsample (k4 s4) + PromptEncoder.mask_downscaling up to its 1x1 conv (prompt_encoder.py:52-60)."""
    torch.manual_seed(23)
    B, S = 2, 128
    m = (torch.rand(B, 1, S, S, device=DEV) > 0.5).float()
    wds, bds = torch.randn(1, 1, 4, 4, device=DEV) / 4, torch.randn(1, device=DEV)
    w0, b0 = torch.randn(4, 1, 2, 2, device=DEV) / 2, torch.randn(4, device=DEV)
    g0, be0 = torch.rand(4, device=DEV) + 0.5, torch.randn(4, device=DEV) * 0.1
    w3, b3 = torch.randn(16, 4, 2, 2, device=DEV) / 4, torch.randn(16, device=DEV)
    g3, be3 = torch.rand(16, device=DEV) + 0.5, torch.randn(16, device=DEV) * 0.1
    out = torch.empty(B * (S // 16) ** 2, 16, device=DEV, dtype=torch.bfloat16)
    ops.mask_prompt_embed(m.view(B, S, S).contiguous(), wds.reshape(-1).contiguous(), bds, w0.reshape(4, 4).contiguous(), b0, g0, be0,
                          w3.reshape(16, 16).contiguous(), b3, g3, be3, out)
    x = F.conv2d(m, wds, bds, stride=4)
    x = F.gelu(_ln2d_b(F.conv2d(x, w0, b0, stride=2), g0, be0))
    x = F.gelu(_ln2d_b(F.conv2d(x, w3, b3, stride=2), g3, be3))
    ref = x.permute(0, 2, 3, 1).reshape(-1, 16)
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


def _ln2d_b(x, w, b):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[None, :, None, None] * ((x - u) / torch.sqrt(s + 1e-6)) + b[None, :, None, None]


@pytest.mark.parametrize("N,H,W", [(3, 64, 256), (2, 37, 100), (16, 256, 1024)])
def test_mask_pack_stats_exact(ops, N, H, W):
    """Integer path: packed bits == numpy.packbits(little) per row, stats == exact raw moments (cv2.moments m00/m10/m01)."""
    torch.manual_seed(21)
    x = torch.randn(N, 1, H, W, device=DEV)
    x[0, 0, : H // 2] = -1.0           # empty region
    if N > 1:
        x[1] = -1.0                    # an object with an empty mask: area 0, sums 0
    bits, stats = ops.mask_pack_stats(x)
    m = (x[:, 0] > 0).cpu().numpy()
    ref_bits = np.packbits(m, axis=-1, bitorder="little")
    assert np.array_equal(bits.cpu().numpy(), ref_bits)
    ys, xs = np.mgrid[0:H, 0:W]
    ref = np.stack([m.sum((1, 2)), (m * xs).sum((1, 2)), (m * ys).sum((1, 2))], 1).astype(np.int64)
    assert np.array_equal(stats.cpu().numpy(), ref)
    # run twice: atomics on integers are order-independent
    bits2, stats2 = ops.mask_pack_stats(x)
    assert torch.equal(bits, bits2) and torch.equal(stats, stats2)


def test_connected_components_and_fill_holes(ops):
    """Integer work: bit-exact against the C oracle (oracle/cc_oracle.c) on seeded masks."""
    import numpy as np
    from oracle import cc_oracle
    rng = np.random.default_rng(0)
    N, H, W = 5, 64, 48
    m = (rng.random((N, 1, H, W)) < np.array([0.1, 0.45, 0.6, 0.9, 0.0])[:, None, None, None]).astype(np.uint8)
    m[4, 0, 10:20, 10:30] = 1
    m[4, 0, 13:15, 15:18] = 0
    labels, counts = ops.connected_components(torch.from_numpy(m).to(DEV))
    el, ec = cc_oracle.connected_components(m)
    assert (counts.cpu().numpy() == ec).all()
    assert (labels.cpu().numpy() == el).all()
    scores = torch.from_numpy(rng.standard_normal((N, H, W)).astype(np.float32)).to(DEV)
    exp = cc_oracle.fill_holes(scores.cpu().numpy(), 8)
    ws1 = torch.empty(N, H, W, device=DEV, dtype=torch.int32)
    ws2 = torch.empty_like(ws1)
    ops.fill_holes(scores, ws1, ws2, N, H, W, 8)
    assert (scores.cpu().numpy() == exp).all()
    with pytest.raises(Exception):
        ops.connected_components(torch.zeros(1, 1, 5, 4, dtype=torch.uint8, device=DEV))


def test_bank_assemble_matches_per_frame_gather(ops):
    """Table-driven whole-bank assembly == the per-frame gather + per-pointer kernels it replaces."""
    import numpy as np
    torch.manual_seed(5)
    B, T, C, nf, npt = 3, 256, 64, 4, 5
    mems = [torch.randn(B, T, C, device=DEV).bfloat16() for _ in range(nf)]
    ptrs = [torch.randn(B, 256, device=DEV) for _ in range(npt)]
    tidx = [6, 0, 3, 5]
    dist = [0.0, 1 / 15, -3 / 15, 7 / 15, 1.0]
    pos = torch.randn(T, C, device=DEV)
    tpos = torch.randn(7, C, device=DEV)
    w = torch.randn(64, 256, device=DEV) / 16
    bias = torch.randn(64, device=DEV) * 0.1
    N = nf * T + 4 * npt
    kin_ref = torch.zeros(B, N, C, device=DEV, dtype=torch.bfloat16)
    val_ref = torch.zeros_like(kin_ref)
    for f in range(nf):
        ops.bank_gather(mems[f], pos, tpos[tidx[f]], kin_ref, val_ref, B, T, C, N * C, f * T)
    for j in range(npt):
        ops.bank_ptr_pe(ptrs[j], dist[j], w, bias, kin_ref, val_ref, B, N * C, nf * T + 4 * j)
    tab_f = torch.tensor([m.data_ptr() for m in mems], dtype=torch.int64, device=DEV)
    tab_t = torch.tensor(tidx, dtype=torch.int32, device=DEV)
    tab_p = torch.tensor([p.data_ptr() for p in ptrs], dtype=torch.int64, device=DEV)
    tab_d = torch.tensor(dist, dtype=torch.float32, device=DEV)
    kin = torch.zeros_like(kin_ref)
    val = torch.zeros_like(kin_ref)
    ops.bank_assemble(tab_f.data_ptr(), tab_t.data_ptr(), nf, tab_p.data_ptr(), tab_d.data_ptr(), npt, pos, tpos, w, bias,
                      kin, val, B, T, C)
    assert torch.equal(kin, kin_ref) and torch.equal(val, val_ref)
    # reference semantics (sam2_base.py:564-650) in fp32
    exp = mems[1].float() + pos[None] + tpos[0][None, None]
    assert (kin[:, T:2 * T].float() - exp).abs().max().item() < 0.04
