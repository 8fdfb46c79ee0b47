"""Small-shape launches of the hand-rolled mbarrier / TMEM / TMA kernels for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitizer_cases.py
    compute-sanitizer --tool racecheck python tools/sanitizer_cases.py
    compute-sanitizer --tool synccheck python tools/sanitizer_cases.py

Every case also checks its result against torch, so a sanitizer-clean run is a run of correct kernels.
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

BF16 = torch.bfloat16


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""     # "mha": the Hiera attention cases alone
    from detsam2_b200 import ops
    from detsam2_b200.engine import _rope_axial
    dev = "cuda"
    torch.manual_seed(0)
    done = []

    def gemm_case(M, N, K, impl, **kw):
        a = torch.randn(M, K, device=dev).to(BF16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
        b = torch.randn(N, device=dev)
        ref = a.float() @ w.float().t() + b
        o = torch.zeros(M, N, device=dev)
        ops.gemm(a, w, bias=b, out_f32=o, impl=impl, **kw)
        assert (o - ref).abs().max().item() < 1e-3
        ob = torch.zeros(M, N, device=dev, dtype=BF16)
        ops.gemm(a, w, bias=b, act=2, out_bf16=ob, impl=impl)
        assert (ob.float() - F.gelu(ref)).abs().max().item() < 4e-2
        x = torch.randn(M, N, device=dev)
        x0 = x.clone()
        ops.gemm(a, w, bias=b, residual=x, out_f32=x, impl=impl)
        assert (x - (x0 + ref)).abs().max().item() < 1e-3
        done.append(f"gemm impl {impl} {M}x{N}x{K}")

    for impl in (() if only == "mha" else (4, 3)):           # single-CTA kernel, CTA-pair kernel
        gemm_case(600, 192, 160, impl)
        gemm_case(1024, 320, 576, impl)
    # rotary epilogue with the table staged in shared memory
    side, Bq = 32, 2
    T = side * side
    a = torch.randn(Bq * (T + 8), 64, device=dev).to(BF16)
    w = (torch.randn(256, 64, device=dev) / 8).to(BF16)
    outs = []
    for impl in (() if only == "mha" else (4, 3)):
        o = torch.zeros(Bq * (T + 8), 256, device=dev, dtype=BF16)
        ops.gemm(a, w, out_bf16=o, rope=(_rope_axial(256, side, 10000.0).to(dev), 0, 256, T + 8, T), impl=impl)
        outs.append(o.float())
    if outs:
        assert (outs[0] - outs[1]).abs().max().item() < 8e-3
        done.append("gemm rotary epilogue (both kernels)")

    # flash attention: plain, two key halves (whole items / every item on two CTAs), self-attention width
    for B, Lq, Lk, DV, flags in () if only == "mha" else ((2, 300, 700, 64, None), (2, 256, 16 * 128 + 5, 64, 1), (2, 256, 16 * 128 + 5, 64, 2),
                                 (1, 256, 17 * 128, 64, 0), (2, 200, 333, 256, None)):
        q = torch.randn(B, Lq, 256, device=dev).to(BF16)
        k = torch.randn(B, Lk, 256, device=dev).to(BF16)
        v = torch.randn(B, Lk, DV, device=dev).to(BF16)
        o = torch.zeros(B, Lq, DV, device=dev, dtype=BF16)
        ws = None if flags is None else torch.zeros(max(ops.flash_workspace_bytes(B, Lq, DV), 16), dtype=torch.uint8, device=dev)
        ops.flash_attn(q, k, v, o, 1.0 / 16, workspace=ws, flags=flags or 0)
        ref = F.scaled_dot_product_attention(q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
        assert (o.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
        done.append(f"flash B{B} Lq{Lq} Lk{Lk} DV{DV} flags {flags}")

    # Hiera attention kernels: tcgen05 window-16 / global, mma.sync windows
    def mha_case(T, do, heads, window, Hm, pool=0, B=1, env=None):
        for k_, v_ in (env or {}).items():
            os.environ[k_] = v_              # kernel selection switches the library reads per call
        qkv = torch.randn(B * T, 3 * do, device=dev).to(BF16)
        Tq = T // 4 if pool else T
        att = torch.zeros(B * Tq, do, device=dev, dtype=BF16)
        hd = do // heads
        ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=hd, scale=1.0 / math.sqrt(hd), B=B,
                Lq=Tq if window == 0 else 0, Lk=T if window == 0 else 0,
                strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, Tq * do), window=window, Hm=Hm, Wm=Hm,
                q_pool=pool)
        assert torch.isfinite(att.float()).all()
        if window == 0:                      # global attention: checked against fp32 softmax attention
            q, k, v = (qkv[:, i * do:(i + 1) * do].reshape(B, T, heads, hd).float() for i in range(3))
            p_ = (torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(hd)).softmax(-1)
            ref = torch.einsum("bhqk,bkhd->bqhd", p_, v).reshape(B * T, do)
            assert (att.float() - ref).abs().max().item() < 2e-2
        for k_ in (env or {}):
            os.environ.pop(k_)
        done.append(f"mha B{B} T{T} d{do} h{heads} win{window} pool{pool} {env or ''}")

    mha_case(1024, 576, 8, 16, 32)       # win16_attn_tc
    mha_case(1024, 576, 8, 0, 32)        # flash kernel, multi-head variant (global blocks)
    mha_case(512, 576, 8, 0, 32, B=2, env={"DS2_GLOB_FLASH": "2"})    # same, two softmax threads per row
    mha_case(1024, 576, 8, 0, 32, env={"DS2_GLOB_FLASH": "0"})        # glob_attn_tc (serial-chain kernel)
    mha_case(1024, 576, 8, 16, 32, B=2, env={"DS2_WIN_FLASH": "1"})   # window mode of the multi-head variant (rank-4 TMA boxes)
    mha_case(1024, 288, 4, 4, 32)        # mma.sync windows
    mha_case(1024, 288, 4, 8, 32, pool=1)

    # stencil / integer kernels
    x = torch.randn(2, 16, 16, 256, device=dev)
    wd, bd = torch.randn(256, 49, device=dev) / 7, torch.randn(256, device=dev)
    y = torch.empty_like(x)
    ops.dwconv7(x, wd, bd, y, 2, 16, 16, 256)
    ref = F.conv2d(x.permute(0, 3, 1, 2), wd.view(256, 1, 7, 7), bd, padding=3, groups=256).permute(0, 2, 3, 1)
    assert (y - ref).abs().max().item() < 1e-3
    done.append("dwconv7")
    m = (torch.rand(3, 1, 64, 64, device=dev) < 0.55).to(torch.uint8)
    labels, counts = ops.connected_components(m)
    assert int(counts.max()) > 0
    done.append("connected components")
    torch.cuda.synchronize()
    print("sanitizer cases run:", len(done))
    for d in done:
        print("  ", d)


if __name__ == "__main__":
    main()
