"""Isolated timing of the hot kernel shapes of the large model (back-to-back launches, CUDA events), and a
one-launch-per-shape mode for `ncu --set full`.

usage: python tools/kernel_probe.py [--iters 50] [--once] [--only NAME_SUBSTR[,NAME_SUBSTR...]]
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--once", action="store_true", help="one launch per case (for ncu)")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from detsam2_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)
    cases = []

    def gemm_case(name, M, N, K, out="bf16", res=False, act=0, rope=False, rows_per_batch=None):
        a = torch.randn(M, K, device=dev).to(BF16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
        b = torch.randn(N, device=dev)
        o = torch.zeros(M, N, device=dev, dtype=BF16 if out == "bf16" else F32)
        kw = dict(bias=b, act=act)
        kw["out_bf16" if out == "bf16" else "out_f32"] = o
        if res:
            kw["residual"] = o
        if rope:
            from detsam2_b200.engine import _rope_axial, _rope_table
            cs = (_rope_axial if rope == "axial" else _rope_table)(256, 64, 10000.0).to(dev)
            rpb = rows_per_batch or 4096
            kw["rope"] = (cs, 0, min(N, 512), rpb, rpb - (rpb % 4096))
        cases.append((name, lambda: ops.gemm(a, w, **kw), 2.0 * M * N * K, None))

    def cublas_case(name, M, N, K):
        a = torch.randn(M, K, device=dev).to(BF16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
        o = torch.zeros(M, N, device=dev, dtype=BF16)
        cases.append((name, lambda: torch.matmul(a, w.t(), out=o), 2.0 * M * N * K, None))

    gemm_case("s3.qkv   4096x1728x576", 4096, 1728, 576)
    gemm_case("s3.proj  4096x576x576 +res", 4096, 576, 576, out="f32", res=True)
    gemm_case("s3.fc1   4096x2304x576 gelu", 4096, 2304, 576, act=2)
    gemm_case("s3.fc2   4096x576x2304 +res", 4096, 576, 2304, out="f32", res=True)
    # the same Hiera stage-3 block inside the 4-frame encoder pass (M = 4 x 4096 rows), with the epilogue varied, and
    # cuBLAS (torch.matmul, plain bf16 store, no epilogue) on the same shapes as the library bar
    gemm_case("e4.qkv   16384x1728x576", 16384, 1728, 576)
    gemm_case("e4.proj  16384x576x576 +res", 16384, 576, 576, out="f32", res=True)
    gemm_case("e4.fc1   16384x2304x576 gelu", 16384, 2304, 576, act=2)
    gemm_case("e4.fc1   16384x2304x576 relu", 16384, 2304, 576, act=1)
    gemm_case("e4.fc1   16384x2304x576 none", 16384, 2304, 576, act=0)
    gemm_case("e4.fc2   16384x576x2304 +res", 16384, 576, 2304, out="f32", res=True)
    gemm_case("e4.fc2   16384x576x2304 bf16", 16384, 576, 2304)
    cublas_case("cublas   16384x1728x576", 16384, 1728, 576)
    cublas_case("cublas   16384x576x576", 16384, 576, 576)
    cublas_case("cublas   16384x2304x576", 16384, 2304, 576)
    cublas_case("cublas   16384x576x2304", 16384, 576, 2304)
    cublas_case("cublas   65536x2048x256", 65536, 2048, 256)
    cublas_case("cublas   65536x256x2048", 65536, 256, 2048)
    cublas_case("cublas   8192x8192x8192", 8192, 8192, 8192)
    gemm_case("s2.fc1   16384x1152x288 gelu", 16384, 1152, 288, act=2)
    gemm_case("s2.fc2   16384x288x1152 +res", 16384, 288, 1152, out="f32", res=True)
    gemm_case("s4.fc2   1024x1152x4608 +res", 1024, 1152, 4608, out="f32", res=True)
    gemm_case("ma.sa_qkv 65536x768x256 rope", 65536, 768, 256, rope=True)
    gemm_case("ma.ca_k  459776x256x64 rope", 459776, 256, 64, rope=True, rows_per_batch=28736)
    gemm_case("ma.ca_k  459776x256x64 rope axial", 459776, 256, 64, rope="axial", rows_per_batch=28736)
    gemm_case("ma.sa_qkv 65536x768x256 rope axial", 65536, 768, 256, rope="axial")
    gemm_case("ma.ca_k  459776x256x64 norope", 459776, 256, 64)
    gemm_case("ma.ff1   65536x2048x256 relu", 65536, 2048, 256, act=1)
    gemm_case("ma.ff2   65536x256x2048 +res", 65536, 256, 2048, out="f32", res=True)
    gemm_case("ma.out   65536x256x256 +res", 65536, 256, 256, out="f32", res=True)
    gemm_case("big      8192x8192x8192", 8192, 8192, 8192)

    def mha_case(name, T, do, heads, window, Hm, pool=0, B=1, env=None):
        qkv = torch.randn(B * T, 3 * do, device=dev).to(BF16)
        Tq = T // 4 if pool else T
        att = torch.zeros(B * Tq, do, device=dev, dtype=BF16)
        hd = do // heads
        Lk = window * window if window else T
        Lq = (Lk // 4 if pool else Lk) if window else Tq
        nseq = (Hm // window) ** 2 if window else 1
        fl = 4.0 * B * nseq * heads * Lq * Lk * hd

        def run():
            if env:
                os.environ.update(env)      # kernel selection switches that the library reads per call
            ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=hd,
                    scale=1.0 / math.sqrt(hd), B=B, Lq=Tq if window == 0 else 0,
                    Lk=T if window == 0 else 0,
                    strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, Tq * do),
                    window=window, Hm=Hm, Wm=Hm, q_pool=pool)
        cases.append((name, run, fl, None))

    mha_case("mha s3 win16 8h", 4096, 576, 8, 16, 64)
    mha_case("mha s3 global 8h", 4096, 576, 8, 0, 64)
    for mode, nm in ((0, "window kernel"), (1, "flash variant, 1 softmax thread/row"), (2, "flash variant, 2 softmax threads/row")):
        mha_case(f"mha s3 win16 8h x4 frames ({nm})", 4096, 576, 8, 16, 64, B=4, env={"DS2_WIN_FLASH": str(mode)})
    for mode, nm in ((0, "serial-chain kernel"), (1, "flash variant, 1 softmax thread/row"), (2, "flash variant, 2 softmax threads/row")):
        mha_case(f"mha s3 global 8h x4 frames ({nm})", 4096, 576, 8, 0, 64, B=4, env={"DS2_GLOB_FLASH": str(mode), "DS2_GLOB_DBG": "0"})
    mha_case("mha s3 global 8h x4 frames (flash variant, first TMA box only: wrong results)", 4096, 576, 8, 0, 64, B=4,
             env={"DS2_GLOB_FLASH": "2", "DS2_GLOB_DBG": "5"})
    mha_case("mha s2 win4 4h", 16384, 288, 4, 4, 128)
    mha_case("mha s1 win8 2h", 65536, 144, 2, 8, 256)

    x = torch.randn(4096, 576, device=dev)
    g, b_ = torch.ones(576, device=dev), torch.zeros(576, device=dev)
    xo = torch.zeros(4096, 576, device=dev, dtype=BF16)
    cases.append(("ln 4096x576", lambda: ops.layernorm(x, g, b_, 1e-6, out_bf16=xo), 0.0, 4096 * 576 * 6))
    x2 = torch.randn(65536, 256, device=dev)
    g2, b2 = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    xo2 = torch.zeros(65536, 256, device=dev, dtype=BF16)
    cases.append(("ln 65536x256", lambda: ops.layernorm(x2, g2, b2, 1e-5, out_bf16=xo2), 0.0, 65536 * 256 * 6))

    xd = torch.randn(16, 64, 64, 256, device=dev)
    wd, bd = torch.randn(256, 49, device=dev) / 7, torch.randn(256, device=dev)
    yd = torch.empty_like(xd)
    cases.append(("dwconv7 16x64x64x256", lambda: ops.dwconv7(xd, wd, bd, yd, 16, 64, 64, 256), 2.0 * 49 * xd.numel(), None))
    hs = torch.randn(16 * 8, 256, device=dev)
    mw = [torch.randn(4, 256, 256, device=dev) / 16, torch.randn(4, 256, device=dev), torch.randn(4, 256, 256, device=dev) / 16,
          torch.randn(4, 256, device=dev), torch.randn(4, 256, 32, device=dev) / 16, torch.randn(4, 32, device=dev)]
    ym = torch.empty(64, 32, device=dev)
    cases.append(("mlp3 hyper 64 items", lambda: ops.mlp3(hs, *mw, ym, rows=64, nmlp=4), 2.0 * 64 * (2 * 65536 + 8192), None))
    q = torch.randn(16, 4096, 256, device=dev).to(BF16)
    k = torch.randn(16, 28736, 256, device=dev).to(BF16)
    v = torch.randn(16, 28736, 64, device=dev).to(BF16)
    o = torch.zeros(16, 4096, 64, device=dev, dtype=BF16)
    cases.append(("flash cross B16 N28736", lambda: ops.flash_attn(q, k, v, o, 1 / 16.0),
                  2.0 * 16 * 4096 * 28736 * 320, None))
    fws = torch.zeros(ops.flash_workspace_bytes(16, 4096, 64), dtype=torch.uint8, device=dev)
    for fl, nm in ((0, "two key halves, tail wave split"), (1, "two key halves, whole items")):
        cases.append((f"flash cross B16 N28736 workspace ({nm})",
                      lambda fl=fl: ops.flash_attn(q, k, v, o, 1 / 16.0, workspace=fws, flags=fl), 2.0 * 16 * 4096 * 28736 * 320, None))
    for im, nm in ((9, "2 softmax wg"), (2, "Q in smem"), (6, "no exps"), (5, "half of each K tile loaded"),
                   (13, "MMA pipeline only, softmax relays")):
        cases.append((f"flash cross B16 N28736 impl{im} ({nm})", lambda im=im: ops.flash_attn(q, k, v, o, 1 / 16.0, impl=im),
                      2.0 * 16 * 4096 * 28736 * 320, None))
    qs = torch.randn(16, 4096, 768, device=dev).to(BF16)
    os_ = torch.zeros(16, 4096, 256, device=dev, dtype=BF16)
    cases.append(("flash self B16", lambda: ops.flash_attn(qs[:, :, :256], qs[:, :, 256:512], qs[:, :, 512:], os_, 1 / 16.0),
                  2.0 * 16 * 4096 * 4096 * 512, None))

    cases.append(("flash self B16 impl9 (2 softmax wg)", lambda: ops.flash_attn(qs[:, :, :256], qs[:, :, 256:512], qs[:, :, 512:], os_, 1 / 16.0, impl=9),
                  2.0 * 16 * 4096 * 4096 * 512, None))
    lines = []
    for name, fn, flops, nbytes in cases:
        if args.only and not any(o in name for o in args.only.split(",")):
            continue
        if args.once:
            fn()
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # the launches are captured into one graph so the host's ctypes launch rate cannot bound the timing
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(args.iters):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        extra = f"{flops / us / 1e6:8.1f} TFLOP/s" if flops else f"{nbytes / us / 1e3:8.1f} GB/s"
        lines.append(f"{name:36s} {us:9.2f} us  {extra}")
        print(lines[-1], flush=True)
    if lines:
        with open(os.path.join(ROOT, "gpurun_out", "kernel_probe.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
