"""GPU bring-up of the tcgen05 kernels (GEMM, flash attention) against torch, one case per
subprocess so a device trap in one case cannot poison the others.  Writes gpurun_out/bringup_tc.json.

usage: python tools/bringup_tc.py            (driver)
       python tools/bringup_tc.py CASE_NAME  (single case, internal)
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _lib():
    from detsam2_b200 import capi
    lib = C.CDLL(capi.lib_path())
    lib.ds2_gemm.restype = C.c_int
    lib.ds2_gemm.argtypes = [C.POINTER(capi.GemmArgs), C.c_void_p]
    lib.ds2_flash_attn.restype = C.c_int
    lib.ds2_flash_attn.argtypes = [C.POINTER(capi.FlashArgs), C.c_void_p]
    lib.ds2_last_error.restype = C.c_char_p
    return lib, capi


def gemm_case(M, N, K, impl=0, bias=True, act=0, gamma=False, residual=False, res_mod=0, rope=False,
              out="both", lda_pad=0, iters=0):
    import torch
    torch.manual_seed(0)
    lib, capi = _lib()
    dev = "cuda"
    lda = K + lda_pad
    A = torch.randn(M, lda, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=dev) if bias else None
    g = torch.rand(N, device=dev) + 0.5 if gamma else None
    rrows = res_mod if res_mod > 0 else M
    R = torch.randn(rrows, N, device=dev) if residual else None
    of = torch.zeros(M, N, device=dev)
    ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    cs = None
    rows_per_batch = M
    period = 64
    if rope:
        ang = torch.rand(period, 128, device=dev) * 6.28
        cs = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
        rows_per_batch = max(M // 2, 1)
    a = capi.GemmArgs()
    a.A, a.W, a.lda, a.ldw = A.data_ptr(), W.data_ptr(), lda, K
    a.M, a.N, a.K = M, N, K
    a.bias = b.data_ptr() if bias else None
    a.gamma = g.data_ptr() if gamma else None
    a.residual = R.data_ptr() if residual else None
    a.ldr = N
    a.res_row_mod = res_mod
    a.act = act
    a.out_f32 = of.data_ptr() if out in ("both", "f32") else None
    a.out_bf16 = ob.data_ptr() if out in ("both", "bf16") else None
    a.ldc = N
    a.ldc_bf16 = N
    if rope:
        cs_t = cs.permute(1, 0, 2).contiguous()  # pair-major table
        a.rope_cs = cs_t.data_ptr()
        a.rope_col0, a.rope_col1 = 0, 256
        a.rope_period, a.rope_rows_per_batch = period, rows_per_batch
        a.rope_row_limit = rows_per_batch - 3
    a.impl = impl
    rc = lib.ds2_gemm(C.byref(a), None)
    torch.cuda.synchronize()
    if rc != 0:
        return {"ok": False, "rc": rc, "err": lib.ds2_last_error().decode()}
    ref = A[:, :K].float() @ W.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = ref.relu()
    if act == 2:
        ref = torch.nn.functional.gelu(ref)
    if gamma:
        ref = ref * g
    if rope:
        rows = torch.arange(M, device=dev)
        rb = rows % rows_per_batch
        pos = rb % period
        x = ref[:, :256].reshape(M, 128, 2)
        c, s = cs[pos][..., 0], cs[pos][..., 1]
        rot = torch.stack([x[..., 0] * c - x[..., 1] * s, x[..., 0] * s + x[..., 1] * c], -1).reshape(M, 256)
        m = (rb < rows_per_batch - 3)[:, None]
        ref = torch.cat([torch.where(m, rot, ref[:, :256]), ref[:, 256:]], 1)
    if residual:
        ref = ref + (R[torch.arange(M, device=dev) % res_mod] if res_mod > 0 else R)
    res = {"ok": True}
    if out in ("both", "f32"):
        err = (of - ref).abs().max().item()
        res["max_err_f32"] = err
        res["ok"] = res["ok"] and err < 2e-2 * max(1.0, ref.abs().max().item())
    if out in ("both", "bf16"):
        err = (ob.float() - ref).abs().max().item()
        res["max_err_bf16"] = err
        res["ok"] = res["ok"] and err < 3e-2 * max(1.0, ref.abs().max().item())
    res["ref_absmax"] = ref.abs().max().item()
    if iters:
        for _ in range(3):
            lib.ds2_gemm(C.byref(a), None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            lib.ds2_gemm(C.byref(a), None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # torch reference timing
        Wt = W.t().contiguous()
        for _ in range(3):
            torch.matmul(A[:, :K], Wt)
        e0.record()
        for _ in range(iters):
            torch.matmul(A[:, :K], Wt)
        e1.record()
        torch.cuda.synchronize()
        res["torch_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    return res


def flash_case(B, Lq, Lk, DV, impl=0, ldk_mult=1, iters=0, qmul=1.0):
    import torch
    torch.manual_seed(0)
    lib, capi = _lib()
    dev = "cuda"
    q = (qmul * torch.randn(B, Lq, 256, device=dev)).to(torch.bfloat16)
    kfull = torch.randn(B, Lk, 256 * ldk_mult, device=dev).to(torch.bfloat16)
    k = kfull[:, :, 256 * (ldk_mult - 1):]
    v = torch.randn(B, Lk, DV, device=dev).to(torch.bfloat16)
    o = torch.zeros(B, Lq, DV, device=dev, dtype=torch.bfloat16)
    a = capi.FlashArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo = 256, 256 * ldk_mult, DV, DV
    a.bsq, a.bsk, a.bsv, a.bso = Lq * 256, Lk * 256 * ldk_mult, Lk * DV, Lq * DV
    a.B, a.Lq, a.Lk, a.DV = B, Lq, Lk, DV
    a.scale = 1.0 / 16.0
    a.impl = impl
    rc = lib.ds2_flash_attn(C.byref(a), None)
    torch.cuda.synchronize()
    if rc != 0:
        return {"ok": False, "rc": rc, "err": lib.ds2_last_error().decode()}
    ref = torch.nn.functional.scaled_dot_product_attention(
        q.float()[:, None], k.float()[:, None], v.float()[:, None])[:, 0]
    err = (o.float() - ref).abs().max().item()
    res = {"ok": err < 2e-2, "max_err": err, "ref_absmax": ref.abs().max().item(),
           "nan": bool(torch.isnan(o.float()).any().item())}
    if iters:
        for _ in range(2):
            lib.ds2_flash_attn(C.byref(a), None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            lib.ds2_flash_attn(C.byref(a), None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops_exec"] = 2.0 * B * Lq * Lk * (256 + DV) / ms / 1e9
        res["tflops_nominal"] = 4.0 * B * Lq * Lk * 256 / ms / 1e9
        qb, kb, vb = q[:, None], k[:, None].contiguous(), v[:, None]
        if DV == 256:
            for _ in range(2):
                torch.nn.functional.scaled_dot_product_attention(qb, kb, vb)
            e0.record()
            for _ in range(iters):
                torch.nn.functional.scaled_dot_product_attention(qb, kb, vb)
            e1.record()
            torch.cuda.synchronize()
            res["torch_sdpa_ms"] = e0.elapsed_time(e1) / iters
    return res


CASES = {
    # name: (fn, kwargs)
    "gemm_simt_small": (gemm_case, dict(M=100, N=48, K=72, impl=1)),
    "gemm_tc_128x128x64": (gemm_case, dict(M=128, N=128, K=64, bias=False)),
    "gemm_tc_128x128x256": (gemm_case, dict(M=128, N=128, K=256)),
    "gemm_tc_256x256x512": (gemm_case, dict(M=256, N=256, K=512)),
    "gemm_tc_ragged": (gemm_case, dict(M=200, N=144, K=160, lda_pad=8)),
    "gemm_tc_small_m": (gemm_case, dict(M=16, N=256, K=256, act=1)),
    "gemm_tc_n432": (gemm_case, dict(M=4096, N=432, K=144, act=2)),
    "gemm_tc_n64_res": (gemm_case, dict(M=4096, N=64, K=256, residual=True, gamma=True)),
    "gemm_tc_resmod": (gemm_case, dict(M=1024, N=256, K=256, residual=True, res_mod=256)),
    "gemm_tc_rope": (gemm_case, dict(M=512, N=768, K=256, rope=True)),
    "gemm_tc_n32": (gemm_case, dict(M=4096, N=32, K=256)),
    "gemm_tc_many_tiles": (gemm_case, dict(M=65536, N=576, K=576, out="bf16", iters=10)),
    "gemm_tc_perf_8k": (gemm_case, dict(M=8192, N=8192, K=8192, bias=False, out="bf16", iters=5)),
    "gemm_tc_perf_mlp": (gemm_case, dict(M=65536, N=2048, K=256, act=1, out="bf16", iters=10)),
    "flash_simt": (flash_case, dict(B=1, Lq=128, Lk=300, DV=64, impl=1)),
    "flash_dv64_1tile": (flash_case, dict(B=1, Lq=128, Lk=128, DV=64)),
    "flash_dv64_q1": (flash_case, dict(B=2, Lq=256, Lk=1000, DV=64)),
    "flash_dv64_q1_ldk": (flash_case, dict(B=2, Lq=300, Lk=517, DV=64, ldk_mult=4)),
    "flash_dv64_q1_peaky": (flash_case, dict(B=2, Lq=256, Lk=3000, DV=64, qmul=6.0)),
    "flash_dv256_peaky": (flash_case, dict(B=2, Lq=256, Lk=3000, DV=256, qmul=6.0)),
    "flash_dv64_q2_small": (flash_case, dict(B=1, Lq=256, Lk=64, DV=64, impl=3)),
    "flash_dv64_q2": (flash_case, dict(B=2, Lq=512, Lk=1000, DV=64, impl=3)),
    "flash_dv64_q2_ldk": (flash_case, dict(B=2, Lq=300, Lk=517, DV=64, ldk_mult=4, impl=3)),
    "flash_dv64_q2_peaky": (flash_case, dict(B=2, Lq=256, Lk=3000, DV=64, qmul=6.0, impl=3)),
    "flash_dv256_small": (flash_case, dict(B=1, Lq=128, Lk=64, DV=256)),
    "flash_dv256": (flash_case, dict(B=2, Lq=512, Lk=1000, DV=256)),
    "flash_dv64_perf_q1": (flash_case, dict(B=16, Lq=4096, Lk=28736, DV=64, iters=3)),
    "flash_dv64_perf_q2": (flash_case, dict(B=16, Lq=4096, Lk=28736, DV=64, impl=3, iters=3)),
    "flash_dv256_perf": (flash_case, dict(B=16, Lq=4096, Lk=4096, DV=256, iters=5)),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        fn, kw = CASES[sys.argv[1]]
        print("RESULT " + json.dumps(fn(**kw)))
        return
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    out = {}
    for name in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True,
                               text=True, timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                out[name] = json.loads(line[-1][7:])
            else:
                out[name] = {"ok": False, "rc": r.returncode, "stdout": r.stdout[-800:], "stderr": r.stderr[-1500:]}
        except subprocess.TimeoutExpired:
            out[name] = {"ok": False, "timeout": True}
        out[name]["wall_s"] = round(time.time() - t0, 1)
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bringup_tc.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("SUMMARY ok=%d fail=%d" % (sum(1 for v in out.values() if v.get("ok")),
                                     sum(1 for v in out.values() if not v.get("ok"))))


if __name__ == "__main__":
    main()
