"""Records every ds2_gemm call of one steady-state tracked frame (large, B objects), then times each
distinct shape in isolation with CUDA events.  Writes gpurun_out/gemm_shapes.json and prints a table.

usage: python tools/gemm_shapes.py [objects]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from detsam2_b200 import ops  # noqa: E402
from detsam2_b200.build_sam import build_sam2_video_predictor  # noqa: E402
from detsam2_b200.synthetic import BilliardVideo  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    pred = build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_l.yaml", device="cuda", seed=0)
    vid = BilliardVideo(num_objects=B, height=1024, width=1024, num_frames=10, seed=0)
    frames = list(vid.frames())
    calls = []
    orig = ops.gemm

    def rec(a, w, **kw):
        calls.append((a.shape[0], w.shape[0], a.shape[1],
                      "f32" if kw.get("out_f32") is not None else "", "bf16" if kw.get("out_bf16") is not None else "",
                      "res" if kw.get("residual") is not None else "", kw.get("act", 0),
                      "rope" if kw.get("rope") is not None else ""))
        return orig(a, w, **kw)

    with torch.inference_mode():
        st = pred.init_state(frames, offload_video_to_cpu=False)
        for oid, box in vid.boxes(0).items():
            pred.add_new_points_or_box(st, 0, oid, box=box)
        gen = pred.propagate_in_video(st)
        for _ in range(9):
            next(gen)
        torch.cuda.synchronize()
        ops.gemm = rec
        import detsam2_b200.engine as E
        E.ops.gemm = rec
        next(gen)
        torch.cuda.synchronize()
        ops.gemm = orig
        E.ops.gemm = orig
    from collections import Counter
    cnt = Counter(calls)
    rows = []
    for key, n in cnt.items():
        M, N, K, of, ob, res, act, rope = key
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        kw = {}
        if of:
            kw["out_f32"] = torch.empty(M, N, device="cuda")
        if ob:
            kw["out_bf16"] = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        if res:
            kw["residual"] = torch.randn(M, N, device="cuda")
        kw["bias"] = torch.randn(N, device="cuda")
        kw["act"] = act
        for _ in range(3):
            orig(a, w, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            orig(a, w, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        fl = 2.0 * M * N * K
        byt = M * K * 2 + N * K * 2 + M * N * ((4 if of else 0) + (2 if ob else 0) + (4 if res else 0))
        rows.append({"M": M, "N": N, "K": K, "out": of + ob, "res": bool(res), "act": act, "rope": bool(rope), "count": n,
                     "us": us, "tflops": fl / us / 1e6, "gbs": byt / us / 1e3, "us_total": us * n})
    rows.sort(key=lambda r: -r["us_total"])
    tot = sum(r["us_total"] for r in rows)
    print(f"{len(calls)} gemm calls / frame, {len(rows)} distinct, sum of isolated times {tot / 1e3:.2f} ms")
    for r in rows:
        print(f"M {r['M']:7d} N {r['N']:5d} K {r['K']:5d} {r['out']:8s} res {int(r['res'])} act {r['act']} rope {int(r['rope'])} "
              f"x{r['count']:3d}  {r['us']:8.1f} us  {r['tflops']:7.1f} TF/s {r['gbs']:7.0f} GB/s  total {r['us_total']:8.1f} us")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
