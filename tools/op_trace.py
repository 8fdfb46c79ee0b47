"""Per-op timing of one tracked frame: every C-ABI launch of the engine is bracketed with CUDA events
(graphs off, so each launch is its own stream op) and aggregated by (op, shape key).  Unlike the ncu
launch list (cold L2, serialised), these are warm-L2 times on the launching stream.

usage: python tools/op_trace.py [--model large] [--objects 16] [--prefill 8] [--steps 3] [--out gpurun_out/op_trace.txt]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DS2_GRAPHS"] = "0"

import torch  # noqa: E402


def key_of(name, args, kw):
    if name == "gemm":
        a, w = args[0], args[1]
        out = "f32" if kw.get("out_f32") is not None else "bf16"
        res = "+res" if kw.get("residual") is not None else ""
        rope = "+rope" if kw.get("rope") is not None else ""
        return f"M={a.shape[0]} N={w.shape[0]} K={a.shape[1]} act={kw.get('act', 0)} {out}{res}{rope}"
    if name == "flash_attn":
        q, k, v = args[0], args[1], args[2]
        return f"B={q.shape[0]} Lq={q.shape[1]} Lk={k.shape[1]} DV={v.shape[2]}"
    if name == "mha":
        return (f"H={kw['heads']} D={kw['head_dim']} B={kw['B']} Lq={kw.get('Lq', 0)} Lk={kw.get('Lk', 0)} "
                f"win={kw.get('window', 0)} Hm={kw.get('Hm', 0)} pool={kw.get('q_pool', 0)}")
    if name == "layernorm":
        x = args[0]
        outs = "".join(t for t, k in (("f", "out_f32"), ("b", "out_bf16"), ("B", "out2_bf16")) if kw.get(k) is not None)
        return f"rows={x.shape[0]} C={x.shape[1]} in={str(x.dtype)[6:]} out={outs}"
    for a in args:
        if torch.is_tensor(a):
            return "x" + "x".join(str(s) for s in a.shape)
    return ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="large")
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--prefill", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_trace.txt"))
    args = ap.parse_args()
    from detsam2_b200 import ops
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.synthetic import BilliardVideo
    yaml = {"tiny": "t", "small": "s", "base_plus": "b+", "large": "l"}[args.model]
    dev = torch.device("cuda", 0)
    pred = build_sam2_video_predictor(f"configs/sam2.1/sam2.1_hiera_{yaml}.yaml", device=dev, seed=0, feature_cache_frames=1)
    pred.encoder_overlap = False     # per-op times: one stream, nothing running beside the op being timed
    S = pred.cfg.image_size
    vid = BilliardVideo(num_objects=args.objects, height=S, width=S, num_frames=2 + args.prefill + args.steps, seed=0)
    st = pred.init_state(list(vid.frames()), offload_video_to_cpu=False)
    for oid, box in vid.boxes(0).items():
        pred.add_new_points_or_box(st, 0, oid, box=box)
    gen = pred.propagate_in_video(st)
    for _ in range(1 + args.prefill):
        next(gen)
    torch.cuda.synchronize()

    records = []
    names = [n for n in dir(ops) if not n.startswith("_") and callable(getattr(ops, n)) and n not in ("launch_count",)
             and getattr(getattr(ops, n), "__module__", "") == ops.__name__]
    orig = {}

    def wrap(n, f):
        def g(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*a, **kw)
            e1.record()
            records.append((n, key_of(n, a, kw), e0, e1))
            return r
        return g

    for n in names:
        orig[n] = getattr(ops, n)
        setattr(ops, n, wrap(n, orig[n]))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        next(gen)
    t1.record()
    torch.cuda.synchronize()
    for n in names:
        setattr(ops, n, orig[n])
    agg = collections.OrderedDict()
    for n, k, e0, e1 in records:
        d = agg.setdefault((n, k), [0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1) * 1e3
    tot = sum(v[1] for v in agg.values())
    lines = [f"# op trace: {args.model}, {args.objects} objects, {args.steps} steps, eager launches; sum of op times "
             f"{tot / args.steps / 1e3:.3f} ms/step (wall incl. host gaps {t0.elapsed_time(t1) / args.steps:.3f} ms/step)",
             f"{'op':12s} {'key':70s} {'n/step':>7s} {'avg us':>9s} {'us/step':>10s} {'share':>6s}"]
    by_op = collections.defaultdict(float)
    for (n, k), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n:12s} {k:70s} {c / args.steps:7.1f} {us / c:9.1f} {us / args.steps:10.1f} {100 * us / tot:5.1f}%")
        by_op[n] += us
    lines.append("# by op")
    for n, us in sorted(by_op.items(), key=lambda kv: -kv[1]):
        lines.append(f"{n:12s} {us / args.steps:10.1f} us/step {100 * us / tot:5.1f}%")
    txt = "\n".join(lines)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
