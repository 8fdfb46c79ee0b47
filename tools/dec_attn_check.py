"""Accuracy of the decoder attention kernels against an fp64 reference on the same bf16 inputs (relative to the output scale)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from detsam2_b200 import ops
torch.manual_seed(0)
dev = "cuda"
for (B, H, D, Lq, Lk, qs) in [(2, 8, 16, 10, 1024, 1.0), (2, 8, 16, 8, 4096, 1.0), (2, 8, 16, 9, 4096, 3.0), (2, 8, 16, 1024, 10, 1.0), (2, 8, 16, 4096, 8, 3.0)]:
    q = (torch.randn(B, Lq, H, D, device=dev) * qs).bfloat16()
    k = (torch.randn(B, Lk, H, D, device=dev) * qs).bfloat16()
    v = torch.randn(B, Lk, H, D, device=dev).bfloat16()
    o = torch.zeros(B, Lq, H, D, device=dev, dtype=torch.bfloat16)
    ops.mha(q, k, v, o, heads=H, head_dim=D, scale=D ** -0.5, B=B, Lq=Lq, Lk=Lk,
            strides=(H * D, H * D, H * D, H * D, Lq * H * D, Lk * H * D, Lk * H * D, Lq * H * D))
    s = torch.einsum("bqhd,bkhd->bhqk", q.double(), k.double()) * D ** -0.5
    ref = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), v.double())
    err = (o.double() - ref)
    print(f"B{B} H{H} D{D} Lq{Lq} Lk{Lk} qscale{qs}: rms err / rms ref = {err.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():.5f}, "
          f"max err / max ref = {err.abs().max().item() / ref.abs().max().item():.5f}")
