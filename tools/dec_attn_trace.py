"""Debug aid: run a golden scenario on the CUDA engine and check every non-windowed ds2_mha call against fp64 on the same inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DS2_GRAPHS"] = "0"
import torch
from detsam2_b200 import ops
from detsam2_b200.engine import CudaEngine
from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.weights import synthetic_state_dict
from oracle import scenarios

name = sys.argv[1] if len(sys.argv) > 1 else "points_api"
real = ops.mha
worst = {}

def mha(q, k, v, out, *, heads, head_dim, scale, B, Lq=0, Lk=0, strides, window=0, **kw):
    real(q, k, v, out, heads=heads, head_dim=head_dim, scale=scale, B=B, Lq=Lq, Lk=Lk, strides=strides, window=window, **kw)
    if window or head_dim != 16:
        return
    qt, kt, vt, ot, qb, kb, vb, ob = strides
    def view(t, L, ts, bs):
        return torch.as_strided(t, (B, L, heads, head_dim), (bs, ts, head_dim, 1), t.storage_offset()).double()
    qq, kk, vv, oo = view(q, Lq, qt, qb), view(k, Lk, kt, kb), view(v, Lk, vt, vb), view(out, Lq, ot, ob)
    s = torch.einsum("bqhd,bkhd->bhqk", qq, kk) * scale
    ref = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), vv)
    err = ((oo - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item()
    key = (B, Lq, Lk)
    if err > worst.get(key, (0,))[0]:
        worst[key] = (err, s.abs().max().item(), ref.abs().max().item(), bool(torch.isfinite(oo).all()))

ops.mha = mha
import detsam2_b200.engine as E
E.ops.mha = mha
cfg = scenarios.scenario_config(name)
pred = SAM2VideoPredictor(CudaEngine(cfg, synthetic_state_dict(cfg, 0), device="cuda:0"), fill_hole_area=0)
scenarios.SCENARIOS[name](pred)
torch.cuda.synchronize()
for k, v in sorted(worst.items()):
    print(f"B={k[0]} Lq={k[1]} Lk={k[2]}: worst rel-rms err {v[0]:.5f}  (max |scaled score| {v[1]:.1f}, max |out| {v[2]:.3f}, finite {v[3]})")
