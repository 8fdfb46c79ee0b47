"""Phase timestamps of one CTA of the tcgen05 windowed-attention kernel (DS2_WIN_DBG=1)."""
import ctypes as C, math, os, sys
os.environ["DS2_WIN_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from detsam2_b200 import ops, capi
lib = capi.load()
T, do, heads = 4096, 576, 8
qkv = torch.randn(T, 3 * do, device="cuda").to(torch.bfloat16)
att = torch.zeros(T, do, device="cuda", dtype=torch.bfloat16)
def run():
    ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=72, scale=1 / math.sqrt(72), B=1, Lq=0, Lk=0,
            strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, T * do), window=16, Hm=64, Wm=64)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.ds2_debug_win_times(buf)
names = ["start", "pdl_sync", "loads returned", "Q,K staged (sync)", "S0 ready", "pass1 done", "pass2 done (P)", "O ready", "epilogue done",
         "MMA: QK issued", "MMA: V staged", "MMA: P0 ready", "MMA: P1 ready"]
for n, v in zip(names, list(buf)):
    print(f"{n:24s} {v:8d} clk")

# global attention (flash loop): timestamps of tile 4 of CTA (0, 0)
def run_glob():
    ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=72, scale=1 / math.sqrt(72), B=1, Lq=T, Lk=T,
            strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, T * do), window=0, Hm=64, Wm=64)
for _ in range(2):
    run_glob()
torch.cuda.synchronize()
lib.ds2_debug_win_times(buf)
for n, v in zip(["loop top", "PV(j-1) retired", "K,V stored + sync", "prefetch issued", "S ready", "pass 1 done", "pass 2 done"], list(buf)):
    print(f"glob {n:24s} {v:8d} clk")
