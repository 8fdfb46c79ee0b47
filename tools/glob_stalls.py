"""Where the multi-head flash variant (Hiera global attention) waits: barrier-stall cycles of the MMA issuer and of a
softmax warp (DS2_GLOB_DBG=8 -> the kernel's stall accounting), per CTA per key tile."""
import ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from detsam2_b200 import ops, capi
BF16 = torch.bfloat16
lib = capi.load()
B, T, do, heads = 4, 4096, 576, 8
qkv = torch.randn(B * T, 3 * do, device="cuda").to(BF16)
att = torch.zeros(B * T, do, device="cuda", dtype=BF16)


def run():
    ops.mha(qkv, qkv[:, do:], qkv[:, 2 * do:], att, heads=heads, head_dim=72, scale=1 / math.sqrt(72), B=B, Lq=T, Lk=T,
            strides=(3 * do, 3 * do, 3 * do, do, T * 3 * do, T * 3 * do, T * 3 * do, T * do), window=0, Hm=64, Wm=64)


for mode in ("1", "2"):
    os.environ["DS2_GLOB_FLASH"] = mode
    os.environ["DS2_GLOB_DBG"] = "8"
    buf = (C.c_ulonglong * 16)()
    run(); torch.cuda.synchronize()
    lib.ds2_debug_flash_stalls(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib.ds2_debug_flash_stalls(buf, 1)
    b = list(buf); n = max(b[7], 1); tiles = T // 128
    print(f"mode {mode}: {e0.elapsed_time(e1)*1e3:.0f} us, {n} CTAs, {tiles} key tiles; per CTA per tile (clk): MMA warp total {b[3]/n/tiles:.0f} = "
          f"wait K {b[0]/n/tiles:.0f} + wait V {b[1]/n/tiles:.0f} + wait P {b[2]/n/tiles:.0f} + issue {(b[3]-b[0]-b[1]-b[2])/n/tiles:.0f}; "
          f"softmax warp total {b[6]/n/tiles:.0f} = wait S {b[4]/n/tiles:.0f} + wait O {b[5]/n/tiles:.0f} + work {(b[6]-b[4]-b[5])/n/tiles:.0f}")
