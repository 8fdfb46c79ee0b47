"""Where the flash kernel waits: barrier-stall cycles of the MMA issuer and of a softmax warp (impl 8)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from detsam2_b200 import ops, capi
BF16 = torch.bfloat16
lib = capi.load()
for name, Lq, Lk, DV in (("cross", 4096, 28736, 64), ("self", 4096, 4096, 256)):
    q = torch.randn(16, Lq, 256, device="cuda").to(BF16); k = torch.randn(16, Lk, 256, device="cuda").to(BF16)
    v = torch.randn(16, Lk, DV, device="cuda").to(BF16); o = torch.zeros(16, Lq, DV, device="cuda", dtype=BF16)
    buf = (C.c_ulonglong * 16)()
    ops.flash_attn(q, k, v, o, 1 / 16.0, impl=8); torch.cuda.synchronize()
    lib.ds2_debug_flash_stalls(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.flash_attn(q, k, v, o, 1 / 16.0, impl=8); e1.record(); torch.cuda.synchronize()
    lib.ds2_debug_flash_stalls(buf, 1)
    b = list(buf); n = max(b[7], 1); tiles = (Lk + (128 if DV == 64 else 64) - 1) // (128 if DV == 64 else 64)
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.0f} us, {n} CTAs, {tiles} key tiles; per CTA per tile (clk): MMA warp total {b[3]/n/tiles:.0f} = "
          f"wait K {b[0]/n/tiles:.0f} + wait V {b[1]/n/tiles:.0f} + wait P {b[2]/n/tiles:.0f} + issue {(b[3]-b[0]-b[1]-b[2])/n/tiles:.0f}; "
          f"softmax warp total {b[6]/n/tiles:.0f} = wait S {b[4]/n/tiles:.0f} + wait O {b[5]/n/tiles:.0f} + work {(b[6]-b[4]-b[5])/n/tiles:.0f}")
