"""SM clock / power while one kernel shape runs back to back for ~2 s (nvidia-smi sampled every 50 ms)."""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from detsam2_b200 import ops
BF16 = torch.bfloat16
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "flash"
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if which == "flash":
    q = torch.randn(16, 4096, 256, device=dev).to(BF16); k = torch.randn(16, 28736, 256, device=dev).to(BF16)
    v = torch.randn(16, 28736, 64, device=dev).to(BF16); o = torch.zeros(16, 4096, 64, device=dev, dtype=BF16)
    fn = lambda: ops.flash_attn(q, k, v, o, 1 / 16.0, impl=impl)
    flops = 2.0 * 16 * 4096 * 28736 * 320
else:
    a = torch.randn(8192, 8192, device=dev).to(BF16); w = torch.randn(8192, 8192, device=dev).to(BF16)
    o = torch.zeros(8192, 8192, device=dev, dtype=BF16)
    if which == "cublas":
        fn = lambda: torch.matmul(a, w.t(), out=o)
    else:
        fn = lambda: ops.gemm(a, w, out_bf16=o)
    flops = 2.0 * 8192 ** 3
fn(); torch.cuda.synchronize()
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
t = threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True); t.start()
time.sleep(0.3)
# single isolated launch (burst clocks)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
single = e0.elapsed_time(e1)
time.sleep(0.3)
n0 = len(rows)
times = []
t0 = time.time()
while time.time() - t0 < 2.0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / 20)
n1 = len(rows)
p.terminate()
print(f"{which} impl={impl}: single launch {single*1e3:.1f} us ({flops/single/1e9:.0f} TFLOP/s); sustained first {times[0]*1e3:.1f} us, last {times[-1]*1e3:.1f} us ({flops/times[-1]/1e9:.0f} TFLOP/s)")
print("idle:", rows[:3]); print("load:", rows[n0:n1][::4])
