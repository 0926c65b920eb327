"""One launch family for profiling: per-channel fake quant along the last axis of [N/4096, 4096] bf16."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
dev = "cuda:0"; n = 1 << 28; rows = n // 4096
x = (torch.randn(n, device=dev) * 4).bfloat16(); y = torch.empty_like(x)
m = qt.FusedAmaxObsFakeQuantize(sys.argv[1] if len(sys.argv) > 1 else "posit8_1", device=dev)
scc = torch.rand(4096, device=dev) * 0.05 + 0.01; am = torch.zeros(4096, device=dev)
def t(fn, reps=10):
    fn(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
for name, fn in [("ax=-1 scale+amax", lambda: qt._C.fq_forward(x, y, rows, 4096, 1, m._fmt, scc, am, m.lut)),
                 ("ax=-1 scale", lambda: qt._C.fq_forward(x, y, rows, 4096, 1, m._fmt, scc, None, m.lut))]:
    ms = t(fn, int(os.environ.get("REPS", "10"))); print(f"{name:18s} {4.0 * n / ms / 1e6:8.0f} GB/s")
