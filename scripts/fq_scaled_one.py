import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
dev = "cuda:0"; n = 1 << 28
x = (torch.randn(n, device=dev) * 4).bfloat16(); y = torch.empty_like(x)
m = qt.FusedAmaxObsFakeQuantize(sys.argv[1] if len(sys.argv) > 1 else "e4m3", device=dev)
sc = torch.full((1,), 0.0123, device=dev); hist = torch.zeros(1, device=dev); unit = torch.ones(1, device=dev)
def t(fn, reps=10):
    fn(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
for name, fn in [("bare", lambda: qt._C.fq_forward(x, y, 1, 1, n, m._fmt, unit, None, m.lut)),
                 ("bare+amax", lambda: qt._C.fq_forward(x, y, 1, 1, n, m._fmt, unit, hist, m.lut)),
                 ("scaled", lambda: qt._C.fq_forward(x, y, 1, 1, n, m._fmt, sc, None, m.lut)),
                 ("scaled+amax", lambda: qt._C.fq_forward(x, y, 1, 1, n, m._fmt, sc, hist, m.lut))]:
    ms = t(fn); print(f"{name:12s} {4.0 * n / ms / 1e6:8.0f} GB/s")
