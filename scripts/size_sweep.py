"""BASELINE configs[2]: standalone fake quant over 2^20 .. 2^32 bf16 elements (one B200), GB/s of algorithmic
read + write bytes per format; inputs up to 2^24 elements (64 MB in + out) are L2-resident between launches -- the
column says so.  python scripts/size_sweep.py > profiles/size_sweep_r01.txt"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
dev = "cuda:0"
specs = ["int4", "int8", "e4m3", "e5m2", "fp6_e3m2", "fp4_e2m1", "posit8_1", "posit8_2"]
mods = {s: qt.FusedAmaxObsFakeQuantize(s, device=dev) for s in specs}
unit = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
print("log2_n  " + "".join(f"{s:>10s}" for s in specs) + "   note")
for L in range(20, 33, 2):
    n = 1 << L
    x = (torch.randn(min(n, 1 << 26), device=dev) * 3).to(torch.bfloat16)
    x = x.repeat(n // x.numel()) if n > x.numel() else x
    y = torch.empty_like(x)
    row = []
    for s in specs:
        m = mods[s]
        f = lambda: qt._C.fq_forward(x, y, 1, 1, n, m._fmt, unit, None, m.lut)
        f(); torch.cuda.synchronize()
        reps = 5
        t = 0.0
        for _ in range(reps):
            flush.zero_()                       # evict x / y from L2 between timed launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            t += a.elapsed_time(b)
        row.append(4.0 * n / (t / reps * 1e-3) / 1e9)
    note = "L2 flushed between launches; launch-latency bound" if L <= 24 else "L2 flushed between launches"
    print(f"{L:6d}  " + "".join(f"{v:10.0f}" for v in row) + "   " + note, flush=True)
    del x, y
