"""Micro-benchmarks of the fused kernels at Llama-2-7B window shapes (graph-timed). Usage: python scripts/fused_micro.py [e4m3|posit8_1]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import _C
dev = "cuda:0"
spec = sys.argv[1] if len(sys.argv) > 1 else "e4m3"
m = qt.FusedAmaxObsFakeQuantize(spec, device=dev); fmt, lut = m._fmt, m.lut
codes = m.fp8_kind is not None
odt = torch.uint8 if codes else torch.bfloat16
def timed(fn, inner=10, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner): fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * inner) * 1e3
S, H, D, HID, I = 1024, 32, 128, 4096, 11008
scores = (torch.randn(1, H, S, S, device=dev) * 3).bfloat16()
mask = torch.full((S, S), torch.finfo(torch.bfloat16).min, device=dev, dtype=torch.bfloat16).triu(1)[None]
probs = torch.empty(scores.shape, dtype=odt, device=dev)
us = timed(lambda: _C.softmax_fq(scores, probs, D ** -0.5, mask, H * S, S, 1, _C.FQ_POST, fmt, lut=lut))
print(f"softmax causal  {us:7.1f} us   {(scores.numel() * (2 + probs.element_size())) / us / 1e6:6.2f} TB/s")
us = timed(lambda: _C.softmax_fq(scores, probs, D ** -0.5, None, H * S, S, 1, _C.FQ_POST, fmt, lut=lut))
print(f"softmax no mask {us:7.1f} us")
x = torch.randn(S, HID, device=dev).bfloat16(); w = torch.ones(HID, device=dev).bfloat16(); y = torch.empty(x.shape, dtype=odt, device=dev)
us = timed(lambda: _C.norm_fq(x, y, 0, w, None, 1e-5, _C.FQ_POST, fmt, lut=lut)); print(f"rmsnorm         {us:7.1f} us")
gu = torch.randn(S, 2 * I, device=dev).bfloat16(); o = torch.empty(S, I, dtype=odt, device=dev)
us = timed(lambda: _C.act_mul_fq(gu[:, :I], gu[:, I:], o, "silu", _C.FQ_POST, fmt, lut=lut)); print(f"silu*up         {us:7.1f} us")
qkv = torch.randn(S, 3 * HID, device=dev).bfloat16(); qk = torch.empty(2, S, H, D, dtype=odt, device=dev)
cos = torch.rand(S, D, device=dev).bfloat16(); sin = torch.rand(S, D, device=dev).bfloat16()
us = timed(lambda: _C.rope_fq(qkv[:, :HID].view(S, H, D), qk[0], qkv[:, HID:2 * HID].view(S, H, D), qk[1], cos, sin, _C.FQ_POST, fmt, lut=lut)); print(f"rope q,k        {us:7.1f} us")
vt = torch.empty(1, H, D, S, dtype=odt, device=dev)
us = timed(lambda: _C.fq_transpose(qkv[:, 2 * HID:].view(1, S, H, D), vt, _C.FQ_POST, fmt, lut=lut)); print(f"fq transpose v  {us:7.1f} us")
ctx = torch.randn(S, HID, device=dev).bfloat16()
us = timed(lambda: m(ctx)); print(f"fq ctx (module) {us:7.1f} us")
