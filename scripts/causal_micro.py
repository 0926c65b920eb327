"""Three-kernel attention chain with and without the causal schedule, per kernel (graph-timed), Llama-2-7B window."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import _C
dev = "cuda:0"
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
B, H, S, D = 1, 32, 1024, 128
for spec, codes in (("posit8_1", False), ("e4m3", True)):
    m = qt.FusedAmaxObsFakeQuantize(spec, device=dev); fmt, lut = m._fmt, m.lut
    qkv = m((torch.randn(B, S, 3 * H * D, device=dev) * 1.5).bfloat16())
    q = qkv[..., :H * D].view(B, S, H, D).transpose(1, 2); k = qkv[..., H * D:2 * H * D].view(B, S, H, D).transpose(1, 2)
    vt = torch.empty(B, H, D, S, device=dev, dtype=torch.bfloat16); _C.fq_transpose(qkv[..., 2 * H * D:].view(B, S, H, D), vt, 0, fmt, lut=lut)
    mask = torch.full((S, S), torch.finfo(torch.bfloat16).min, device=dev, dtype=torch.bfloat16).triu(1)[None].contiguous()
    op = 0
    if codes:
        enc = lambda t: t.contiguous().to(torch.float8_e4m3fn).view(torch.uint8)
        qc = enc(qkv[..., :2 * H * D]); q = qc[..., :H * D].view(B, S, H, D).transpose(1, 2); k = qc[..., H * D:].view(B, S, H, D).transpose(1, 2)
        vt = enc(vt); op = _C.GEMM_E4M3
    scores = torch.empty(B, H, S, S, device=dev, dtype=torch.bfloat16)
    probs = torch.empty(B, H, S, S, device=dev, dtype=torch.uint8 if codes else torch.bfloat16)
    ctx = torch.empty(B, S, H * D, device=dev, dtype=torch.bfloat16); o4 = ctx.view(B, S, H, D).transpose(1, 2)
    flag = _C.causal_mask_check(mask)
    for name, c, f in (("full", False, None), ("causal", True, None), ("causal+flag", True, flag)):
        t1 = timed(lambda: _C.gemm_nt(q, k, out=scores, operand_type=op, causal=_C.CAUSAL_OUT_LOWER if c else 0, causal_flag=f))
        t2 = timed(lambda: _C.softmax_fq(scores, probs, D ** -0.5, mask, H * S, S, 1, _C.FQ_POST | (_C.SOFTMAX_CAUSAL if c else 0), fmt, lut=lut, causal_flag=f))
        t3 = timed(lambda: _C.gemm_nt(probs, vt, out=o4, operand_type=op, causal=_C.CAUSAL_A_LOWER if c else 0, causal_flag=f))
        print(f"{spec:9s} {name:12s} QK^T {t1:6.1f}  softmax {t2:6.1f}  PV {t3:6.1f}  sum {t1 + t2 + t3:6.1f} us", flush=True)
    print(f"{spec:9s} mask check {timed(lambda: _C.causal_mask_check(mask, flag)):6.1f} us")
# does the model path engage it?
from quantized_training import fused
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import llama_bench
model, fwd, ids = llama_bench.setup("e4m3", torch.device(dev), layers=2)
calls = []
orig = _C.gemm_nt
def spy(*a, **k):
    calls.append((k.get("causal", 0), k.get("causal_flag") is not None)); return orig(*a, **k)
_C.gemm_nt = spy
fwd(); fwd(); calls.clear(); fwd()
print("gemm_nt calls (causal, flag):", calls)
