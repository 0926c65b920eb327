"""Fused attention core vs the three-kernel chain at the Llama-2-7B window shape (graph-timed)."""
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
for spec, codes in (("posit8_1", False), ("e4m3", True)):
    for (B, H, S, D) in ((1, 32, 1024, 128), (16, 12, 384, 64)):
        if codes and D != 128: continue
        m = qt.FusedAmaxObsFakeQuantize(spec, device=dev); fmt, lut = m._fmt, m.lut
        qkv = m((torch.randn(B, S, 3 * H * D, device=dev) * 1.5).bfloat16())
        q = qkv[..., :H * D].view(B, S, H, D).transpose(1, 2); k = qkv[..., H * D:2 * H * D].view(B, S, H, D).transpose(1, 2)
        vt = torch.empty(B, H, D, S, device=dev, dtype=torch.bfloat16); _C.fq_transpose(qkv[..., 2 * H * D:].view(B, S, H, D), vt, 0, fmt, lut=lut)
        mask = torch.full((S, S), torch.finfo(torch.bfloat16).min, device=dev, dtype=torch.bfloat16).triu(1)[None].contiguous()
        kw = {}
        if codes:
            enc = lambda t: t.contiguous().to(torch.float8_e4m3fn).view(torch.uint8)
            qc = enc(qkv[..., :2 * H * D]); q = qc[..., :H * D].view(B, S, H, D).transpose(1, 2); k = qc[..., H * D:].view(B, S, H, D).transpose(1, 2)
            vt = enc(vt); kw = dict(qk_type=_C.GEMM_E4M3, pv_type=_C.GEMM_E4M3)
        ctx = torch.empty(B, S, H * D, device=dev, dtype=torch.uint8 if codes else torch.bfloat16)
        o4 = ctx.view(B, S, H, D).transpose(1, 2)
        for name, mk, causal in (("causal-skip", mask, True), ("masked, no skip", mask, False), ("no mask", None, False)):
            us = timed(lambda: _C.attention_fq(q, k, vt, o4, D ** -0.5, mk, causal, _C.FQ_POST | _C.FQ_OUT, fmt, lut, **kw))
            print(f"{spec:9s} B{B} H{H} S{S} D{D}  fused attention {name:16s} {us:8.1f} us", flush=True)
        scores = torch.empty(B, H, S, S, device=dev, dtype=torch.bfloat16); probs = torch.empty(B, H, S, S, device=dev, dtype=ctx.dtype)
        def chain():
            _C.gemm_nt(q, k, out=scores, operand_type=kw.get("qk_type", 0))
            _C.softmax_fq(scores, probs, D ** -0.5, mask, H * S, S, 1, _C.FQ_POST, fmt, lut=lut)
            _C.gemm_nt(probs, vt, out=o4 if not codes else torch.empty(B, S, H * D, device=dev, dtype=torch.bfloat16).view(B, S, H, D).transpose(1, 2), operand_type=kw.get("pv_type", 0))
        print(f"{spec:9s} B{B} H{H} S{S} D{D}  three-kernel chain (no final fq)   {timed(chain):8.1f} us", flush=True)
