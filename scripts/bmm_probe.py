"""Timing experiments on the short-K batched product (QT_GEMM_DEBUG bit mask: 1 no store, 2 no TMEM load,
4/8/16 force tile width 128/64/256)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from quantized_training import _C
dev = "cuda:0"
def timed(fn, inner=20, reps=5):
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
for name, b, M, N, K in [("llama qk^T", 32, 1024, 1024, 128), ("llama pv", 32, 1024, 128, 1024), ("bert qk^T", 192, 384, 384, 64),
                         ("llama qkvo", 1, 1024, 4096, 4096)]:
    a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
    c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
    for flags in (0, 1, 2, 3, 4, 5, 8, 16, 17, 19):
        os.environ["QT_GEMM_DEBUG"] = str(flags)
        us = timed(lambda: _C.gemm_nt(a, w, out=c))
        print(f"{name:12s} debug={flags:2d}  {us:8.1f} us", flush=True)
os.environ["QT_GEMM_DEBUG"] = "0"
