import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
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
for b, M, N in [(1, 32768, 1024), (32, 1024, 1024), (1, 128 * 148, 256), (1, 128 * 148 * 4, 256)]:
    for K in (64, 128, 256, 512, 1024, 2048):
        a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
        c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
        r = []
        for flags in (16, 19):
            os.environ["QT_GEMM_DEBUG"] = str(flags)
            r.append(timed(lambda: _C.gemm_nt(a, w, out=c)))
        tiles = b * (M // 128) * (N // 256)
        print(f"b={b} M={M} N={N} K={K:5d} tiles/CTA={tiles/148:5.2f}  full {r[0]:7.1f} us  no-epilogue {r[1]:7.1f} us", flush=True)
