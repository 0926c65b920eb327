"""QT_GEMM_CODE8_B (weights as one-byte codes, decoded in shared memory) vs the bf16-operand kernel vs cuBLAS, over the
token count M: where does 1 byte / weight win?  Weights are cycled through a pool larger than L2 so that every launch
streams them from HBM (the regime the code form is for).  python scripts/code8_bench.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import _C

dev = "cuda:0"
fqm = qt.FusedAmaxObsFakeQuantize("posit8_1", device=dev)
lut = _C.code_table_host(fqm._fmt).to(dev)


def timed(fns, reps=3):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * len(fns))


out = {}
for N, K in ((4096, 4096), (11008, 4096), (4096, 11008)):
    pool = max(2, (300 << 20) // (N * K * 2) + 1)          # > 126 MB L2 of bf16 weights
    ws = [fqm((torch.randn(N, K, device=dev) * 0.5).bfloat16()) for _ in range(pool)]
    wcs = [_C.quantize_codes8(w, torch.empty(N, K, dtype=torch.uint8, device=dev), fqm._fmt) for w in ws]
    for M in (8, 32, 128, 256, 512, 1024):
        x = fqm(torch.randn(M, K, device=dev).bfloat16())
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t_bf = timed([lambda w=w: _C.gemm_nt(x, w, out=c) for w in ws])
        t_cd = timed([lambda w=w: _C.gemm_nt(x, w, out=c, operand_type=_C.GEMM_CODE8_B, code_lut=lut) for w in wcs])
        t_cb = timed([lambda w=w: torch.matmul(x, w.t(), out=c) for w in ws])
        fl = 2.0 * M * N * K
        out[f"{M}x{N}x{K}"] = {"bf16_us": t_bf * 1e3, "code8_us": t_cd * 1e3, "cublas_us": t_cb * 1e3,
                               "bf16_TF": fl / t_bf / 1e9, "code8_TF": fl / t_cd / 1e9,
                               "weight_GBps_bf16": 2.0 * N * K / t_bf / 1e6, "weight_GBps_code8": 1.0 * N * K / t_cd / 1e6}
        print(f"M={M:5d} N={N:6d} K={K:6d}  bf16 operands {t_bf*1e3:7.1f} us ({fl/t_bf/1e9:6.0f} TF, weights {2.0*N*K/t_bf/1e6:5.0f} GB/s) | "
              f"code8 {t_cd*1e3:7.1f} us ({fl/t_cd/1e9:6.0f} TF, weights {1.0*N*K/t_cd/1e6:5.0f} GB/s) | cuBLAS {t_cb*1e3:7.1f} us", flush=True)
    del ws, wcs
    torch.cuda.empty_cache()
print(json.dumps(out))
