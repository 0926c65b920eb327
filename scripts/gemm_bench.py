"""TFLOP/s of qt_gemm_nt vs torch.matmul (cuBLAS) on the BASELINE shapes.  Every variant is captured in a CUDA graph
(20 launches per replay) and timed with CUDA events, so the numbers are device time, not Python/ctypes dispatch.
Usage: python scripts/gemm_bench.py [--json out.json]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
from quantized_training import _C

INNER = 20


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(INNER):
            fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * INNER)


def run(dev="cuda:0", verbose=True, quick=False):
    """Returns {shape: {qt_bf16_TF, cublas_bf16_TF, qt_fp8_TF, ...}}; quick=True drops the 8192^3 case."""
    global _VERBOSE
    _VERBOSE = verbose
    return _run(dev, quick)


_VERBOSE = True


def say(*a, **k):
    if _VERBOSE:
        print(*a, **k)


def _run(dev, quick):
    shapes = [("llama q/k/v/o 1024x4096x4096", 1, 1024, 4096, 4096), ("llama qkv fused 1024x12288x4096", 1, 1024, 12288, 4096),
              ("llama gate/up 1024x11008x4096", 1, 1024, 11008, 4096), ("llama gate+up fused 1024x22016x4096", 1, 1024, 22016, 4096),
              ("llama down 1024x4096x11008", 1, 1024, 4096, 11008), ("llama lm_head 1024x32000x4096", 1, 1024, 32000, 4096),
              ("bert qkv/out 6144x768x768", 1, 6144, 768, 768), ("bert ffn1 6144x3072x768", 1, 6144, 3072, 768),
              ("bert ffn2 6144x768x3072", 1, 6144, 768, 3072), ("square 8192^3", 1, 8192, 8192, 8192),
              ("bert qk^T 192x(384x384x64)", 192, 384, 384, 64), ("bert pv 192x(384x64x384)", 192, 384, 64, 384),
              ("llama qk^T 32x(1024x1024x128)", 32, 1024, 1024, 128), ("llama pv 32x(1024x128x1024)", 32, 1024, 128, 1024)]
    out = {}
    if quick:
        shapes = [s for s in shapes if "8192" not in s[0]]
    for name, b, M, N, K in shapes:
        a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
        c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * b * M * N * K
        ms_q = timed(lambda: _C.gemm_nt(a, w, out=c))
        ms_c = timed(lambda: torch.matmul(a, w.transpose(-1, -2), out=c))
        a8 = a.to(torch.float8_e4m3fn).view(torch.uint8); w8 = w.to(torch.float8_e4m3fn).view(torch.uint8)
        ms_8 = timed(lambda: _C.gemm_nt(a8, w8, operand_type=_C.GEMM_E4M3, out=c))
        out[name] = {"qt_bf16_TF": fl / ms_q / 1e9, "cublas_bf16_TF": fl / ms_c / 1e9, "qt_fp8_TF": fl / ms_8 / 1e9,
                     "qt_bf16_us": ms_q * 1e3, "cublas_bf16_us": ms_c * 1e3, "qt_fp8_us": ms_8 * 1e3,
                     "out_GBps_qt_bf16": 2.0 * b * M * N / ms_q / 1e6}
        mx = ""
        if b == 1 and K % 128 == 0:   # block-scaled fp8 (microscaling: one UE8M0 scale per 32 K elements)
            sa = _C.mx_pack_scales(torch.exp2(torch.randint(-3, 4, (M, K // 32), device=dev).float()))
            sw = _C.mx_pack_scales(torch.exp2(torch.randint(-3, 4, (N, K // 32), device=dev).float()))
            ms_x = timed(lambda: _C.gemm_nt(a8[0], w8[0], operand_type=_C.GEMM_E4M3, sf_a=sa, sf_b=sw, out=c[0]))
            out[name]["qt_mxfp8_TF"], out[name]["qt_mxfp8_us"] = fl / ms_x / 1e9, ms_x * 1e3
            mx = f" | qt mxfp8 {fl/ms_x/1e9:6.0f} TF {ms_x*1e3:7.1f} us"
        say(f"{name:38s} qt bf16 {fl/ms_q/1e9:6.0f} TF {ms_q*1e3:7.1f} us | cuBLAS bf16 {fl/ms_c/1e9:6.0f} TF {ms_c*1e3:7.1f} us | "
              f"qt fp8 {fl/ms_8/1e9:6.0f} TF {ms_8*1e3:7.1f} us" + mx, flush=True)
    # backward products (MN-major operands) vs the torch.matmul calls autograd would issue
    bwd = [("roberta 2048x768x768", 2048, 768, 768), ("roberta ffn1 2048x3072x768", 2048, 3072, 768),
           ("roberta ffn2 2048x768x3072", 2048, 768, 3072), ("llama o 1024x4096x4096", 1024, 4096, 4096),
           ("llama down 1024x4096x11008", 1024, 4096, 11008)]
    for name, M, N, K in bwd:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16); w = torch.randn(N, K, device=dev).to(torch.bfloat16)
        g = torch.randn(M, N, device=dev).to(torch.bfloat16)
        gx = torch.empty(M, K, device=dev, dtype=torch.bfloat16); gw = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        t = {"dgrad_qt": timed(lambda: _C.gemm_nt(g, w, b_mn=True, out=gx)), "dgrad_cublas": timed(lambda: torch.matmul(g, w, out=gx)),
             "wgrad_qt": timed(lambda: _C.gemm_nt(g, x, a_mn=True, b_mn=True, out=gw)),
             "wgrad_cublas": timed(lambda: torch.matmul(g.t(), x, out=gw))}
        g8 = g.to(torch.float8_e5m2).view(torch.uint8); w8 = w.to(torch.float8_e4m3fn).view(torch.uint8)
        x8 = x.to(torch.float8_e4m3fn).view(torch.uint8)
        t["dgrad_qt_fp8"] = timed(lambda: _C.gemm_nt(g8, w8, operand_type=_C.GEMM_E5M2_E4M3, b_mn=True, out=gx))
        t["wgrad_qt_fp8"] = timed(lambda: _C.gemm_nt(g8, x8, operand_type=_C.GEMM_E5M2_E4M3, a_mn=True, b_mn=True, out=gw))
        out["bwd " + name] = {k + "_TF": fl / v / 1e9 for k, v in t.items()}
        say(f"bwd {name:30s} " + " | ".join(f"{k} {fl/v/1e9:6.0f} TF {v*1e3:6.1f} us" for k, v in t.items()), flush=True)
    return out


if __name__ == "__main__":
    out = run()
    print(json.dumps(out))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
