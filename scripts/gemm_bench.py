"""TFLOP/s of qt_gemm_nt vs torch.matmul (cuBLAS) on the BASELINE shapes. Usage: python scripts/gemm_bench.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
from quantized_training import _C

def timed(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

dev = "cuda:0"
shapes = [("llama q/k/v/o 1024x4096x4096", 1, 1024, 4096, 4096), ("llama gate/up 1024x11008x4096", 1, 1024, 11008, 4096),
          ("llama down 1024x4096x11008", 1, 1024, 4096, 11008), ("llama lm_head 1024x32000x4096", 1, 1024, 32000, 4096),
          ("bert qkv/out 6144x768x768", 1, 6144, 768, 768), ("bert ffn1 6144x3072x768", 1, 6144, 3072, 768),
          ("bert ffn2 6144x768x3072", 1, 6144, 768, 3072), ("square 8192^3", 1, 8192, 8192, 8192),
          ("bert qk^T 192x(384x384x64)", 192, 384, 384, 64), ("llama qk^T 32x(1024x1024x128)", 32, 1024, 1024, 128)]
out = {}
for name, b, M, N, K in shapes:
    a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
    c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * b * M * N * K
    ms_q = timed(lambda: _C.gemm_nt(a, w, out=c))
    ms_c = timed(lambda: torch.matmul(a, w.transpose(-1, -2), out=c))
    a8 = a.to(torch.float8_e4m3fn).view(torch.uint8); w8 = w.to(torch.float8_e4m3fn).view(torch.uint8)
    ms_8 = timed(lambda: _C.gemm_nt(a8, w8, operand_type=_C.GEMM_E4M3, out=c))
    out[name] = {"qt_bf16_TF": fl / ms_q / 1e9, "cublas_bf16_TF": fl / ms_c / 1e9, "qt_fp8_TF": fl / ms_8 / 1e9}
    print(f"{name:36s} qt bf16 {fl/ms_q/1e9:7.0f} TF | cuBLAS bf16 {fl/ms_c/1e9:7.0f} TF | qt fp8 {fl/ms_8/1e9:7.0f} TF", flush=True)
print(json.dumps(out))
