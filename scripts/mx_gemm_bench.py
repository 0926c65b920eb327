"""Block-scaled (microscaling) GEMM on tcgen05.mma kind::mxf8f6f4.block_scale vs the plain fp8 product and vs the
dequantize-then-bf16 route linear_mx used before.  CUDA-graph timed like scripts/gemm_bench.py.
Usage: python scripts/mx_gemm_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import quantized_training as qt
from quantized_training import _C, decomposed
from quantized_training.quantizer import get_quant_min_max
from gemm_bench import timed
dev = "cuda:0"
torch.manual_seed(0)
qmap = qt.get_quantization_map("fp8_e4m3", dev)
qmax = float(get_quant_min_max("fp8_e4m3")[1])
for name, M, N, K in [("llama o 1024x4096x4096", 1024, 4096, 4096), ("llama qkv 1024x12288x4096", 1024, 12288, 4096),
                      ("llama down 1024x4096x11008", 1024, 4096, 11008), ("bert ffn1 6144x3072x768", 6144, 3072, 768),
                      ("square 8192^3", 8192, 8192, 8192)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    xs, xq = torch.ops.quantized_ops.quantize_mx(x, qmap, [-1], 32, qmax, True, None)
    ws, wq = torch.ops.quantized_ops.quantize_mx(w, qmap, [-1], 32, qmax, True, None)
    a8, b8 = xq.to(torch.float8_e4m3fn).view(torch.uint8), wq.to(torch.float8_e4m3fn).view(torch.uint8)
    pa, pb = _C.mx_pack_scales(xs.float().contiguous()), _C.mx_pack_scales(ws.float().contiguous())
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    t_mx = timed(lambda: _C.gemm_nt(a8, b8, operand_type=_C.GEMM_E4M3, sf_a=pa, sf_b=pb, out=out)) * 1e3
    t_f8 = timed(lambda: _C.gemm_nt(a8, b8, operand_type=_C.GEMM_E4M3, out=out)) * 1e3
    decomposed.MX_TENSOR_CORES = "0"
    t_old = timed(lambda: torch.ops.quantized_ops.linear_mx(xq, wq, None, input_scale=xs, weight_scale=ws, block_size=32)) * 1e3
    decomposed.MX_TENSOR_CORES = "assume"
    t_new = timed(lambda: torch.ops.quantized_ops.linear_mx(xq, wq, None, input_scale=xs, weight_scale=ws, block_size=32)) * 1e3
    decomposed.MX_TENSOR_CORES = "1"
    print(f"{name:28s} block-scaled kernel {t_mx:7.1f} us {fl/t_mx/1e6:5.0f} TF | plain fp8 kernel {t_f8:7.1f} us {fl/t_f8/1e6:5.0f} TF | "
          f"linear_mx op: dequantize + bf16 GEMM {t_old:7.1f} us, block-scaled route (codes + scale packing each call) {t_new:7.1f} us", flush=True)
