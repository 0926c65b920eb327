import os, sys, ctypes, torch
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
from quantized_training import _C
dev = "cuda:0"
L = _C.lib()
for (b, M, N, K, flags) in [(32, 1024, 1024, 128, 16 + 32), (32, 1024, 1024, 128, 19 + 32)]:
    a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
    c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
    os.environ["QT_GEMM_DEBUG"] = str(flags)
    for _ in range(3):
        _C.gemm_nt(a, w, out=c)
    torch.cuda.synchronize()
    buf = np.zeros((3, 256), dtype=np.int64)
    L.qt_gemm_debug_trace(ctypes.c_void_p(buf.ctypes.data))
    t0 = min(buf[r][0] for r in range(3) if buf[r][0] > 0)
    print(f"--- b={b} M={M} N={N} K={K}")
    for r, name in enumerate(["producer(after empty wait)", "mma(after tmem_empty / full waits)", "epilogue w2 (after tmem_full; after each chunk)"]):
        v = buf[r]; v = v[v > 0][:40] - t0
        print(name, list(v))
