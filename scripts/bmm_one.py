import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
from quantized_training import _C
dev = "cuda:0"
b, M, N, K = 32, 1024, 1024, 128
a = torch.randn(b, M, K, device=dev).to(torch.bfloat16); w = torch.randn(b, N, K, device=dev).to(torch.bfloat16)
c = torch.empty(b, M, N, device=dev, dtype=torch.bfloat16)
for _ in range(6):
    _C.gemm_nt(a, w, out=c)
torch.cuda.synchronize()
