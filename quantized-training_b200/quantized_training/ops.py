"""GEMM-shaped ops of the quantized modules, routed to the tcgen05 kernel (qt_gemm_nt).

`linear` replaces F.linear in the QAT Linear / LoRA Linear (reference modules/qat/linear.py:40-41, lora.py:52)
and `matmul` replaces torch.matmul in MatmulFunctional (modules/quantizable/functional_modules.py:22-27).
The kernel path is taken for bf16 CUDA operands whose layout the kernel accepts (16-byte aligned rows,
N % 8 == 0); everything else (fp32 models, odd shapes) uses the stock torch op, i.e. the reference's own
K5/K6 cuBLAS path.  Backward GEMMs (dgrad / wgrad) use torch.matmul on the saved quantized operands --
what autograd does in the reference; the forward is the hot path (north star: forward evaluation).
"""
import torch
import torch.nn.functional as F

from . import _C

__all__ = ["linear", "matmul", "kernel_eligible"]

_ENABLED = True


def set_enabled(flag: bool):
    """Route linear/matmul through the tcgen05 kernel (default) or through torch (for A/B timing)."""
    global _ENABLED
    _ENABLED = bool(flag)


def _rows_ok(t):
    return t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and all(s % 8 == 0 for s in t.stride()[:-1])


def kernel_eligible(a, b_nk, n, bias=None):
    return (_ENABLED and a.is_cuda and a.dtype == torch.bfloat16 and b_nk.dtype == torch.bfloat16 and n % 8 == 0
            and a.shape[-1] % 8 == 0 and a.numel() > 0 and b_nk.numel() > 0
            and (bias is None or bias.dtype == torch.bfloat16))


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1])
        if not _rows_ok(x2):
            x2 = x2.contiguous()
        wk = w if _rows_ok(w) else w.contiguous()
        y = _C.gemm_nt(x2, wk, bias=b.contiguous() if b is not None else None)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        g2 = g.reshape(-1, g.shape[-1])
        if ctx.needs_input_grad[0]:
            gx = (g2 @ w).view_as(x)
        if ctx.needs_input_grad[1]:
            gw = g2.t() @ x.reshape(-1, x.shape[-1])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


def linear(x, weight, bias=None):
    if kernel_eligible(x, weight, weight.shape[0], bias):
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class _MatmulFn(torch.autograd.Function):
    """x [..., M, K] @ y [..., K, N]; the kernel wants y as [..., N, K] with a unit-stride K axis, which is free
    when y is itself a transposed view (k^T in attention) and one transpose copy of the small operand otherwise."""

    @staticmethod
    def forward(ctx, x, y):
        yt = y.transpose(-1, -2)
        out = _C.gemm_nt(x, yt)
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        gx = g @ y.transpose(-1, -2) if ctx.needs_input_grad[0] else None
        gy = x.transpose(-1, -2) @ g if ctx.needs_input_grad[1] else None
        return gx, gy


def matmul(x, y):
    if (x.dim() >= 2 and y.dim() >= 2 and x.dim() == y.dim() and x.shape[:-2] == y.shape[:-2]
            and kernel_eligible(x, y, y.shape[-1]) and y.shape[-2] % 8 == 0):
        return _MatmulFn.apply(x, y)
    return torch.matmul(x, y)
