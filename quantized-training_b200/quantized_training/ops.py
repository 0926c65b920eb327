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


_FP8_DTYPE = {"e4m3": torch.float8_e4m3fn, "e5m2": torch.float8_e5m2}
_FP8_OP = {("e4m3", "e4m3"): _C.GEMM_E4M3, ("e5m2", "e5m2"): _C.GEMM_E5M2,
           ("e4m3", "e5m2"): _C.GEMM_E4M3_E5M2, ("e5m2", "e4m3"): _C.GEMM_E5M2_E4M3}


class _LinearFp8Fn(torch.autograd.Function):
    """x holds values of an fp8 format exactly (it left a bare e4m3/e5m2 fake-quantizer), the weight is quantized
    straight to codes: both operands go to the FP8 tensor cores.  Products and fp32 accumulation are those of
    the bf16 path (every fp8 value is a bf16 value), at twice the MMA rate and 3/4 of the weight traffic."""

    @staticmethod
    def forward(ctx, x, weight, bias, wq, x_kind, codes):
        x2 = x.reshape(-1, x.shape[-1])
        xc = x2.contiguous().to(_FP8_DTYPE[x_kind]).view(torch.uint8)   # exact: x is representable
        wc = codes if codes is not None else wq.quantize_to_codes(weight)
        y = _C.gemm_nt(xc, wc, bias=bias.contiguous() if bias is not None else None,
                       operand_type=_FP8_OP[(x_kind, wq.fp8_kind)])
        if any(ctx.needs_input_grad[:3]):
            ctx.save_for_backward(x, wc)
            ctx.w_kind = wq.fp8_kind
            ctx.w_scale = wq.scale
        ctx.has_bias = bias is not None
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        x, wc = ctx.saved_tensors
        w = (wc.view(_FP8_DTYPE[ctx.w_kind]).to(g.dtype) * ctx.w_scale.to(g.dtype))   # the fake-quantized weight
        g2 = g.reshape(-1, g.shape[-1])
        gx = (g2 @ w).view_as(x) if ctx.needs_input_grad[0] else None
        gw = g2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None   # STE through the quantizer
        gb = g2.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None


def linear_fp8(x, weight, bias, weight_fq, x_kind, codes=None):
    """F.linear(x, weight_fq(weight), bias) with both operands as fp8 codes.  Caller guarantees that x already holds
    `x_kind` values exactly and that weight_fq is a bare (scale 1) e4m3/e5m2 quantizer."""
    return _LinearFp8Fn.apply(x, weight, bias, weight_fq, x_kind, codes)


def fp8_route(module, x, weight_fq):
    """Which fp8 format the input of `module` is already quantized to by its forward pre-hook, if the whole
    product can run on the FP8 tensor cores; else None."""
    if not (_ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[-1] % 16 == 0):
        return None
    hooks = getattr(module, "activation_pre_process", None)
    act = hooks["0"] if hooks is not None and "0" in hooks else None
    for fq in (act, weight_fq):
        if fq is None or getattr(fq, "fp8_kind", None) is None or fq.qscheme is not None or fq.is_per_channel:
            return None
        if fq._flags() != (False, True):
            return None
    if module.weight.shape[0] % 8:
        return None
    return act.fp8_kind


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
