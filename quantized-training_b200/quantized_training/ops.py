"""GEMM-shaped ops of the quantized modules -- forward AND backward -- on the tcgen05 kernel (qt_gemm_nt_ex).

`linear` replaces F.linear in the QAT Linear / LoRA Linear (reference modules/qat/linear.py:40-41, lora.py:52)
and `matmul` replaces torch.matmul in MatmulFunctional (modules/quantizable/functional_modules.py:22-27).
Their autograd (what torch derives for the reference) is the same kernel with MN-major operand descriptors:

    y  = x W^T          A = x   [M, K]           B = W  [N, K]                 both K-major
    gx = g W            A = g   [M, N]           B = W  read MN-major (contraction over its row axis)
    gW = g^T x          A = g   read MN-major    B = x  read MN-major

so no operand is ever transposed or copied.  There is no automatic fallback to torch: operands the kernel cannot
address directly (row strides that are not multiples of 16 bytes, N or K not a multiple of 8 -- e.g. a 3-class
classifier head) are zero-padded into aligned buffers and still run on the kernel; anything else (fp32 / fp16
models, CPU tensors) raises.  `set_enabled(False)` is an explicit A/B switch for tests and benches that want the
reference's own cuBLAS op sequence beside the kernel; nothing in the library turns it off.
"""
import torch
import torch.nn.functional as F

from . import _C

__all__ = ["linear", "matmul", "gemm", "set_enabled"]

_ENABLED = True


def set_enabled(flag: bool):
    """A/B switch for tests and benches: True (default) = tcgen05 kernel, False = the stock torch ops (the reference's
    K5/K6 cuBLAS path).  Never flipped by the library itself."""
    global _ENABLED
    _ENABLED = bool(flag)


def _require_bf16_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("quantized_training GEMMs run on CUDA tensors only: there is no CPU fallback")
        if t.dtype != torch.bfloat16:
            raise TypeError(f"quantized_training GEMMs take bfloat16 operands (run the model with --bf16 as every "
                            f"BASELINE configuration does); got {t.dtype}")


def _pad_to(t, dim, mult):
    """zero-pad dimension `dim` of t up to a multiple of `mult` (one small copy; zeros add nothing to a product)."""
    dim %= t.dim()
    size = t.shape[dim]
    extra = (-size) % mult
    if extra == 0:
        return t
    pad = [0, 0] * t.dim()
    pad[2 * (t.dim() - 1 - dim) + 1] = extra
    return F.pad(t, pad)


def _operand(t, align=8):
    """Logical [..., rows, k] operand -> (tensor to hand to the kernel, mn flag).  A unit-stride k axis is K-major
    (passed as is); a unit-stride rows axis is MN-major (passed as the stored [..., k, rows] view); anything else --
    including row / batch strides that are not multiples of 16 bytes -- is copied once into a contiguous buffer."""
    def strides_ok(v):
        return v.data_ptr() % 16 == 0 and all(v.shape[i] == 1 or (v.stride(i) % align == 0 and v.stride(i) > 0)
                                              for i in range(v.dim() - 1))
    if t.stride(-1) == 1 and strides_ok(t):
        return t, False
    tt = t.transpose(-1, -2)
    if tt.stride(-1) == 1 and strides_ok(tt):
        return tt, True
    return t.contiguous(), False


def gemm(a, b, **kw):
    """out[..., m, n] = sum_k a[..., m, k] * b[..., n, k] for LOGICAL operands given as any strided views; each is
    read K-major or MN-major as it lies.  K and N that are not multiples of 8 (a 3-class classifier head) are
    zero-padded into aligned copies -- zeros add nothing to a product -- and the result is sliced."""
    _require_bf16_cuda(a, b, kw.get("bias"), kw.get("residual"))
    n, k = b.shape[-2], b.shape[-1]
    if a.shape[-1] != k:
        raise ValueError(f"inner dimensions differ: {tuple(a.shape)} x {tuple(b.shape)}^T")
    if a.numel() == 0 or b.numel() == 0:
        return a.new_zeros(*a.shape[:-1], n)
    if k % 8:
        a, b = _pad_to(a, -1, 8), _pad_to(b, -1, 8)
    if n % 8:
        if kw.get("residual") is not None or kw.get("out") is not None or kw.get("glu"):
            raise ValueError("N % 8 != 0 with a residual / out= / glu is not supported")
        bias = kw.pop("bias", None)
        out = gemm(a, _pad_to(b, -2, 8), bias=None if bias is None else _pad_to(bias, 0, 8), **kw)
        return out[..., :n]
    a_op, a_mn = _operand(a)
    b_op, b_mn = _operand(b)
    bias = kw.get("bias")
    if bias is not None and not bias.is_contiguous():
        kw["bias"] = bias.contiguous()
    return _C.gemm_nt(a_op, b_op, a_mn=a_mn, b_mn=b_mn, **kw)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1])
        y = gemm(x2, w, bias=b)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        g2 = g.reshape(-1, g.shape[-1])
        if ctx.needs_input_grad[0]:
            gx = gemm(g2, w.t()).view_as(x)                            # dgrad: W read MN-major
        if ctx.needs_input_grad[1]:
            gw = gemm(g2.t(), x.reshape(-1, x.shape[-1]).t())          # wgrad: g and x read MN-major
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


def linear(x, weight, bias=None):
    if not _ENABLED:
        return F.linear(x, weight, bias)
    _require_bf16_cuda(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)


_FP8_DTYPE = {"e4m3": torch.float8_e4m3fn, "e5m2": torch.float8_e5m2}
_FP8_OP = {("e4m3", "e4m3"): _C.GEMM_E4M3, ("e5m2", "e5m2"): _C.GEMM_E5M2,
           ("e4m3", "e5m2"): _C.GEMM_E4M3_E5M2, ("e5m2", "e4m3"): _C.GEMM_E5M2_E4M3}


def _fp8_codes(t, kind):
    return t.contiguous().to(_FP8_DTYPE[kind]).view(torch.uint8)   # exact: t holds values of that format


class _LinearFp8Fn(torch.autograd.Function):
    """x holds values of an fp8 format exactly (it left a bare e4m3/e5m2 fake-quantizer), the weight is quantized
    straight to codes: both operands go to the FP8 tensor cores.  Products and fp32 accumulation are those of
    the bf16 path (every fp8 value is a bf16 value), at twice the MMA rate and 3/4 of the weight traffic.
    Backward: the activation is saved as codes (1 B / element).  When the incoming gradient is itself fp8-valued
    (`g_kind`: the module's error_pre_process hook is a bare e4m3 / e5m2 quantizer) dgrad and wgrad run on the FP8
    tensor cores too (QT_GEMM_E5M2_E4M3: gradient x weight, gradient^T x activation, operands read MN-major);
    otherwise the codes are widened to bf16 once and the bf16 kernel is used."""

    @staticmethod
    def forward(ctx, x, weight, bias, wq, x_kind, codes, g_kind):
        x2 = x.reshape(-1, x.shape[-1])
        xc = _fp8_codes(x2, x_kind)
        wc = codes if codes is not None else wq.quantize_to_codes(weight)
        y = _C.gemm_nt(xc, wc, bias=bias.contiguous() if bias is not None else None,
                       operand_type=_FP8_OP[(x_kind, wq.fp8_kind)])
        if any(ctx.needs_input_grad[:3]):
            ctx.save_for_backward(xc, wc)
            ctx.kinds = (x_kind, wq.fp8_kind, g_kind)
            ctx.x_shape = x.shape
        ctx.has_bias = bias is not None
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, g):
        xc, wc = ctx.saved_tensors
        x_kind, w_kind, g_kind = ctx.kinds
        g2 = g.reshape(-1, g.shape[-1])
        n, k = wc.shape
        m = g2.shape[0]
        gx = gw = None
        fp8_ok = g_kind is not None and n % 16 == 0 and k % 16 == 0
        if fp8_ok:
            gc = _fp8_codes(g2, g_kind)
            if ctx.needs_input_grad[0]:
                gx = _C.gemm_nt(gc, wc, operand_type=_FP8_OP[(g_kind, w_kind)], b_mn=True).view(ctx.x_shape)
            if ctx.needs_input_grad[1]:   # STE through the weight quantizer
                gw = _C.gemm_nt(gc, xc, operand_type=_FP8_OP[(g_kind, x_kind)], a_mn=True, b_mn=True)
        else:
            if ctx.needs_input_grad[0]:
                gx = gemm(g2, wc.view(_FP8_DTYPE[w_kind]).to(g.dtype).t()).view(ctx.x_shape)
            if ctx.needs_input_grad[1]:
                gw = gemm(g2.t(), xc.view(_FP8_DTYPE[x_kind]).to(g.dtype).t())
        gb = g2.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None, None


def linear_fp8(x, weight, bias, weight_fq, x_kind, codes=None, g_kind=None):
    """F.linear(x, weight_fq(weight), bias) with both operands as fp8 codes.  Caller guarantees that x already holds
    `x_kind` values exactly and that weight_fq is a bare (scale 1) e4m3/e5m2 quantizer; g_kind: the fp8 format the
    gradient of the output will hold exactly (bare error quantizer), or None."""
    return _LinearFp8Fn.apply(x, weight, bias, weight_fq, x_kind, codes, g_kind)


def _bare_fp8_kind(fq):
    if fq is None or getattr(fq, "fp8_kind", None) is None or fq.qscheme is not None or fq.is_per_channel:
        return None
    return fq.fp8_kind if fq._flags() == (False, True) else None


def fp8_route(module, x, weight_fq):
    """Which fp8 format the input of `module` is already quantized to by its forward pre-hook, if the whole
    product can run on the FP8 tensor cores; else None."""
    if not (_ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[-1] % 16 == 0):
        return None
    hooks = getattr(module, "activation_pre_process", None)
    act = hooks["0"] if hooks is not None and "0" in hooks else None
    if _bare_fp8_kind(act) is None or _bare_fp8_kind(weight_fq) is None or module.weight.shape[0] % 8:
        return None
    return act.fp8_kind


def fp8_grad_kind(module):
    """fp8 format of the gradient arriving at `module`'s output (its error_pre_process hook is a bare e4m3 / e5m2
    quantizer, quantize.py:142-150), or None."""
    hooks = getattr(module, "error_pre_process", None)
    err = hooks["0"] if hooks is not None and "0" in hooks else None
    return _bare_fp8_kind(err)


class _MatmulFn(torch.autograd.Function):
    """x [..., M, K] @ y [..., K, N].  Logical B = y^T: K-major when y is itself a transposed view (k^T in attention),
    MN-major when y is row-major (P V) -- read as it lies either way.  Backward: gx = g y^T (B = y), gy = x^T g."""

    @staticmethod
    def forward(ctx, x, y):
        out = gemm(x, y.transpose(-1, -2))
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        gx = gemm(g, y) if ctx.needs_input_grad[0] else None
        gy = gemm(x.transpose(-1, -2), g.transpose(-1, -2)) if ctx.needs_input_grad[1] else None
        return gx, gy


def matmul(x, y):
    if not _ENABLED:
        return torch.matmul(x, y)
    _require_bf16_cuda(x, y)
    if x.dim() < 2 or y.dim() < 2:
        raise ValueError("quantized_training.ops.matmul takes operands with at least 2 dimensions")
    if x.shape[:-2] != y.shape[:-2]:   # broadcast batch dimensions like torch.matmul (views, no copies)
        lead = torch.broadcast_shapes(x.shape[:-2], y.shape[:-2])
        x, y = x.expand(*lead, *x.shape[-2:]), y.expand(*lead, *y.shape[-2:])
    return _MatmulFn.apply(x, y)
