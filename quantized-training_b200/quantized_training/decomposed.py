"""``torch.ops.quantized_ops.{vmap, quantize, dequantize}`` backed by the sm_100a kernels.

Operator surface of the reference's ``decomposed.py`` (:143-262): the same namespace, op names and schemas, so graphs
and drivers that call ``torch.ops.quantized_ops.quantize(x, scale, zp, axes, block_size, qmap)`` keep working.  Only a
CUDA implementation is registered -- a CPU tensor gets the dispatcher's "not implemented for CPU" error (there is no
CPU fallback in this build).

The reference passes the rounding function as an explicit 65 536-entry table tensor (``qmap``), which may be any
codebook; these ops therefore gather from the caller's table (``qt_table_op``) instead of using the bitwise rounders
of the module path.  Type promotion follows torch: ``input / scale`` of a bf16 tensor and a multi-element fp32 scale
is fp32 (and the lookup then truncates with round-to-odd, decomposed.py:151-153); 0-dim operands do not promote.
"""
import torch

from . import _C
from .fake_quantize import _block_view

__all__ = ["vmap", "quantize", "dequantize", "expand"]

_lib = torch.library.Library("quantized_ops", "DEF")
_lib.define("vmap(Tensor self, Tensor other) -> Tensor")
_lib.define("quantize(Tensor input, Tensor scale, Tensor? zero_point=None, SymInt[]? axes=None, "
            "int? block_size=None, Tensor? qmap=None, Tensor? output_code=None) -> Tensor")
_lib.define("dequantize(Tensor input, Tensor scale, Tensor? zero_point=None, SymInt[]? axes=None, "
            "int? block_size=None, Tensor? input_qmap=None, Tensor? output_qmap=None) -> Tensor")


def expand(input, shape, block_size):
    """decomposed.py:127-140 (host-side helper of the reference, kept for drivers that import it)."""
    while input.ndim < len(shape):
        input = input.unsqueeze(0)
    for dim in range(len(shape)):
        if input.shape[dim] != shape[dim]:
            input = torch.repeat_interleave(input, block_size, dim)
    if list(input.shape) != list(shape):
        input = input[tuple(slice(0, x) for x in shape)]
    return input


def _table(t, device):
    if t is None:
        return None
    if t.numel() != 65536:
        raise ValueError(f"qmap must have 65536 entries (one per bf16 bit pattern), got {t.numel()}")
    return t.to(device=device, dtype=torch.bfloat16).contiguous()


def _params_layout(x, scale, axes, block_size):
    """(dims, block_size, block_axis2, scale laid out for the kernel) for a scale that is a scalar, a block grid
    (block_size given: `expand` semantics) or any tensor broadcastable to x (materialised per element)."""
    if scale.numel() == 1:
        return (1, x.numel(), 1, 1, 1), 1, False, None
    if block_size is not None:
        shape = tuple(x.shape)
        # the axes whose extent differs are the tiled ones (expand(), decomposed.py:131-134)
        sc = scale
        while sc.dim() < len(shape):
            sc = sc.unsqueeze(0)
        tiled = [d for d in range(len(shape)) if sc.shape[d] != shape[d]]
        if tiled and len(tiled) <= 2:
            dims, axis2, grid = _block_view(shape, tiled, block_size)
            if list(sc.shape) == list(grid):
                return dims, block_size, axis2, None
        full = expand(scale, shape, block_size)
        return (1, x.numel(), 1, 1, 1), 1, False, full
    return (1, x.numel(), 1, 1, 1), 1, False, scale.expand(x.shape)


def _run(op, input, scale, zero_point, axes, block_size, table_a, table_b):
    if not input.is_cuda:
        raise RuntimeError("quantized_ops: the B200 build runs on CUDA tensors only (no CPU fallback)")
    dt = input.dtype
    for p in (scale, zero_point):
        if p is not None and p.dim() > 0:
            dt = torch.promote_types(dt, p.dtype)
    if dt not in (torch.bfloat16, torch.float32):
        raise TypeError(f"quantized_ops kernels take bfloat16 / float32 tensors, got {dt}")
    x = input.to(dt).contiguous()
    dims, bs, axis2, full = _params_layout(x, scale, axes, block_size)

    def lay(p):
        if p is None:
            return None
        if p.numel() == 1:
            return p.reshape(1).to(dt).contiguous()
        src = p if full is None else (expand(p, tuple(x.shape), block_size) if block_size is not None
                                      else p.expand(x.shape))
        return src.to(dt).contiguous()

    s, z = lay(scale), lay(zero_point)
    if z is not None and z.numel() != s.numel():
        raise ValueError("scale and zero_point must have the same layout")
    y = torch.empty_like(x)
    if x.numel():
        _C.table_op(op, x, y, dims, bs, axis2, s, z, _table(table_a, x.device), _table(table_b, x.device))
    return y


def vmap(input, qmap):
    """out = qmap[bits(input)] (decomposed.py:146-163): contiguous, input's dtype."""
    one = torch.ones(1, dtype=input.dtype, device=input.device)
    return _run(_C.TABLE_LOOKUP, input, one, None, None, None, qmap, None)


def quantize(input, scale, zero_point=None, axes=None, block_size=None, qmap=None, output_code=None):
    """vmap(input / expand(scale) [+ expand(zero_point)], qmap) (decomposed.py:171-210)."""
    assert qmap is not None, "qmap must be provided for quantization"
    return _run(_C.TABLE_QUANTIZE, input, scale, zero_point, axes, block_size, qmap, None)


def dequantize(input, scale, zero_point=None, axes=None, block_size=None, input_qmap=None, output_qmap=None):
    """[vmap(., input_qmap)] -> (. - zero_point) * scale -> [vmap(., output_qmap)] (decomposed.py:218-262)."""
    return _run(_C.TABLE_DEQUANTIZE, input, scale, zero_point, axes, block_size, input_qmap, output_qmap)


_lib.impl("vmap", vmap, "CUDA")
_lib.impl("quantize", quantize, "CUDA")
_lib.impl("dequantize", dequantize, "CUDA")
