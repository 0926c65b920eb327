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
import os
import weakref

import torch

from . import _C
from .fake_quantize import _block_view

__all__ = ["vmap", "quantize", "dequantize", "expand", "calculate_mx_qparam", "quantize_mx", "linear_mx", "matmul_mx",
           "conv2d_mx"]

_lib = torch.library.Library("quantized_ops", "DEF")
_lib.define("vmap(Tensor self, Tensor other) -> Tensor")
_lib.define("quantize(Tensor input, Tensor scale, Tensor? zero_point=None, SymInt[]? axes=None, "
            "int? block_size=None, Tensor? qmap=None, Tensor? output_code=None) -> Tensor")
_lib.define("dequantize(Tensor input, Tensor scale, Tensor? zero_point=None, SymInt[]? axes=None, "
            "int? block_size=None, Tensor? input_qmap=None, Tensor? output_qmap=None) -> Tensor")
_lib.define("calculate_mx_qparam(Tensor self, SymInt[] axes, int block_size, float quant_max, "
            "bool force_scale_power_of_two=False, Tensor scale_qmap=None) -> Tensor")
_lib.define("quantize_mx(Tensor self, Tensor qmap, SymInt[] axes, int block_size, float quant_max, "
            "bool force_scale_power_of_two=False, Tensor scale_qmap=None, Tensor output_code=None) -> (Tensor, Tensor)")
_lib.define("linear_mx(Tensor input, Tensor weight, Tensor? bias=None, *, Tensor? input_scale=None, "
            "Tensor? weight_scale=None, int? block_size=None, Tensor? input_code=None, "
            "Tensor? weight_code=None) -> Tensor")
_lib.define("conv2d_mx(Tensor input, Tensor weight, Tensor? bias=None, SymInt[2] stride=1, SymInt[2] padding=0, "
            "SymInt[2] dilation=1, SymInt groups=1, *, Tensor? input_scale=None, Tensor? weight_scale=None, "
            "int? block_size=None, Tensor? input_code=None, Tensor? weight_code=None) -> Tensor")
_lib.define("matmul_mx(Tensor self, Tensor other, *, Tensor? input_scale=None, Tensor? weight_scale=None, "
            "int? block_size=None, Tensor? input_code=None, Tensor? weight_code=None) -> Tensor")


def expand(input, shape, block_size):
    """Per-block parameters -> one value per element of a tensor of `shape` (the host-side helper drivers import from
    the reference, decomposed.py:127-140): leading axes are added, every axis whose extent differs is repeated
    block_size times and cut back to the tensor's extent (edge blocks)."""
    t = input.reshape((1,) * (len(shape) - input.ndim) + tuple(input.shape))
    for d, want in enumerate(shape):
        if t.shape[d] != want:
            t = t.repeat_interleave(block_size, dim=d).narrow(d, 0, want)
    return t


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


_POW2_TABLES = {}


def calculate_mx_qparam(input, axes, block_size, quant_max, force_scale_power_of_two=False, scale_qmap=None):
    """Per-block scale of the microscaling scheme (decomposed.py:372-419): amax / quant_max [through scale_qmap], or
    2^(floor(log2 amax) - floor(log2 quant_max)); non-positive / NaN -> 1.  Shape: the block grid; dtype: input's."""
    if not input.is_cuda:
        raise RuntimeError("quantized_ops: the B200 build runs on CUDA tensors only (no CPU fallback)")
    if input.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f"quantized_ops kernels take bfloat16 / float32 tensors, got {input.dtype}")
    x = input.contiguous()
    axes = [axes] if isinstance(axes, int) else list(axes)
    dims, axis2, grid = _block_view(tuple(x.shape), axes, block_size)
    scale = torch.empty(grid, dtype=torch.float32, device=x.device)
    pow2_tab = None
    if force_scale_power_of_two:
        key = (x.dtype, x.device)
        if key not in _POW2_TABLES:
            _POW2_TABLES[key] = _C.pow2_table_host(_C.QT_F32 if x.dtype == torch.float32 else _C.QT_BF16).to(x.device)
        pow2_tab = _POW2_TABLES[key]
    if x.numel():
        _C.fq_block(x, None, dims, block_size, axis2, _C.BLOCK_MX, -float(quant_max), float(quant_max),
                    _C.format_from_string("bfloat16"), scale, force_pow2=force_scale_power_of_two,
                    pow2_table=pow2_tab, scale_table=_table(scale_qmap, x.device))
    return scale.to(input.dtype)


def quantize_mx(input, qmap, axes, block_size, quant_max, force_scale_power_of_two=False, scale_qmap=None,
                output_code=None):
    """(scale, vmap(input / expand(scale), qmap)) -- decomposed.py:428-448."""
    scale = calculate_mx_qparam(input, axes, block_size, quant_max, force_scale_power_of_two, scale_qmap)
    return scale, quantize(input, scale, None, axes, block_size, qmap)


def _decode(t, scale, block_size, code):
    """codebook decode + block-scale multiply of one *_mx operand (decomposed.py:291-300)."""
    if code is not None:
        t = code[t.to(torch.long)].to(t.dtype)
    if scale is not None:
        t = dequantize(t, scale, None, None, block_size)
    return t


# ---- block-scaled tensor-core route of linear_mx ----------------------------------------------------------------
# OCP microscaling operands (fp8 / fp6 / fp4 element values, one power-of-two scale per 32 elements of K) are what
# tcgen05.mma kind::mxf8f6f4.block_scale multiplies natively: the elements travel as one-byte fp8 codes (fp6 / fp4 values
# are exact e4m3 values), the scales as UE8M0 exponent bytes in TMEM, and nothing is dequantized to bf16 in HBM.
# Whether a call qualifies is a property of its DATA (element values on an fp8 grid, scales powers of two), so it is
# checked on the device and read back once per call; QT_MX_TENSOR_CORES=0 turns the route off, =assume skips the check of
# the activations (the caller guarantees MX-formatted operands; also the only mode usable under CUDA-graph capture).
MX_TENSOR_CORES = os.environ.get("QT_MX_TENSOR_CORES", "1")
_MX_WEIGHTS = {}
_F8 = ((torch.float8_e4m3fn, "e4m3"), (torch.float8_e5m2, "e5m2"))
_MX_TYPES = {("e4m3", "e4m3"): _C.GEMM_E4M3, ("e5m2", "e5m2"): _C.GEMM_E5M2, ("e4m3", "e5m2"): _C.GEMM_E4M3_E5M2,
             ("e5m2", "e4m3"): _C.GEMM_E5M2_E4M3}


def _mx_operand(t2, scale2, check=True, transposed=False):
    """One operand ([rows, K], or a batch [..., rows, K]; transposed: [..., K, rows] with scales [..., K / 32, rows]) as
    (fp8 kind, one-byte codes in the same layout, packed scales), or None when its elements are on neither fp8 grid or a
    scale is not a power of two.  check=False: e4m3 codes without looking (the caller's guarantee)."""
    ok = torch.ones(1, dtype=torch.int32, device=t2.device)
    packed = _C.mx_pack_scales(scale2.float().contiguous(), ok, transposed)
    for dt, kind in _F8:
        c = t2.to(dt)
        if not check:
            return kind, c.view(torch.uint8), packed
        fits, scales_ok = torch.stack([(c.to(t2.dtype) == t2).all(), ok[0] != 0]).tolist()   # the read-back
        if not scales_ok:
            return None
        if fits:
            return kind, c.view(torch.uint8), packed
    return None


def _linear_mx_tensor_cores(input, weight, bias, input_scale, weight_scale, block_size):
    """The block-scaled product, or None when the call does not qualify (the caller then dequantizes)."""
    mode = MX_TENSOR_CORES
    if mode == "0" or block_size != 32 or input_scale is None or weight_scale is None:
        return None
    if not (input.is_cuda and input.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.dim() == 2):
        return None
    N, K = weight.shape
    kb = (K + 31) // 32
    if K % 16 or N % 8 or input.shape[-1] != K or input.numel() == 0:
        return None
    if tuple(weight_scale.shape) != (N, kb) or tuple(input_scale.shape) != (*input.shape[:-1], kb):
        return None
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing and mode != "assume":
        return None
    x2 = input.reshape(-1, K)
    # cached per weight OBJECT (weak references: an address alone may be reused by a different tensor after a free)
    key = (id(weight), id(weight_scale))
    w = _MX_WEIGHTS.get(key)
    if w is not None and not (w[3]() is weight and w[4]() is weight_scale and
                              w[5] == (weight._version, weight_scale._version, weight.data_ptr(), weight_scale.data_ptr())):
        w = None
    if w is None:
        if capturing:
            return None
        w = (_mx_operand(weight, weight_scale), None, None, weakref.ref(weight), weakref.ref(weight_scale),
             (weight._version, weight_scale._version, weight.data_ptr(), weight_scale.data_ptr()))
        if len(_MX_WEIGHTS) >= 256:
            _MX_WEIGHTS.clear()
        _MX_WEIGHTS[key] = w
    if w[0] is None:
        return None
    w_kind, w_codes, w_sf = w[0]
    a = _mx_operand(x2, input_scale.reshape(-1, kb), check=mode != "assume")
    if a is None:
        return None
    a_kind, a_codes, packed = a
    if bias is not None:
        bias = bias.to(torch.bfloat16).contiguous()
    y = _C.gemm_nt(a_codes, w_codes, operand_type=_MX_TYPES[(a_kind, w_kind)], bias=bias, sf_a=packed, sf_b=w_sf)
    return y.reshape(*input.shape[:-1], N)


def linear_mx(input, weight, bias=None, *, input_scale=None, weight_scale=None, block_size=None, input_code=None,
              weight_code=None):
    """F.linear on the dequantized operands (decomposed.py:311-331).  Microscaling operands (block_size 32, power-of-two
    scales, fp8-representable elements) are multiplied by the block-scaled tensor-core instruction without being
    dequantized; everything else is dequantized first and runs on the bf16 tcgen05 GEMM."""
    from . import ops
    if input_code is None and weight_code is None:
        y = _linear_mx_tensor_cores(input, weight, bias, input_scale, weight_scale, block_size)
        if y is not None:
            return y
    return ops.linear(_decode(input, input_scale, block_size, input_code),
                      _decode(weight, weight_scale, block_size, weight_code), bias)


def _matmul_mx_tensor_cores(a, b, a_scale, b_scale, block_size):
    """a [..., M, K] @ b [..., K, N] with scales [..., M, K / 32] and [..., K / 32, N] as ONE batched block-scaled product
    (b is read as stored, MN-major), or None when the call does not qualify."""
    mode = MX_TENSOR_CORES
    if mode == "0" or block_size != 32 or a_scale is None or b_scale is None:
        return None
    if not (a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() >= 2 and a.dim() == b.dim()
            and a.dim() <= 4 and a.shape[:-2] == b.shape[:-2]):
        return None
    M, K, N = a.shape[-2], a.shape[-1], b.shape[-1]
    kb = (K + 31) // 32
    if b.shape[-2] != K or K % 16 or N % 16 or a.numel() == 0 or b.numel() == 0:
        return None
    if tuple(a_scale.shape) != (*a.shape[:-2], M, kb) or tuple(b_scale.shape) != (*b.shape[:-2], kb, N):
        return None
    if torch.cuda.is_current_stream_capturing() and mode != "assume":
        return None
    check = mode != "assume"
    oa = _mx_operand(a.contiguous(), a_scale, check)
    if oa is None:
        return None
    ob = _mx_operand(b.contiguous(), b_scale, check, transposed=True)
    if ob is None:
        return None
    return _C.gemm_nt(oa[1], ob[1], operand_type=_MX_TYPES[(oa[0], ob[0])], b_mn=True, sf_a=oa[2], sf_b=ob[2],
                      sf_batched=(True, True))


def matmul_mx(self, other, *, input_scale=None, weight_scale=None, block_size=None, input_code=None,
              weight_code=None):
    """torch.matmul on the dequantized operands (decomposed.py:341-363); microscaling operands (see linear_mx) are
    multiplied by the block-scaled tensor-core instruction, the second one read as it is stored."""
    from . import ops
    if input_code is None and weight_code is None:
        y = _matmul_mx_tensor_cores(self, other, input_scale, weight_scale, block_size)
        if y is not None:
            return y
    return ops.matmul(_decode(self, input_scale, block_size, input_code),
                      _decode(other, weight_scale, block_size, weight_code))


_lib.impl("calculate_mx_qparam", calculate_mx_qparam, "CUDA")
_lib.impl("quantize_mx", quantize_mx, "CUDA")
_lib.impl("linear_mx", linear_mx, "CUDA")
def conv2d_mx(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, *, input_scale=None,
              weight_scale=None, block_size=None, input_code=None, weight_code=None):
    """F.conv2d on the dequantized operands (decomposed.py:273-300): codebook decode and block scales by this library's
    table kernel, the convolution itself by the library call the reference makes too (cuDNN)."""
    return torch.nn.functional.conv2d(_decode(input, input_scale, block_size, input_code),
                                      _decode(weight, weight_scale, block_size, weight_code), bias, stride, padding,
                                      dilation, groups)


_lib.impl("matmul_mx", matmul_mx, "CUDA")
_lib.impl("conv2d_mx", conv2d_mx, "CUDA")
_lib.impl("vmap", vmap, "CUDA")
_lib.impl("quantize", quantize, "CUDA")
_lib.impl("dequantize", dequantize, "CUDA")
