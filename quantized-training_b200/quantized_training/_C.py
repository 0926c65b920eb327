"""ctypes binding of libqt_b200.so (include/qt_b200.h).  This is the ONLY route to compute:
there is no CPU or eager-PyTorch fallback -- if the library is missing, or a tensor is not on
a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# QT_B200_LIB: load another build of the library (A/B runs of kernel changes on the same box)
LIB_PATH = os.environ.get("QT_B200_LIB") or os.path.join(_HERE, "_lib", "libqt_b200.so")

QT_BF16, QT_F32 = 0, 1
QT_NO_LUT = 5
QT_LUT_BYTES = 8192
_ERR = {1: ValueError, 2: ValueError, 3: RuntimeError, 4: ValueError}


class QtFormat(ctypes.Structure):
    """qt_format_t"""
    _fields_ = [
        ("kind", ctypes.c_int32), ("flavour", ctypes.c_int32), ("nbits", ctypes.c_int32),
        ("ebits", ctypes.c_int32), ("mbits", ctypes.c_int32), ("is_unsigned", ctypes.c_int32),
        ("max_value", ctypes.c_float), ("min_value", ctypes.c_float),
    ]


_lib = None
# every symbol include/qt_b200.h declares; tests check the .so exports all of them
EXPORTS = {
    "qt_version": (ctypes.c_char_p, []),
    "qt_stream_capture_id": (ctypes.c_ulonglong, [ctypes.c_void_p]),
    "qt_last_error": (ctypes.c_char_p, []),
    "qt_format_from_string": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(QtFormat)]),
    "qt_format_min_max": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double)]),
    "qt_table_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_void_p]),
    "qt_lut_build_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_void_p]),
    "qt_scale_pow2_host": (ctypes.c_float, [ctypes.c_float]),
    "qt_scale_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_void_p]),
    "qt_fq_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_mx_pack_scales": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "qt_mx_pack_scales_ex": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_quantize_codes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.POINTER(QtFormat), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "qt_gemm_nt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] +
                   [ctypes.c_int64] * 10 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "qt_gemm_nt_ex": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qt_softmax_fq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                     ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_causal_mask_check": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "qt_norm_fq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                  ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(QtFormat), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p]),
    "qt_add_norm_fq": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(QtFormat), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "qt_act_mul_fq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_size_t] * 5 +
                      [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p,
                       ctypes.c_void_p, ctypes.c_void_p]),
    "qt_code_table_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_int, ctypes.c_void_p]),
    "qt_encode_codes_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_size_t]),
    "qt_quantize_codes8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.POINTER(QtFormat), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "qt_lora_merge_fq": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "qt_rope_fq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                  ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                  ctypes.c_int, ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_fq_transpose": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 +
                        [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(QtFormat),
                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_fq_block": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qt_block_pow2_table_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "qt_table_op": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "qt_amax": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python quantized-training_b200/build.py, or __graft_entry__.build()). "
                "quantized_training (B200) has no CPU / eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().qt_last_error().decode()
        raise _ERR.get(rc, RuntimeError)(msg)


def version():
    return lib().qt_version().decode()


def format_from_string(dtype):
    """Parse a dtype string; ValueError("Unsupported dtype: ...") like the reference (fake_quantize.py:95)."""
    fmt = QtFormat()
    if not isinstance(dtype, str):
        raise ValueError(f"Unsupported dtype: {dtype}")
    _check(lib().qt_format_from_string(dtype.encode(), ctypes.byref(fmt)))
    return fmt


def format_min_max(dtype):
    lo, hi = ctypes.c_double(), ctypes.c_double()
    if not isinstance(dtype, str):
        raise ValueError(f"Unsupported dtype: {dtype}")
    _check(lib().qt_format_min_max(dtype.encode(), ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def table_host(fmt):
    """The rounding logic evaluated on the host for all 65 536 bf16 patterns -> bf16 CPU tensor."""
    out = torch.empty(65536, dtype=torch.int16)
    _check(lib().qt_table_host(ctypes.byref(fmt), out.data_ptr()))
    return out.view(torch.bfloat16)


def lut_host(fmt):
    """int32[2048] CPU tensor holding the fast-path constants of `fmt` (512 x {p1, p2, d, l} as float32 BIT PATTERNS),
    or None for formats that run on the direct path (int / uint / native dtypes).  Integer on purpose: the table is a
    module buffer, and model.bfloat16() / .half() / .to(dtype) cast every floating-point buffer -- which would turn raw
    rounding constants into garbage -- but leave integer buffers alone."""
    out = torch.empty(QT_LUT_BYTES // 4, dtype=torch.int32)
    rc = lib().qt_lut_build_host(ctypes.byref(fmt), out.data_ptr())
    if rc == QT_NO_LUT:
        return None
    _check(rc)
    return out


def _elem_type(t):
    if t.dtype == torch.bfloat16:
        return QT_BF16
    if t.dtype == torch.float32:
        return QT_F32
    raise TypeError(f"fake-quant kernels take bfloat16 or float32 tensors, got {t.dtype}")


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: the B200 build of quantized_training runs on CUDA only (no CPU fallback)")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_dev = getattr(torch._C, "_cuda_getDevice", None)
_set_dev = getattr(torch._C, "_cuda_setDevice", None)


def _stream(t):
    """cudaStream_t of torch's current stream on t's device (the raw getter: no Stream object per call -- these
    wrappers run ~600 times per fine-tune step when a model executes eagerly)."""
    if _raw_stream is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


class _on:
    """`with torch.cuda.device(t.device)` without its cost when t already lives on the current device."""
    __slots__ = ("idx", "prev")

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = -1

    def __enter__(self):
        if _get_dev is None:
            self.prev = torch.cuda.current_device()
            if self.prev != self.idx:
                torch.cuda.set_device(self.idx)
            else:
                self.prev = -1
            return self
        cur = _get_dev()
        if cur != self.idx:
            self.prev = cur
            _set_dev(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            (_set_dev or torch.cuda.set_device)(self.prev)
        return False


def fq_forward(x, y, outer, channels, inner, fmt, scale=None, amax_out=None, lut=None):
    """y = round_fmt(x / s) * s on the current stream; amax_out[c] max-accumulates max|x|.
    lut: device tensor from lut_host(fmt) (fast path) or None (direct bitwise path)."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and y.is_contiguous() and y.dtype == x.dtype and y.device == x.device
    assert outer * channels * inner == x.numel() == y.numel()
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.device == x.device and scale.numel() == channels \
            and scale.is_contiguous()
    if amax_out is not None:
        assert amax_out.dtype == torch.float32 and amax_out.device == x.device and amax_out.numel() >= channels
    if lut is not None:
        assert lut.dtype in (torch.int32, torch.float32) and lut.device == x.device and lut.numel() * 4 == QT_LUT_BYTES \
            and lut.is_contiguous()
    with _on(x):
        _check(lib().qt_fq_forward(x.data_ptr(), y.data_ptr(), outer, channels, inner, _elem_type(x),
                                   ctypes.byref(fmt), scale.data_ptr() if scale is not None else None,
                                   amax_out.data_ptr() if amax_out is not None else None,
                                   lut.data_ptr() if lut is not None else None, _stream(x)))


def quantize_codes(x, codes, fmt, scale=None, amax_out=None, lut=None):
    """codes (uint8, same numel) = fp8 encoding of round_fmt(x / s); per tensor."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and codes.is_contiguous() and codes.dtype == torch.uint8 and codes.numel() == x.numel()
    assert lut is not None and lut.device == x.device
    with _on(x):
        _check(lib().qt_quantize_codes(x.data_ptr(), codes.data_ptr(), x.numel(), _elem_type(x), ctypes.byref(fmt),
                                       scale.data_ptr() if scale is not None else None,
                                       amax_out.data_ptr() if amax_out is not None else None, lut.data_ptr(),
                                       _stream(x)))


CODE_NATIVE, CODE_E4M3, CODE_E5M2 = 0, 1, 2


def code_table_host(fmt, code_kind=CODE_NATIVE):
    """bf16 tensor [256]: decode(byte) of the format's one-byte codes (ValueError if it has none of that kind)."""
    t = torch.empty(256, dtype=torch.int16)
    _check(lib().qt_code_table_host(ctypes.byref(fmt), int(code_kind), t.data_ptr()))
    return t.view(torch.bfloat16)


def encode_codes_host(fmt, bf16_bits, code_kind=CODE_NATIVE):
    """HOST: uint8 codes of round_fmt(value) for an int16/uint16 numpy-or-torch array of bf16 bit patterns."""
    b = torch.as_tensor(bf16_bits).contiguous().view(torch.int16)
    out = torch.empty(b.numel(), dtype=torch.uint8)
    _check(lib().qt_encode_codes_host(ctypes.byref(fmt), int(code_kind), b.data_ptr(), out.data_ptr(), b.numel()))
    return out


def quantize_codes8(x, codes, fmt, code_kind=CODE_NATIVE, scale=None, amax_out=None):
    """codes (uint8, same numel) = encode(round_fmt(x / s)) for any <= 8-bit format; per tensor."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and codes.is_contiguous() and codes.dtype == torch.uint8 and codes.numel() == x.numel()
    with _on(x):
        _check(lib().qt_quantize_codes8(x.data_ptr(), codes.data_ptr(), x.numel(), _elem_type(x), ctypes.byref(fmt),
                                        int(code_kind), _ptr_or_none(scale), _ptr_or_none(amax_out), _stream(x)))
    return codes


def amax(x, outer, channels, inner, amax_out):
    _require_cuda(x, "input")
    assert x.is_contiguous() and amax_out.dtype == torch.float32 and amax_out.device == x.device
    with _on(x):
        _check(lib().qt_amax(x.data_ptr(), outer, channels, inner, _elem_type(x), amax_out.data_ptr(), _stream(x)))


def scale_update(history, ahl, channels, scale, quant_max, force_pow2):
    _require_cuda(history, "amax_history")
    assert history.dtype == torch.float32 and scale.dtype == torch.float32 and history.is_contiguous()
    assert history.numel() == ahl * channels and scale.numel() == channels and scale.device == history.device
    with _on(history):
        _check(lib().qt_scale_update(history.data_ptr(), ahl, channels, scale.data_ptr(), float(quant_max),
                                     int(bool(force_pow2)), _stream(history)))


BLOCK_MX, BLOCK_AFFINE = 0, 1
QT_POW2_TABLE_WORDS = 288


class QtBlockDesc(ctypes.Structure):
    """qt_block_desc_t"""
    _fields_ = [
        ("x", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("elem_type", ctypes.c_int32), ("qscheme", ctypes.c_int32),
        ("d0", ctypes.c_int64), ("n1", ctypes.c_int64), ("d1", ctypes.c_int64), ("n2", ctypes.c_int64),
        ("d2", ctypes.c_int64),
        ("block_size", ctypes.c_int32), ("block_axis2", ctypes.c_int32),
        ("quant_min", ctypes.c_float), ("quant_max", ctypes.c_float),
        ("force_scale_power_of_two", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("fmt", ctypes.POINTER(QtFormat)), ("lut", ctypes.c_void_p),
        ("scale_fmt", ctypes.POINTER(QtFormat)), ("pow2_table", ctypes.c_void_p),
        ("scale", ctypes.c_void_p), ("zero_point", ctypes.c_void_p), ("scale_table", ctypes.c_void_p),
    ]


def pow2_table_host(elem_type):
    """int32[288] CPU tensor: floor(log2(amax)) thresholds in the tensor's dtype (force_scale_power_of_two)."""
    out = torch.empty(QT_POW2_TABLE_WORDS, dtype=torch.int32)
    _check(lib().qt_block_pow2_table_host(int(elem_type), out.data_ptr()))
    return out


def fq_block(x, y, dims, block_size, block_axis2, qscheme, quant_min, quant_max, fmt, scale, zero_point=None,
             lut=None, scale_fmt=None, force_pow2=False, pow2_table=None, scale_table=None):
    """Block-scaled fake quant (qt_fq_block).  dims = (d0, n1, d1, n2, d2) view of the contiguous x;
    scale / zero_point: float32 outputs, one entry per block.  y = None: parameters only.  scale_table: the parameter
    codebook as a 65 536-entry bf16 table instead of scale_fmt."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and (y is None or (y.is_contiguous() and y.dtype == x.dtype and y.device == x.device))
    d0, n1, d1, n2, d2 = (int(v) for v in dims)
    assert d0 * n1 * d1 * n2 * d2 == x.numel() and (y is None or y.numel() == x.numel())
    nb1 = -(-n1 // block_size)
    nb2 = -(-n2 // block_size) if block_axis2 else n2
    for t in (scale, zero_point):
        if t is not None:
            assert t.dtype == torch.float32 and t.device == x.device and t.is_contiguous() \
                and t.numel() == d0 * nb1 * d1 * nb2 * d2
    d = QtBlockDesc()
    d.x, d.y = x.data_ptr(), (y.data_ptr() if y is not None else None)
    d.elem_type, d.qscheme = _elem_type(x), int(qscheme)
    if scale_table is not None:
        assert scale_table.numel() == 65536 and scale_table.element_size() == 2 and scale_table.device == x.device
        d.scale_table = scale_table.data_ptr()
    d.d0, d.n1, d.d1, d.n2, d.d2 = d0, n1, d1, n2, d2
    d.block_size, d.block_axis2 = int(block_size), int(bool(block_axis2))
    d.quant_min, d.quant_max = float(quant_min), float(quant_max)
    d.force_scale_power_of_two = int(bool(force_pow2))
    d.fmt = ctypes.pointer(fmt) if fmt is not None else None
    d.lut = lut.data_ptr() if lut is not None else None
    d.scale_fmt = ctypes.pointer(scale_fmt) if scale_fmt is not None else None
    if pow2_table is not None:
        assert pow2_table.device == x.device and pow2_table.numel() == QT_POW2_TABLE_WORDS
        d.pow2_table = pow2_table.data_ptr()
    d.scale = scale.data_ptr()
    d.zero_point = zero_point.data_ptr() if zero_point is not None else None
    with _on(x):
        _check(lib().qt_fq_block(ctypes.byref(d), _stream(x)))


class QtTableOpDesc(ctypes.Structure):
    """qt_table_op_desc_t"""
    _fields_ = [
        ("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("elem_type", ctypes.c_int32), ("op", ctypes.c_int32),
        ("d0", ctypes.c_int64), ("n1", ctypes.c_int64), ("d1", ctypes.c_int64), ("n2", ctypes.c_int64),
        ("d2", ctypes.c_int64),
        ("block_size", ctypes.c_int32), ("block_axis2", ctypes.c_int32), ("scalar_params", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("scale", ctypes.c_void_p), ("zero_point", ctypes.c_void_p), ("table_a", ctypes.c_void_p),
        ("table_b", ctypes.c_void_p),
    ]


TABLE_QUANTIZE, TABLE_DEQUANTIZE, TABLE_LOOKUP = 0, 1, 2


def _ptr_or_none(t):
    return t.data_ptr() if t is not None else None


def table_op(op, x, y, dims, block_size, block_axis2, scale, zero_point=None, table_a=None, table_b=None):
    """qt_table_op: quantize / dequantize through caller-supplied 65 536-entry bf16 tables (int16 / bf16 tensors).
    x, y, scale, zero_point share one dtype (bf16 or fp32); scale.numel() == 1 means a scalar scale."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and y.is_contiguous() and y.dtype == x.dtype and scale.dtype == x.dtype
    assert scale.is_contiguous() and (zero_point is None or (zero_point.dtype == x.dtype and zero_point.is_contiguous()
                                                             and zero_point.numel() == scale.numel()))
    for t in (table_a, table_b):
        assert t is None or (t.numel() == 65536 and t.element_size() == 2 and t.is_contiguous() and t.device == x.device)
    d = QtTableOpDesc()
    d.x, d.y, d.elem_type, d.op = x.data_ptr(), y.data_ptr(), _elem_type(x), int(op)
    d.d0, d.n1, d.d1, d.n2, d.d2 = (int(v) for v in dims)
    d.block_size, d.block_axis2 = int(block_size), int(bool(block_axis2))
    d.scalar_params = int(scale.numel() == 1)
    d.scale = scale.data_ptr()
    d.zero_point = _ptr_or_none(zero_point)
    d.table_a, d.table_b = _ptr_or_none(table_a), _ptr_or_none(table_b)
    with _on(x):
        _check(lib().qt_table_op(ctypes.byref(d), _stream(x)))


GEMM_BF16, GEMM_E4M3, GEMM_E5M2, GEMM_E4M3_E5M2, GEMM_E5M2_E4M3, GEMM_CODE8_B, GEMM_CODE8_AB = range(7)
ACTIVATIONS = {None: 0, "none": 0, "relu": 1, "gelu": 2, "silu": 3}


class QtGemmDesc(ctypes.Structure):
    """qt_gemm_desc_t"""
    _fields_ = [
        ("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("C", ctypes.c_void_p),
        ("operand_type", ctypes.c_int32), ("activation", ctypes.c_int32),
        ("M", ctypes.c_int64), ("N", ctypes.c_int64), ("K", ctypes.c_int64),
        ("batch_inner", ctypes.c_int64), ("batch_outer", ctypes.c_int64),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64),
        ("strideA_inner", ctypes.c_int64), ("strideA_outer", ctypes.c_int64),
        ("strideB_inner", ctypes.c_int64), ("strideB_outer", ctypes.c_int64),
        ("strideC_inner", ctypes.c_int64), ("strideC_outer", ctypes.c_int64),
        ("alpha", ctypes.c_float), ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("ldr", ctypes.c_int64), ("strideR_inner", ctypes.c_int64), ("strideR_outer", ctypes.c_int64),
        ("fq_fmt", ctypes.POINTER(QtFormat)), ("fq_lut", ctypes.c_void_p),
        ("out_type", ctypes.c_int32), ("glu", ctypes.c_int32),
        ("causal", ctypes.c_int32), ("reserved", ctypes.c_int32), ("causal_flag", ctypes.c_void_p),
        ("a_major", ctypes.c_int32), ("b_major", ctypes.c_int32), ("code_lut", ctypes.c_void_p),
        ("sf_a", ctypes.c_void_p), ("sf_b", ctypes.c_void_p), ("sf_rows_a", ctypes.c_int64), ("sf_rows_b", ctypes.c_int64),
        ("sf_a_batched", ctypes.c_int32), ("sf_b_batched", ctypes.c_int32),
    ]


def _as4d(t, name, align):
    """[..., rows, cols] tensor -> a [outer, inner, rows, cols] view the kernel's 4-D tensor maps can address:
    unit-stride last axis, every other stride a multiple of `align` elements (16 bytes), 16-byte aligned base.
    No copy when the tensor already qualifies (projections viewed as [B, H, S, D], k^T, strided outputs)."""
    if t.dim() < 2:
        raise ValueError(f"{name} must have at least 2 dimensions")
    if t.dim() > 4:
        t = t.reshape(-1, *t.shape[-3:])
    while t.dim() < 4:
        t = t.unsqueeze(0)
    ok = t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and all(
        t.shape[i] == 1 or (t.stride(i) % align == 0 and t.stride(i) > 0) for i in range(3))
    return t if ok else t.contiguous()


def gemm_nt(a, b, alpha=1.0, bias=None, activation=None, residual=None, operand_type=GEMM_BF16, out=None,
            fq=None, out_codes=False, glu=False, causal=0, causal_flag=None, a_mn=False, b_mn=False, code_lut=None,
            sf_a=None, sf_b=None, sf_batched=(False, False)):
    """out[..., m, n] = epilogue(alpha * sum_k a[..., m, k] * b[..., n, k]) on the tcgen05 kernel.
    a, b: bf16 (GEMM_BF16) or uint8 fp8 codes, up to two leading batch dimensions with arbitrary strides;
    bias bf16 [n]; residual bf16 broadcastable to out's shape; out: optional destination (any 16-byte
    aligned strides, e.g. a [B, H, S, D] view of a [B, S, H*D] buffer).
    fq: (fmt, lut) of a BARE fake-quantizer applied to the result in the epilogue (the consumer's input hook);
    out_codes: with an e4m3 / e5m2 `fq`, store one-byte fp8 codes (out is uint8); glu: b is a gate|up projection
    interleaved in blocks of 64 rows and the result is activation(gate) * up with n / 2 columns.
    causal: CAUSAL_OUT_LOWER (skip output tiles above the diagonal) / CAUSAL_A_LOWER (a is lower triangular);
    causal_flag: optional device int32 tensor (causal_mask_check) that gates the schedule on the device.
    a_mn / b_mn: the operand is handed over as it is stored, transposed -- a as [..., k, m], b as [..., k, n] with a
    unit-stride last axis -- and read MN-major by the tensor cores (dgrad / wgrad / x @ y without transpose copies)."""
    _require_cuda(a, "a")
    want_a = torch.bfloat16 if operand_type in (GEMM_BF16, GEMM_CODE8_B) else torch.uint8
    want_b = torch.bfloat16 if operand_type == GEMM_BF16 else torch.uint8
    if a.dtype != want_a or b.dtype != want_b:
        raise TypeError(f"operand_type {operand_type} takes {want_a} x {want_b} operands, got {a.dtype} and {b.dtype}")
    if operand_type in (GEMM_CODE8_B, GEMM_CODE8_AB):
        if code_lut is None or code_lut.numel() != 256 or code_lut.element_size() != 2 or code_lut.device != a.device:
            raise ValueError("GEMM_CODE8* needs code_lut: the format's 256-entry decode table on the device")
    if b.dim() == 2 and a.dim() > 2 and not a_mn:   # one weight for every batch entry: the batch is just more rows
        a = a.reshape(-1, a.shape[-1])
    a4, b4 = _as4d(a, "a", 16 if a.dtype == torch.uint8 else 8), _as4d(b, "b", 16 if b.dtype == torch.uint8 else 8)
    if a4.shape[:2] != b4.shape[:2]:
        raise ValueError(f"batch mismatch: {tuple(a.shape)} x {tuple(b.shape)}")
    outer, inner = a4.shape[:2]
    (K, M) = a4.shape[2:] if a_mn else a4.shape[:1:-1]
    (Kb, N) = b4.shape[2:] if b_mn else b4.shape[:1:-1]
    if Kb != K:
        raise ValueError(f"inner dimensions differ: {tuple(a.shape)} (a_mn={a_mn}) x {tuple(b.shape)} (b_mn={b_mn})")
    lead = a.shape[:-2] if a.dim() >= b.dim() else b.shape[:-2]
    out_shape = (*lead, M, N // 2 if glu else N)
    out_dtype = torch.uint8 if out_codes else torch.bfloat16
    if out_codes and fq is None:
        raise ValueError("fp8 code output needs the fake-quantizer (fq) whose codes they are")
    if out is None:
        out = torch.empty(out_shape, dtype=out_dtype, device=a.device)
    elif tuple(out.shape) != tuple(out_shape) or out.dtype != out_dtype:
        raise ValueError(f"out must be {out_dtype} of shape {tuple(out_shape)}, got {out.dtype} {tuple(out.shape)}")
    o4 = _as4d(out, "out", 16 if out_codes else 8)
    if o4.data_ptr() != out.data_ptr() or o4.numel() != out.numel():
        raise ValueError("out needs a unit-stride last axis and 16-byte aligned strides")
    d = QtGemmDesc()
    d.A, d.B, d.C = a4.data_ptr(), b4.data_ptr(), o4.data_ptr()
    d.operand_type, d.activation = operand_type, ACTIVATIONS[activation]
    d.M, d.N, d.K, d.batch_inner, d.batch_outer = M, N, K, inner, outer
    d.lda, d.ldb, d.ldc = a4.stride(2), b4.stride(2), o4.stride(2)
    d.strideA_outer, d.strideA_inner = a4.stride(0), a4.stride(1)
    d.strideB_outer, d.strideB_inner = b4.stride(0), b4.stride(1)
    d.strideC_outer, d.strideC_inner = o4.stride(0), o4.stride(1)
    d.alpha = float(alpha)
    d.glu = 1 if glu else 0
    d.causal = int(causal)
    d.a_major, d.b_major = int(bool(a_mn)), int(bool(b_mn))
    if code_lut is not None:
        d.code_lut = code_lut.data_ptr()
    if sf_a is not None or sf_b is not None:
        # block-scaled fp8 product: packed UE8M0 scale factors of both operands (mx_pack_scales)
        assert sf_a.dtype == torch.uint8 and sf_b.dtype == torch.uint8 and sf_a.device == a.device == sf_b.device
        # sf_batched: (a, b) -- the operand's scales come one packed set per batch entry (else one set for all)
        k128, nb = (K + 127) // 128, inner * outer
        d.sf_a, d.sf_b = sf_a.data_ptr(), sf_b.data_ptr()
        d.sf_a_batched, d.sf_b_batched = int(bool(sf_batched[0])), int(bool(sf_batched[1]))
        d.sf_rows_a = sf_a.numel() // (4 * k128 * (nb if sf_batched[0] else 1))
        d.sf_rows_b = sf_b.numel() // (4 * k128 * (nb if sf_batched[1] else 1))
    if causal_flag is not None:
        assert causal_flag.dtype == torch.int32 and causal_flag.device == a.device
        d.causal_flag = causal_flag.data_ptr()
    if fq is not None:
        fq_fmt, fq_lut = fq
        d.fq_fmt = ctypes.pointer(fq_fmt)
        d.fq_lut = fq_lut.data_ptr() if fq_lut is not None else None
        d.out_type = _codes_type(fq_fmt) if out_codes else OUT_BF16
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == N and bias.is_contiguous()
        d.bias = bias.data_ptr()
    r4 = None
    if residual is not None:
        r4 = _as4d(residual.expand(out_shape), "residual", 8)
        if r4.dtype != torch.bfloat16:
            raise TypeError("residual must be bf16")
        d.residual, d.ldr, d.strideR_outer, d.strideR_inner = r4.data_ptr(), r4.stride(2), r4.stride(0), r4.stride(1)
    with _on(a):
        _check(lib().qt_gemm_nt_ex(ctypes.addressof(d), _stream(a)))
    return out


def mx_pack_scales(scale, ok=None, transposed=False):
    """fp32 [rows, K / 32] (or a batch [..., rows, K / 32]) power-of-two block scales -> the packed UE8M0 bytes of the
    block-scaled GEMM (qt_mx_pack_scales_ex), one set per batch entry; transposed: the matrices are [K / 32, rows].
    `ok` (int32[1] on the device, preset to 1) is cleared if a scale is not such a power of two."""
    _require_cuda(scale, "scale")
    assert scale.dtype == torch.float32 and scale.dim() >= 2 and scale.is_contiguous()
    rows, kb32 = (scale.shape[-1], scale.shape[-2]) if transposed else (scale.shape[-2], scale.shape[-1])
    batch = scale.numel() // (rows * kb32)
    rows_pad, k128 = (rows + 127) // 128 * 128, (kb32 + 3) // 4
    out = torch.empty(batch * k128 * rows_pad * 4, dtype=torch.uint8, device=scale.device)
    if ok is not None:
        assert ok.dtype == torch.int32 and ok.device == scale.device
    with _on(scale):
        _check(lib().qt_mx_pack_scales_ex(scale.data_ptr(), batch, rows, kb32, int(bool(transposed)), out.data_ptr(),
                                          _ptr(ok), _stream(scale)))
    return out


# ---- fused ops (qt_fused.cu): thin wrappers; argument checking beyond dtype/device lives in the C library --------
FQ_PRE, FQ_MID, FQ_POST = 1, 2, 4
FQ_RES_A, FQ_RES_B = 32, 64
SOFTMAX_CAUSAL = 16
CAUSAL_OUT_LOWER, CAUSAL_A_LOWER = 1, 2
NORM_RMS, NORM_LAYER, NORM_NONE, NORM_IDENTITY = 0, 1, 2, 3
OUT_BF16, OUT_E4M3, OUT_E5M2 = 0, 1, 2


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _bf16_cuda(t, what):
    _require_cuda(t, what)
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{what} must be bfloat16, got {t.dtype}")


def _out_type(out):
    """bf16 destination -> values; uint8 destination -> fp8 codes (the format decides e4m3 / e5m2: see callers)."""
    if out.dtype == torch.bfloat16:
        return OUT_BF16
    if out.dtype != torch.uint8:
        raise TypeError(f"output must be bfloat16 (values) or uint8 (fp8 codes), got {out.dtype}")
    return None


def _codes_type(fmt):
    if fmt.kind == 2 and not fmt.is_unsigned and (fmt.ebits, fmt.mbits) == (4, 3):
        return OUT_E4M3
    if fmt.kind == 2 and not fmt.is_unsigned and (fmt.ebits, fmt.mbits) == (5, 2):
        return OUT_E5M2
    raise ValueError("fp8 code output needs an e4m3 / e5m2 format")


def _resolve_out(out, fmt):
    t = _out_type(out)
    return _codes_type(fmt) if t is None else t


def stream_capture_id(t):
    """0 if the current stream of t's device is not capturing, else the id of the CUDA-graph capture in progress."""
    return int(lib().qt_stream_capture_id(_stream(t)))


def causal_mask_check(mask3, flag=None):
    """int32 device flag: 1 iff the additive bf16 mask [batches, rows, rows] is the standard causal mask.  Asynchronous
    (no host read-back): hand the tensor to gemm_nt(causal_flag=...) / softmax_fq(causal_flag=...)."""
    _bf16_cuda(mask3, "mask")
    assert mask3.is_contiguous() and mask3.dim() == 3 and mask3.shape[1] == mask3.shape[2]
    if flag is None:
        flag = torch.empty(1, dtype=torch.int32, device=mask3.device)
    with _on(mask3):
        _check(lib().qt_causal_mask_check(mask3.data_ptr(), mask3.shape[0], mask3.shape[1], flag.data_ptr(),
                                          _stream(mask3)))
    return flag


def softmax_fq(scores, probs, alpha, mask, rows_per_batch, mask_rows, mask_batches, fq_points, fmt,
               scale_pre=None, scale_mid=None, scale_post=None, lut=None, causal_flag=None):
    _bf16_cuda(scores, "scores")
    assert scores.is_contiguous() and probs.is_contiguous() and probs.shape == scores.shape
    cols = scores.shape[-1]
    with _on(scores):
        _check(lib().qt_softmax_fq(scores.data_ptr(), probs.data_ptr(), scores.numel() // cols, cols, float(alpha),
                                   _ptr(mask), rows_per_batch, mask_rows, mask_batches, fq_points,
                                   _resolve_out(probs, fmt), ctypes.byref(fmt), _ptr(scale_pre), _ptr(scale_mid),
                                   _ptr(scale_post), _ptr(lut), _ptr(causal_flag), _stream(scores)))


def norm_fq(x, y, kind, weight, bias, eps, fq_points, fmt, scale_pre=None, scale_post=None, lut=None, y_raw=None):
    _bf16_cuda(x, "x")
    assert x.is_contiguous() and y.is_contiguous() and weight.is_contiguous() and weight.dtype == torch.bfloat16
    assert y.shape == x.shape
    assert y_raw is None or (y_raw.is_contiguous() and y_raw.shape == x.shape and y_raw.dtype == torch.bfloat16)
    cols = x.shape[-1]
    with _on(x):
        _check(lib().qt_norm_fq(x.data_ptr(), y.data_ptr(), _ptr(y_raw), x.numel() // cols, cols, kind, weight.data_ptr(),
                                _ptr(bias), float(eps), fq_points, _resolve_out(y, fmt), ctypes.byref(fmt),
                                _ptr(scale_pre), _ptr(scale_post), _ptr(lut), _stream(x)))


def add_norm_fq(x, res, y, kind, weight, bias, eps, fq_points, fmt, scale_pre=None, scale_post=None, lut=None, y_raw=None):
    """y = fq_post(norm(fq_pre(bf16(fq_a(x) + fq_b(res)))))  (qt_add_norm_fq)."""
    _bf16_cuda(x, "x")
    _bf16_cuda(res, "res")
    assert x.is_contiguous() and res.is_contiguous() and y.is_contiguous() and res.shape == x.shape and y.shape == x.shape
    assert weight is None or (weight.is_contiguous() and weight.dtype == torch.bfloat16)
    assert y_raw is None or (y_raw.is_contiguous() and y_raw.shape == x.shape and y_raw.dtype == torch.bfloat16)
    cols = x.shape[-1]
    with _on(x):
        _check(lib().qt_add_norm_fq(x.data_ptr(), res.data_ptr(), y.data_ptr(), _ptr(y_raw), x.numel() // cols, cols, kind,
                                    _ptr(weight), _ptr(bias), float(eps), fq_points, _resolve_out(y, fmt),
                                    ctypes.byref(fmt), _ptr(scale_pre), _ptr(scale_post), _ptr(lut), _stream(x)))


def act_mul_fq(gate, up, out, activation, fq_points, fmt, scale_post=None, lut=None):
    """gate, up, out: 2-D views [rows, cols] with a unit-stride last axis (row strides may differ)."""
    _bf16_cuda(gate, "gate")
    assert gate.dim() == 2 and out.shape == gate.shape and gate.stride(1) == 1 and out.stride(1) == 1
    assert up is None or (up.shape == gate.shape and up.stride(1) == 1)
    with _on(gate):
        _check(lib().qt_act_mul_fq(gate.data_ptr(), _ptr(up), out.data_ptr(), gate.shape[0], gate.shape[1],
                                   gate.stride(0), up.stride(0) if up is not None else 0, out.stride(0),
                                   ACTIVATIONS[activation], fq_points, _resolve_out(out, fmt), ctypes.byref(fmt),
                                   _ptr(scale_post), _ptr(lut), _stream(gate)))


def lora_merge_fq(w, a, b, out, scaling, fq_points, fmt, scale_post=None, lut=None):
    """out = fq_post(w + (fq(b) @ fq(a)) * scaling) with the reference's bf16 roundings (qt_lora_merge_fq)."""
    for t, what in ((w, "w"), (a, "a"), (b, "b"), (out, "out")):
        _bf16_cuda(t, what)
        assert t.is_contiguous(), what
    n, k = w.shape
    r = a.shape[0]
    assert a.shape == (r, k) and b.shape == (n, r) and out.shape == w.shape
    with _on(w):
        _check(lib().qt_lora_merge_fq(w.data_ptr(), a.data_ptr(), b.data_ptr(), out.data_ptr(), n, k, r, float(scaling),
                                      int(fq_points), ctypes.byref(fmt), _ptr(scale_post), _ptr(lut), _stream(w)))
    return out


def rope_fq(q, q_out, k, k_out, cos, sin, fq_points, fmt, scale_q=None, scale_k=None, lut=None):
    """q, k: [tokens, heads, head_dim] bf16 views (contiguous heads, any token stride); cos, sin: [rows, head_dim]."""
    _bf16_cuda(q, "q")
    tokens, qh, d = q.shape
    assert q.stride(2) == 1 and q.stride(1) == d and q_out.stride(2) == 1 and q_out.stride(1) == d
    assert cos.is_contiguous() and sin.is_contiguous() and cos.shape == sin.shape and cos.shape[-1] == d
    kh = 0
    if k is not None:
        kh = k.shape[1]
        assert k.shape[0] == tokens and k.stride(2) == 1 and k.stride(1) == d and k_out.stride(1) == d
        assert k_out.dtype == q_out.dtype
    with _on(q):
        _check(lib().qt_rope_fq(q.data_ptr(), q_out.data_ptr(), q.stride(0), q_out.stride(0), qh,
                                _ptr(k), _ptr(k_out), k.stride(0) if k is not None else 0,
                                k_out.stride(0) if k is not None else 0, kh, tokens, d, cos.data_ptr(), sin.data_ptr(),
                                cos.numel() // d, fq_points, _resolve_out(q_out, fmt), ctypes.byref(fmt), _ptr(scale_q),
                                _ptr(scale_k), _ptr(lut), _stream(q)))


def fq_transpose(v, out, fq_points, fmt, scale_post=None, lut=None):
    """v: [B, S, H, D] bf16 view (contiguous heads; token / batch strides free) -> out [B, H, D, S] contiguous."""
    _bf16_cuda(v, "v")
    b, s_, h, d = v.shape
    assert v.stride(3) == 1 and v.stride(2) == d and out.is_contiguous() and tuple(out.shape) == (b, h, d, s_)
    with _on(v):
        _check(lib().qt_fq_transpose(v.data_ptr(), out.data_ptr(), b, s_, h, d, v.stride(1), v.stride(0), fq_points,
                                     _resolve_out(out, fmt), ctypes.byref(fmt), _ptr(scale_post), _ptr(lut),
                                     _stream(v)))
