"""ctypes binding of libqt_b200.so (include/qt_b200.h).  This is the ONLY route to compute:
there is no CPU or eager-PyTorch fallback -- if the library is missing, or a tensor is not on
a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libqt_b200.so")

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
    "qt_last_error": (ctypes.c_char_p, []),
    "qt_format_from_string": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(QtFormat)]),
    "qt_format_min_max": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double)]),
    "qt_table_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_void_p]),
    "qt_lut_build_host": (ctypes.c_int, [ctypes.POINTER(QtFormat), ctypes.c_void_p]),
    "qt_scale_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_void_p]),
    "qt_fq_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(QtFormat), ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qt_quantize_codes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.POINTER(QtFormat), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "qt_gemm_nt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] +
                   [ctypes.c_int64] * 10 + [ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
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
    """float32[2048] CPU tensor with the fast-path constants of `fmt` (512 x {p1, p2, d, l}), or None for
    formats that run on the direct path (int / uint / native dtypes)."""
    out = torch.empty(QT_LUT_BYTES // 4, dtype=torch.float32)
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


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


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
        assert lut.dtype == torch.float32 and lut.device == x.device and lut.numel() * 4 == QT_LUT_BYTES \
            and lut.is_contiguous()
    with torch.cuda.device(x.device):
        _check(lib().qt_fq_forward(x.data_ptr(), y.data_ptr(), outer, channels, inner, _elem_type(x),
                                   ctypes.byref(fmt), scale.data_ptr() if scale is not None else None,
                                   amax_out.data_ptr() if amax_out is not None else None,
                                   lut.data_ptr() if lut is not None else None, _stream(x)))


def quantize_codes(x, codes, fmt, scale=None, amax_out=None, lut=None):
    """codes (uint8, same numel) = fp8 encoding of round_fmt(x / s); per tensor."""
    _require_cuda(x, "input")
    assert x.is_contiguous() and codes.is_contiguous() and codes.dtype == torch.uint8 and codes.numel() == x.numel()
    assert lut is not None and lut.device == x.device
    with torch.cuda.device(x.device):
        _check(lib().qt_quantize_codes(x.data_ptr(), codes.data_ptr(), x.numel(), _elem_type(x), ctypes.byref(fmt),
                                       scale.data_ptr() if scale is not None else None,
                                       amax_out.data_ptr() if amax_out is not None else None, lut.data_ptr(),
                                       _stream(x)))


def amax(x, outer, channels, inner, amax_out):
    _require_cuda(x, "input")
    assert x.is_contiguous() and amax_out.dtype == torch.float32 and amax_out.device == x.device
    with torch.cuda.device(x.device):
        _check(lib().qt_amax(x.data_ptr(), outer, channels, inner, _elem_type(x), amax_out.data_ptr(), _stream(x)))


def scale_update(history, ahl, channels, scale, quant_max, force_pow2):
    _require_cuda(history, "amax_history")
    assert history.dtype == torch.float32 and scale.dtype == torch.float32 and history.is_contiguous()
    assert history.numel() == ahl * channels and scale.numel() == channels and scale.device == history.device
    with torch.cuda.device(history.device):
        _check(lib().qt_scale_update(history.data_ptr(), ahl, channels, scale.data_ptr(), float(quant_max),
                                     int(bool(force_pow2)), _stream(history)))


GEMM_BF16, GEMM_E4M3, GEMM_E5M2, GEMM_E4M3_E5M2, GEMM_E5M2_E4M3 = range(5)
ACTIVATIONS = {None: 0, "none": 0, "relu": 1, "gelu": 2, "silu": 3}


def _as_batched(t, name):
    """[..., rows, K] tensor -> (batch, rows, K, ld, batch_stride) with a unit-stride K axis, no copy if possible."""
    if t.dim() < 2:
        raise ValueError(f"{name} must have at least 2 dimensions")
    if t.dim() == 2:
        t3 = t.unsqueeze(0)
    else:
        t3 = t.reshape(-1, t.shape[-2], t.shape[-1])  # a view whenever the leading dims are collapsible
    if t3.stride(-1) != 1 or (t3.shape[1] > 1 and t3.stride(1) < t3.shape[2]):
        t3 = t3.contiguous()
    return t3


def gemm_nt(a, b, alpha=1.0, bias=None, activation=None, residual=None, operand_type=GEMM_BF16, out=None):
    """out[..., m, n] = epilogue(alpha * sum_k a[..., m, k] * b[..., n, k]) on the tcgen05 kernel.
    a, b: bf16 (GEMM_BF16) or uint8 fp8 codes; bias bf16 [n]; residual bf16 broadcastable to out's shape."""
    _require_cuda(a, "a")
    a3, b3 = _as_batched(a, "a"), _as_batched(b, "b")
    if b3.shape[0] != a3.shape[0]:
        if b3.shape[0] == 1:
            b3 = b3.expand(a3.shape[0], -1, -1)
        else:
            raise ValueError(f"batch mismatch: {tuple(a.shape)} x {tuple(b.shape)}")
    batch, M, K = a3.shape
    N = b3.shape[1]
    if b3.shape[2] != K:
        raise ValueError(f"inner dimensions differ: {tuple(a.shape)} x {tuple(b.shape)}^T")
    want = torch.uint8 if operand_type != GEMM_BF16 else torch.bfloat16
    if a3.dtype != want or b3.dtype != want:
        raise TypeError(f"operand_type {operand_type} takes {want} operands, got {a3.dtype} and {b3.dtype}")
    out_shape = (*a.shape[:-1], N) if a.dim() > 2 or b.dim() <= 2 else (*b.shape[:-2], M, N)
    if out is None:
        out = torch.empty((batch, M, N), dtype=torch.bfloat16, device=a.device)
    o3 = out.view(batch, M, N)
    r3 = None
    if residual is not None:
        r3 = residual.expand(out_shape).reshape(batch, M, N)
        if r3.stride(-1) != 1:
            r3 = r3.contiguous()
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == N and bias.is_contiguous()
    sa = a3.stride(0) if batch > 1 else 0
    sb = b3.stride(0) if batch > 1 else 0
    with torch.cuda.device(a.device):
        _check(lib().qt_gemm_nt(
            a3.data_ptr(), b3.data_ptr(), o3.data_ptr(), operand_type, batch, M, N, K,
            a3.stride(1), b3.stride(1), o3.stride(1), sa, sb, o3.stride(0) if batch > 1 else 0,
            float(alpha), bias.data_ptr() if bias is not None else None, ACTIVATIONS[activation],
            r3.data_ptr() if r3 is not None else None, r3.stride(1) if r3 is not None else 0,
            (r3.stride(0) if batch > 1 else 0) if r3 is not None else 0, _stream(a)))
    return out.view(out_shape)
