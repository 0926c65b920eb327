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
