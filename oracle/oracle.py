"""ctypes front end of the CPU oracle (oracle/qt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package never imports this module (tests/test_no_oracle_in_product.py checks).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqt_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "qt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libqt_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, sz, i32, f32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
        L.qto_qmap.argtypes = [ctypes.c_char_p, vp]
        L.qto_qmap.restype = i32
        L.qto_posit.argtypes = [i32, i32, vp, vp]
        L.qto_posit.restype = i32
        L.qto_vmap_bf16.argtypes = [vp, vp, sz, vp]
        L.qto_vmap_f32.argtypes = [vp, vp, sz, vp]
        L.qto_amax.argtypes = [vp, i32, sz, sz, sz, vp]
        L.qto_scale_update.argtypes = [vp, i32, sz, vp, vp, f32, i32]
        L.qto_fake_quant_bf16.argtypes = [vp, vp, sz, sz, sz, vp, vp]
        L.qto_fake_quant_f32.argtypes = [vp, vp, sz, sz, sz, vp, vp]
        L.qto_mx_fake_quant.argtypes = [vp, vp, i32, i32, vp, vp, i32, f32, i32, vp, vp, vp]
        L.qto_gwa_fake_quant.argtypes = [vp, vp, i32, i32, vp, vp, i32, f32, f32, vp, vp, vp]
        L.qto_mx_fake_quant.restype = None
        L.qto_gwa_fake_quant.restype = None
        L.qto_num_threads.restype = i32
        L.qto_set_num_threads.argtypes = [i32]
        for f in (L.qto_vmap_bf16, L.qto_vmap_f32, L.qto_amax, L.qto_scale_update,
                  L.qto_fake_quant_bf16, L.qto_fake_quant_f32, L.qto_set_num_threads):
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def qmap(dtype):
    """get_quantization_map(dtype) -> uint16[65536] (bf16 bit patterns). ValueError if unsupported."""
    out = np.empty(65536, dtype=np.uint16)
    rc = lib().qto_qmap(dtype.encode() if dtype is not None else None, _p(out))
    if rc != 0:
        raise ValueError(f"Unsupported dtype: {dtype}")
    return out


def posit(nbits, es):
    """(value table uint16[65536], signed posit codes int32[65536])."""
    vals = np.empty(65536, dtype=np.uint16)
    pb = np.empty(65536, dtype=np.int32)
    if lib().qto_posit(nbits, es, _p(vals), _p(pb)) != 0:
        raise ValueError(f"Unsupported dtype: posit{nbits}_{es}")
    return vals, pb


def vmap(x, table):
    """x: uint16 array (bf16 bits) or float32 array."""
    x = np.ascontiguousarray(x)
    y = np.empty_like(x)
    if x.dtype == np.uint16:
        lib().qto_vmap_bf16(_p(x), _p(y), x.size, _p(table))
    elif x.dtype == np.float32:
        lib().qto_vmap_f32(_p(x), _p(y), x.size, _p(table))
    else:
        raise TypeError(x.dtype)
    return y


def channel_view(shape, ch_axis):
    """[outer, C, inner] factorisation of a contiguous tensor for a channel axis (None -> per tensor)."""
    n = int(np.prod(shape)) if len(shape) else 1
    if ch_axis is None:
        return 1, 1, n
    ax = ch_axis + len(shape) if ch_axis < 0 else ch_axis
    outer = int(np.prod(shape[:ax])) if ax > 0 else 1
    inner = int(np.prod(shape[ax + 1:])) if ax + 1 < len(shape) else 1
    return outer, int(shape[ax]), inner


class FakeQuant:
    """FusedAmaxObsFakeQuantize restated on numpy buffers (fake_quantize.py:255-404).

    x arrays are uint16 (bf16 bits) or float32; shape information is passed explicitly.
    """

    def __init__(self, dtype, qscheme=None, quant_max=None, amax_history_len=None, ch_axis=None,
                 force_scale_power_of_two=False):
        self.table = qmap(dtype)
        self.qscheme = qscheme
        self.quant_max = quant_max
        self.ahl = amax_history_len
        self.ch_axis = ch_axis
        self.pow2 = force_scale_power_of_two
        self.observer_enabled = qscheme is not None
        self.fake_quant_enabled = True
        self.scale = np.ones(1, dtype=np.float32)
        self.history = np.zeros(0, dtype=np.float32)

    def __call__(self, x, shape):
        x = np.ascontiguousarray(x)
        is_f32 = x.dtype == np.float32
        per_channel = self.qscheme == "per_channel_symmetric"
        outer, C, inner = channel_view(shape, self.ch_axis if per_channel else None)
        L = lib()
        if self.observer_enabled:
            cur = np.empty(C, dtype=np.float32)
            L.qto_amax(_p(x), int(is_f32), outer, C, inner, _p(cur))
            if self.history.size == 0:
                self.history = np.zeros(self.ahl * C, dtype=np.float32)
                self.scale = np.ones(C, dtype=np.float32)
            L.qto_scale_update(_p(self.history), self.ahl, C, _p(cur), _p(self.scale),
                               float(self.quant_max), int(self.pow2))
        if not self.fake_quant_enabled:
            return x
        y = np.empty_like(x)
        sC = self.scale.size
        if sC == 1:
            outer, C, inner = 1, 1, x.size
        fn = L.qto_fake_quant_f32 if is_f32 else L.qto_fake_quant_bf16
        fn(_p(x), _p(y), outer, C, inner, _p(self.scale), _p(self.table))
        return y


def _block_args(shape, axes, block_size):
    shape = [int(d) for d in shape]
    axes = [axes] if isinstance(axes, int) else list(axes)
    axes = [a + len(shape) if a < 0 else a for a in axes]
    flags = np.array([1 if d in axes else 0 for d in range(len(shape))], dtype=np.int32)
    bshape = [-(-d // block_size) if f else d for d, f in zip(shape, flags)]
    return np.array(shape, dtype=np.uint64), flags, bshape


def mx_fake_quant(x, shape, axes, block_size, quant_max, table, force_pow2=False, scale_table=None):
    """MXFakeQuantFunction.forward (fake_quantize.py:105-129) -> (y, scale[block grid] float32)."""
    x = np.ascontiguousarray(x)
    shp, flags, bshape = _block_args(shape, axes, block_size)
    y = np.empty_like(x)
    scale = np.empty(int(np.prod(bshape)) if bshape else 1, dtype=np.float32)
    lib().qto_mx_fake_quant(_p(x), _p(y), int(x.dtype == np.float32), len(shape), _p(shp), _p(flags),
                            int(block_size), float(quant_max), int(force_pow2),
                            None if scale_table is None else _p(scale_table), _p(table), _p(scale))
    return y, scale.reshape(bshape)


def gwa_fake_quant(x, shape, axes, block_size, quant_min, quant_max, scale_table=None):
    """GroupWiseAffineFakeQuantFunction.forward (fake_quantize.py:138-190) -> (y, scale, zero_point)."""
    x = np.ascontiguousarray(x)
    shp, flags, bshape = _block_args(shape, axes, block_size)
    y = np.empty_like(x)
    nb = int(np.prod(bshape)) if bshape else 1
    scale = np.empty(nb, dtype=np.float32)
    zp = np.empty(nb, dtype=np.float32)
    lib().qto_gwa_fake_quant(_p(x), _p(y), int(x.dtype == np.float32), len(shape), _p(shp), _p(flags),
                             int(block_size), float(quant_min), float(quant_max),
                             None if scale_table is None else _p(scale_table), _p(scale), _p(zp))
    return y, scale.reshape(bshape), zp.reshape(bshape)


def num_threads():
    return lib().qto_num_threads()


def set_num_threads(n):
    lib().qto_set_num_threads(int(n))
