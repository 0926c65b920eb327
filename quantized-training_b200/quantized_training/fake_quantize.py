"""Fake-quantize module backed by the sm_100a kernels.

Interface mirror of the reference's ``fake_quantize.py``: ``get_quantization_map`` (:31-95),
``FusedAmaxObsFakeQuantFunction`` (:197-252) and ``FusedAmaxObsFakeQuantize`` (:255-435) --
same constructor, buffers, state-dict keys, lazily shaped ``scale`` / ``amax_history``,
straight-through backward.  What differs is the engine: one fused CUDA pass
(``qt_scale_update`` + ``qt_fq_forward``) instead of ~15 ATen launches, a Python chunk loop
over a 65 536-entry table and two host synchronisations per call.
"""
import logging
from typing import Optional

import torch
from torch.ao.quantization import FakeQuantizeBase

from . import _C
from .quantizer.quantizer import QScheme

__all__ = ["FusedAmaxObsFakeQuantize", "get_quantization_map"]

logger = logging.getLogger(__name__)


def get_quantization_map(dtype, device=None):
    """bf16[65536]: value every bf16 bit pattern rounds to.  The kernels do not use this table
    (they round bitwise); it is produced from the same rounding code for API parity."""
    if dtype is None:
        t = torch.arange(2 ** 16, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    else:
        t = _C.table_host(_C.format_from_string(dtype))
    return t.to(device) if device is not None else t


def _channel_view(shape, ch_axis):
    """Contiguous tensor of `shape` seen as [outer, channels, inner] around ch_axis."""
    if ch_axis is None or isinstance(ch_axis, (tuple, list)):
        raise TypeError("per_channel_symmetric needs an integer ch_axis (qspec key 'ax')")
    nd = len(shape)
    ax = ch_axis + nd if ch_axis < 0 else ch_axis
    if not 0 <= ax < nd:
        raise IndexError(f"ch_axis {ch_axis} out of range for a {nd}-d tensor")
    outer = inner = 1
    for d in shape[:ax]:
        outer *= d
    for d in shape[ax + 1:]:
        inner *= d
    keep = tuple(shape[i] if i == ax else 1 for i in range(nd))
    return outer, shape[ax], inner, keep


def _run(mod, x, emit_codes=False):
    """Observer step + one fused pass.  emit_codes=False: fake-quantized tensor (x's dtype).
    emit_codes=True: uint8 fp8 codes of round_fmt(x / s) (per tensor, e4m3 / e5m2 formats)."""
    observe, quantize = mod._flags()
    if emit_codes:
        quantize = True
    if not (observe or quantize):
        return x
    if not x.is_cuda:
        raise RuntimeError(
            f"FusedAmaxObsFakeQuantize got a tensor on {x.device}: the B200 build runs on CUDA only "
            "(no CPU fallback)")
    if x.dtype == torch.float16:
        if emit_codes:
            raise TypeError("fp8 codes are produced from bfloat16 / float32 tensors")
        return _run_half(mod, x, observe, quantize)
    xd = x.detach()
    perm = None
    if mod.preserve_strides and not xd.is_contiguous() and not mod.is_per_channel and xd.dim() > 1:
        # A dense permutation (e.g. key.transpose(-1, -2)): the op is elementwise, so quantize the storage
        # order and hand back the same view -- no transpose copy, and the GEMM keeps a unit-stride K axis.
        order = sorted(range(xd.dim()), key=lambda i: -xd.stride(i))
        xp = xd.permute(order)
        if xp.is_contiguous():
            perm = [order.index(i) for i in range(xd.dim())]
            xd = xp
    xc = xd.contiguous()
    amax_slot = None
    outer, channels, inner = 1, 1, xc.numel()
    if observe:
        if xc.numel() == 0:
            raise RuntimeError("amax(): cannot observe an empty tensor")
        if mod.is_per_channel:
            outer, channels, inner, stat_shape = _channel_view(tuple(xc.shape), mod.ch_axis)
        else:
            stat_shape = ()
        if mod.amax_history.numel() == 0:  # first observed call: shapes become known
            mod.amax_history.resize_((mod.amax_history_len,) + stat_shape).fill_(0.0)
            mod.scale.resize_(stat_shape).fill_(1.0)
        _C.scale_update(mod.amax_history, mod.amax_history_len, channels, mod.scale, mod.quant_max,
                        mod.force_scale_power_of_two)
        # the kernel writes `scale` through a raw pointer, which torch's version counter does not see: consumers that
        # cache something derived from the scale (quantized weights) key on this epoch instead
        mod._scale_epoch += 1
        amax_slot = mod.amax_history
    if not quantize:
        _C.amax(xc, outer, channels, inner, amax_slot)
        return x
    scale = mod.scale
    if scale.numel() == 1:
        # per tensor, or a bare spec whose scale buffer is the constant 1.0 (the kernel then
        # takes its exact unit-scale path: the branch is made on the device, no read-back)
        outer, channels, inner = 1, 1, xc.numel()
    elif not observe:
        # frozen per-channel scale: recover the layout from the scale's keepdim shape
        axes = [i for i, d in enumerate(scale.shape) if d != 1]
        if len(axes) != 1 or scale.dim() != xc.dim() or scale.shape[axes[0]] != xc.shape[axes[0]]:
            raise RuntimeError(f"scale of shape {tuple(scale.shape)} does not broadcast over a single "
                               f"axis of input {tuple(xc.shape)}")
        outer, channels, inner, _ = _channel_view(tuple(xc.shape), axes[0])
    if emit_codes:
        if channels != 1:
            raise NotImplementedError("fp8 codes are produced per tensor (bare or per_tensor_symmetric specs)")
        y = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device)
        _C.quantize_codes(xc, y, mod._fmt, scale.reshape(1), amax_slot, mod.lut)
    else:
        y = torch.empty_like(xc)
        _C.fq_forward(xc, y, outer, channels, inner, mod._fmt, scale, amax_slot, mod.lut)
    return y if perm is None else y.permute(perm)


def _run_half(mod, x, observe, quantize):
    """float16 tensors.  The reference's arithmetic is dtype-generic (fake_quantize.py:217-246): amax of the fp16
    tensor, `scale.to(float16)`, an fp16 division, the table lookup through an exact fp32 widening with round-to-odd
    truncation (decomposed.py:150-153), the table value narrowed to fp16, an fp16 multiply.  The observer and the
    lookup run on the fp32 kernels (widening fp16 is exact); the two fp16 roundings of the scaled path are torch's own
    division and multiplication, which is where the reference gets them too.  A compatibility path: five passes
    instead of one (the models of the BASELINE configs are bf16)."""
    xc = x.detach().contiguous()
    x32 = xc.float()
    if observe:
        if xc.numel() == 0:
            raise RuntimeError("amax(): cannot observe an empty tensor")
        outer, channels, inner, stat_shape = (_channel_view(tuple(xc.shape), mod.ch_axis) if mod.is_per_channel
                                              else (1, 1, xc.numel(), ()))
        if mod.amax_history.numel() == 0:
            mod.amax_history.resize_((mod.amax_history_len,) + stat_shape).fill_(0.0)
            mod.scale.resize_(stat_shape).fill_(1.0)
        _C.scale_update(mod.amax_history, mod.amax_history_len, channels, mod.scale, mod.quant_max,
                        mod.force_scale_power_of_two)
        mod._scale_epoch += 1
        _C.amax(x32, outer, channels, inner, mod.amax_history)
    if not quantize:
        return x
    s16 = mod.scale.to(torch.float16)
    if s16.numel() > 1 and s16.dim() != xc.dim():
        raise RuntimeError(f"scale of shape {tuple(s16.shape)} does not broadcast over input {tuple(xc.shape)}")
    t32 = (xc / s16).float().contiguous()
    q32 = torch.empty_like(t32)
    _C.fq_forward(t32, q32, 1, 1, t32.numel(), mod._fmt, None, None, mod.lut)
    return q32.to(torch.float16) * s16


def _block_view(shape, axes, block_size):
    """(d0, n1, d1, n2, d2), tiled-second-axis flag and the block-grid shape for tiling `axes` of a contiguous
    tensor with block_size (mx_utils.py:62-121).  One or two tiled axes."""
    nd = len(shape)
    axes = [axes] if isinstance(axes, int) else list(axes)
    axes = sorted({a + nd if a < 0 else a for a in axes})
    if not axes or any(not 0 <= a < nd for a in axes):
        raise IndexError(f"block axes {axes} out of range for a {nd}-d tensor")
    if len(axes) > 2:
        raise NotImplementedError("block-scaled qschemes tile one or two axes")

    def prod(v):
        r = 1
        for d in v:
            r *= d
        return r

    grid = [-(-d // block_size) if i in axes else d for i, d in enumerate(shape)]
    a1 = axes[0]
    if len(axes) == 1:
        return (prod(shape[:a1]), shape[a1], prod(shape[a1 + 1:]), 1, 1), False, grid
    a2 = axes[1]
    return (prod(shape[:a1]), shape[a1], prod(shape[a1 + 1:a2]), shape[a2], prod(shape[a2 + 1:])), True, grid


def _run_block(mod, x):
    """microscaling / group_wise_affine: statistic, parameters and quantize-dequantize of every block in one
    call; the parameters land in the module's `scale` (and `zero_point`) buffers, shaped like the block grid."""
    if not mod._flags()[1]:  # only fake_quant_enabled gates these schemes (fake_quantize.py:116, 150)
        return x
    if not x.is_cuda:
        raise RuntimeError(
            f"FusedAmaxObsFakeQuantize got a tensor on {x.device}: the B200 build runs on CUDA only "
            "(no CPU fallback)")
    xc = x.detach().contiguous()
    bs = mod.block_size
    if not isinstance(bs, int) or bs <= 0:
        raise AssertionError("block_size must be a positive integer")  # decomposed.py:381 `assert block_size > 0`
    dims, axis2, grid = _block_view(tuple(xc.shape), mod.ch_axis, bs)
    affine = mod.qscheme == QScheme.GROUP_WISE_AFFINE
    mod.scale.resize_(grid)
    if affine:
        mod.zero_point.resize_(grid)
    y = torch.empty_like(xc)
    if xc.numel() == 0:
        return y
    pow2 = mod.force_scale_power_of_two and not affine
    _C.fq_block(xc, y, dims, bs, axis2, _C.BLOCK_AFFINE if affine else _C.BLOCK_MX,
                mod.quant_min if mod.quant_min is not None else 0.0, mod.quant_max, None if affine else mod._fmt,
                mod.scale, mod.zero_point if affine else None, None if affine else mod.lut, mod._scale_fmt,
                pow2, mod._pow2_table(xc) if pow2 else None)
    return y


class BlockScaledFakeQuantFunction(torch.autograd.Function):
    """MXFakeQuantFunction / GroupWiseAffineFakeQuantFunction (fake_quantize.py:98-194); STE backward."""

    @staticmethod
    def forward(ctx, x, mod):
        return _run_block(mod, x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


class FusedAmaxObsFakeQuantFunction(torch.autograd.Function):
    """Delayed-scaling observer + quantize-dequantize, one fused pass; STE backward."""

    @staticmethod
    def forward(ctx, x, mod):
        return _run(mod, x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


class FusedAmaxObsFakeQuantize(FakeQuantizeBase):
    r"""Simulated quantization with amax-history ("delayed") scaling.

    ``dtype`` is a format string (``int8``, ``e4m3``, ``fp6_e3m2``, ``posit8_1`` ...).  With
    ``qscheme=None`` values are rounded to the format as they are (scale 1).  With
    ``per_tensor_symmetric`` / ``per_channel_symmetric`` the scale used by call *k* is
    ``max(amax of the previous amax_history_len calls) / quant_max``.
    """

    amax_history: torch.Tensor
    scale: torch.Tensor
    zero_point: torch.Tensor

    def __init__(
        self,
        dtype: str,
        qscheme: Optional[QScheme] = None,
        quant_min: Optional[float] = None,
        quant_max: Optional[float] = None,
        amax_history_len: int = None,
        ch_axis: Optional[int] = None,
        block_size: Optional[int] = None,
        record_histogram: bool = False,
        scale_dtype: Optional[str] = None,
        force_scale_power_of_two: bool = False,
        outlier_threshold: Optional[float] = None,
        **kwargs,
    ) -> None:
        super().__init__()
        if isinstance(qscheme, str):
            qscheme = QScheme(qscheme)
        self.is_block_scaled = qscheme in (QScheme.MICROSCALING, QScheme.GROUP_WISE_AFFINE)
        if outlier_threshold is not None:
            raise NotImplementedError("outlier_threshold belongs to the PT2E flow, outside the B200 hot path")
        if self.is_block_scaled:
            if quant_max is None or block_size is None or ch_axis is None:
                raise ValueError("quant_max, block_size and ch_axis are required for block-scaled qschemes")
            if qscheme == QScheme.GROUP_WISE_AFFINE and quant_min is None:
                raise ValueError("quant_min is required for group_wise_affine")
        elif qscheme is not None and (quant_max is None or amax_history_len is None):
            raise ValueError("quant_max and amax_history_len are required when a qscheme is given")
        self.dtype = dtype
        self.qscheme = qscheme
        self.quant_min = quant_min
        self.quant_max = quant_max
        self.amax_history_len = amax_history_len
        self.ch_axis = ch_axis
        self.block_size = block_size
        self.scale_dtype = scale_dtype
        self.force_scale_power_of_two = force_scale_power_of_two
        self.outlier_threshold = outlier_threshold
        self.record_histogram = record_histogram
        # False (default): outputs are contiguous, as in the reference.  True: a dense permuted input (k^T) comes
        # back as the same view; set by `prepare` on the fake-quantizers that feed our own matmul.
        self.preserve_strides = False
        self._fmt = _C.format_from_string(dtype)  # ValueError("Unsupported dtype: ...")
        self._scale_fmt = _C.format_from_string(scale_dtype) if scale_dtype is not None else None
        self._qmap = None
        self._scale_qmap = None
        self._pow2_tables = {}
        device = kwargs.get("device", None)
        f32 = dict(device=device, dtype=torch.float)
        self.register_buffer("amax_history", torch.tensor([], **f32))
        self.register_buffer("scale", torch.tensor([1.0], **f32))
        self.register_buffer("zero_point", torch.tensor([1.0], **f32))
        self.register_buffer("histogram", torch.zeros(254, **f32), persistent=False)
        # 8 KB of per-binade rounding constants for the kernels' fast path (None: int / native dtypes).
        # The counterpart of the reference's 128 KB `qmap` buffer; non-persistent like it.
        self.register_buffer("lut", _C.lut_host(self._fmt), persistent=False)
        self.is_per_channel = qscheme == QScheme.PER_CHANNEL_SYMMETRIC
        self._flag_versions = None
        self._scale_epoch = 0   # bumped by every observer update and every external write to `scale` (state_key())
        self.enable_observer(qscheme is not None)
        if device is not None:
            self.to(device)

    # -- enable flags: uint8 buffers (state-dict / DDP compatible) mirrored on the host so that the
    #    hot path never reads device memory back.  The mirror is refreshed whenever a buffer's
    #    version counter moved (in-place writes, load_state_dict), i.e. never in steady state.
    def _flags(self):
        versions = (self.observer_enabled._version, self.fake_quant_enabled._version,
                    id(self.observer_enabled), id(self.fake_quant_enabled))
        if versions != self._flag_versions:
            self._observe = bool(self.observer_enabled[0].item() == 1)
            self._quantize = bool(self.fake_quant_enabled[0].item() == 1)
            self._flag_versions = versions
        return self._observe, self._quantize

    def _apply(self, fn, *args, **kwargs):
        """model.bfloat16() / .half() / .to(dtype) cast every floating-point buffer.  The observer state (`scale`,
        `amax_history`, `zero_point`, `histogram`) is fp32 by definition -- the kernels read it through raw pointers --
        so a dtype cast is undone here (device moves are kept); the reference tolerates such casts because its buffers
        only ever meet torch ops."""
        keep = {n: getattr(self, n).detach().clone() for n in ("scale", "amax_history", "zero_point", "histogram")
                if getattr(self, n, None) is not None}
        out = super()._apply(fn, *args, **kwargs)
        for n, old in keep.items():
            cur = getattr(self, n)
            if cur.dtype != torch.float32:
                self._buffers[n] = old.to(device=cur.device, dtype=torch.float32)
        self._flag_versions = None
        return out

    def state_key(self):
        """Hashable token that changes whenever the function this module computes may have changed: the observer
        updated the scale (epoch), the scale buffer was written or replaced by torch (version / storage), or the
        enable flags moved.  Used to key caches of quantized weights (qat.Linear, fused.py)."""
        observe, quantize = self._flags()
        return (self._scale_epoch, self.scale.data_ptr(), self.scale._version, observe, quantize)

    @property
    def qmap(self):
        if self._qmap is None or self._qmap.device != self.scale.device:
            self._qmap = get_quantization_map(self.dtype, self.scale.device)
        return self._qmap

    @property
    def scale_qmap(self):
        if self.scale_dtype is None:
            return None
        if self._scale_qmap is None or self._scale_qmap.device != self.scale.device:
            self._scale_qmap = get_quantization_map(self.scale_dtype, self.scale.device)
        return self._scale_qmap

    def _pow2_table(self, x):
        """Device copy of the floor(log2()) thresholds for x's dtype (force_scale_power_of_two, block-scaled)."""
        key = (x.dtype, x.device)
        t = self._pow2_tables.get(key)
        if t is None:
            t = _C.pow2_table_host(_C.QT_F32 if x.dtype == torch.float32 else _C.QT_BF16).to(x.device)
            self._pow2_tables[key] = t
        return t

    @torch.jit.export
    def calculate_qparams(self):
        if self.qscheme == QScheme.GROUP_WISE_AFFINE:
            return self.scale, self.zero_point
        return self.scale

    @property
    def fp8_kind(self):
        """"e4m3" / "e5m2" when the format has a one-byte OCP encoding the FP8 tensor cores read, else None."""
        f = self._fmt
        if f.kind == 2 and not f.is_unsigned and (f.ebits, f.mbits) in ((4, 3), (5, 2)):
            return "e4m3" if f.ebits == 4 else "e5m2"
        return None

    def quantize_to_codes(self, X: torch.Tensor) -> torch.Tensor:
        """uint8 tensor of fp8 codes of round(X / scale): same observer side effects as forward(), 3 bytes of
        traffic per bf16 element, and the operand format of ops.linear_fp8.  No gradient flows through it."""
        if self.fp8_kind is None:
            raise ValueError(f"dtype {self.dtype} has no fp8 code form")
        if self.scale.device != X.device:
            self.to(X.device)
        return _run(self, X, emit_codes=True)

    @torch.jit.export
    def extra_repr(self):
        return (f"fake_quant_enabled={self.fake_quant_enabled}, observer_enabled={self.observer_enabled}, "
                f"dtype={self.dtype}, amax_history_len={self.amax_history_len}, quant_max={self.quant_max}, "
                f"qscheme={self.qscheme}, ch_axis={self.ch_axis}, block_size={self.block_size}, "
                f"force_scale_power_of_two={self.force_scale_power_of_two}, scale={self.scale}")

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if self.scale.device != X.device:
            self.to(X.device)
        if self.record_histogram:
            mag = X.detach().float().abs()
            self.histogram += torch.histc(torch.log2(mag).floor(), bins=254, min=-126, max=127)
        if not (X.requires_grad and torch.is_grad_enabled()):
            # nothing to differentiate (inference, frozen activations, gradient hooks in backward): skip the
            # autograd.Function round trip -- ~5 us of host time per call, ~600 calls per eager fine-tune step
            return _run_block(self, X) if self.is_block_scaled else _run(self, X)
        if self.is_block_scaled:
            return BlockScaledFakeQuantFunction.apply(X, self)
        return FusedAmaxObsFakeQuantFunction.apply(X, self)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict,
                              missing_keys, unexpected_keys, error_msgs):
        # `scale` / `amax_history` are shaped lazily by the first observed call; adopt checkpoint shapes
        for name in ("scale", "amax_history", "zero_point"):
            key = prefix + name
            if key in state_dict:
                getattr(self, name).resize_(state_dict[key].shape)
            elif strict:
                missing_keys.append(key)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict,
                                      missing_keys, unexpected_keys, error_msgs)
        self._flag_versions = None
        self._scale_epoch += 1
