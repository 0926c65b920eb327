"""QAT convolutions: the filter passes through ``qconfig.weight()`` on every forward (reference:
modules/qat/conv.py:12-100 `_ConvNd`, :102-260 Conv1d/2d/3d).  The fake-quant is this repo's kernel; the convolution
itself is the library call the reference makes too (`F.convNd` -> cuDNN): none of the reference's headline models has a
convolution on the measured path, so there is no hand-written implicit-GEMM here."""
import torch
import torch.nn as nn
from torch.nn.utils.parametrize import (
    is_parametrized,
    transfer_parametrizations_and_params,
    type_before_parametrizations,
)

__all__ = ["Conv1d", "Conv2d", "Conv3d"]

_CTOR_ARGS = ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups", "padding_mode")


class _QatConv:
    """Mixin placed in front of an ``nn.ConvNd``: adds ``weight_fake_quant`` and the float <-> QAT conversions."""

    _FLOAT_MODULE = None

    def _attach(self, qconfig, device, dtype):
        assert qconfig, "qconfig must be provided for QAT module"
        self.qconfig = qconfig
        self.weight_fake_quant = qconfig.weight(factory_kwargs={"device": device, "dtype": dtype})

    def forward(self, input):
        return self._conv_forward(input, self.weight_fake_quant(self.weight), self.bias)

    @classmethod
    def from_float(cls, mod):
        """Wrap a float convolution; Parameters are shared, not copied (reference conv.py:47-72)."""
        assert type_before_parametrizations(mod) == cls._FLOAT_MODULE, (
            f"qat.{cls.__name__}.from_float only works for {cls._FLOAT_MODULE.__name__}")
        assert getattr(mod, "qconfig", None), "Input float module must have a valid qconfig"
        kw = {k: getattr(mod, k) for k in _CTOR_ARGS}
        qat = cls(**kw, bias=mod.bias is not None, qconfig=mod.qconfig, device="meta")
        qat.weight_fake_quant = mod.qconfig.weight()
        for name in ("weight", "bias"):
            if is_parametrized(mod, name):
                transfer_parametrizations_and_params(mod, qat, name)
            else:
                setattr(qat, name, getattr(mod, name))
        return qat

    def to_float(self):
        """Back to the plain convolution, dropping the quantizer (reference conv.py:74-100)."""
        kw = {k: getattr(self, k) for k in _CTOR_ARGS}
        conv = self._FLOAT_MODULE(**kw, bias=self.bias is not None)
        conv.weight = nn.Parameter(self.weight.detach())
        if self.bias is not None:
            conv.bias = nn.Parameter(self.bias.detach())
        conv.train(self.training)
        return conv


def _make(float_cls):
    class _Conv(_QatConv, float_cls):
        _FLOAT_MODULE = float_cls

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     bias=True, padding_mode="zeros", qconfig=None, device=None, dtype=None) -> None:
            float_cls.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                               dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode,
                               device=device, dtype=dtype)
            self._attach(qconfig, device, dtype)

    _Conv.__name__ = _Conv.__qualname__ = float_cls.__name__
    _Conv.__doc__ = f"``nn.{float_cls.__name__}`` whose weight is fake-quantized before the convolution."
    return _Conv


Conv1d = _make(nn.Conv1d)
Conv2d = _make(nn.Conv2d)
Conv3d = _make(nn.Conv3d)
