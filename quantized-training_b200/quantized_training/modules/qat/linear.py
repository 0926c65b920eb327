"""QAT Linear: the weight is fake-quantized on every forward (reference: modules/qat/linear.py:15-80)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from torch.nn.utils.parametrize import (
    is_parametrized,
    transfer_parametrizations_and_params,
    type_before_parametrizations,
)

__all__ = ["Linear"]


class Linear(nn.Linear):
    """``nn.Linear`` whose weight passes through ``qconfig.weight()`` before the GEMM.
    Input quantization is not done here: `prepare` installs it as a forward pre-hook."""

    _FLOAT_MODULE = nn.Linear

    def __init__(self, in_features, out_features, bias=True, qconfig=None, device=None, dtype=None) -> None:
        super().__init__(in_features, out_features, bias, device=device, dtype=dtype)
        assert qconfig, "qconfig must be provided for QAT module"
        self.qconfig = qconfig
        self.weight_fake_quant = qconfig.weight(factory_kwargs={"device": device, "dtype": dtype})

    def forward(self, input):
        kind = ops.fp8_route(self, input, self.weight_fake_quant)
        if kind is not None:  # bare e4m3/e5m2 on both sides: operands go to the FP8 tensor cores as codes
            return ops.linear_fp8(input, self.weight, self.bias, self.weight_fake_quant, kind,
                                  codes=self._quantized_weight(codes=True), g_kind=ops.fp8_grad_kind(self))
        return ops.linear(input, self._quantized_weight(), self.bias)

    def _quantized_weight(self, codes=False):
        """weight_fake_quant(weight), or its fp8 codes.  The reference re-quantizes the weight on every forward
        (linear.py:41); that is kept whenever it is observable -- live observer (amax history advances per call)
        or autograd recording through the quantizer.  Otherwise (evaluation with a frozen or bare quantizer) the
        result is a pure function of (weight, scale) and is reused until either is written to: for Llama-2-7B
        this removes 6.6 G elements of re-quantization traffic per forward."""
        fq, w = self.weight_fake_quant, self.weight
        run = (lambda: fq.quantize_to_codes(w)) if codes else (lambda: fq(w))
        flags = getattr(fq, "_flags", None)
        if flags is None or (torch.is_grad_enabled() and w.requires_grad):
            return None if codes else run()
        observe, quantize = flags()
        if observe or not getattr(self, "cache_quantized_weight", True):
            return None if codes else run()
        # fq.state_key(): observer epoch + scale version + flags -- a scale updated by a later calibration pass (the
        # kernel writes it through a raw pointer) invalidates the cache.  The weight side uses torch's version counter,
        # which in-place ops bump but writes through `.data` do NOT: after such writes call invalidate_weight_cache()
        # or set `cache_quantized_weight = False` on the module (the reference re-quantizes on every forward).
        key = (codes, w.data_ptr(), w._version, fq.state_key(), w.dtype, w.device)
        if self.__dict__.get("_wq_key") != key:
            self.__dict__["_wq"] = run().detach()
            self.__dict__["_wq_key"] = key
        return self.__dict__["_wq"]

    def invalidate_weight_cache(self):
        """Drop the cached quantized weight (needed only after writing the weight through `.data`)."""
        self.__dict__.pop("_wq", None)
        self.__dict__.pop("_wq_key", None)

    @classmethod
    def from_float(cls, mod):
        """Wrap a float ``nn.Linear``; Parameters are shared, not copied."""
        assert type_before_parametrizations(mod) == cls._FLOAT_MODULE, (
            f"qat.{cls.__name__}.from_float only works for {cls._FLOAT_MODULE.__name__}")
        assert getattr(mod, "qconfig", None), "Input float module must have a valid qconfig"
        # build on the meta device: the Parameters are replaced by the float module's right below
        qat = cls(mod.in_features, mod.out_features, bias=mod.bias is not None, qconfig=mod.qconfig, device="meta")
        qat.weight_fake_quant = mod.qconfig.weight()
        for name in ("weight", "bias"):
            if is_parametrized(mod, name):
                transfer_parametrizations_and_params(mod, qat, name)
            else:
                setattr(qat, name, getattr(mod, name))
        return qat

    def to_float(self):
        linear = nn.Linear(self.in_features, self.out_features, self.bias is not None)
        linear.weight = nn.Parameter(self.weight.detach())
        if self.bias is not None:
            linear.bias = nn.Parameter(self.bias.detach())
        linear.train(self.training)
        return linear
