"""QAT LoRA linear: fake-quantize A, B and the merged weight W + s*B@A on every forward, then ONE
GEMM with the merged weight (reference: modules/qat/lora.py:34-55).  Note what that implies, and is
kept: LoRA dropout is not applied on this path and the ``lora_A`` / ``lora_B`` sub-Linears are never
called (their ``.weight`` is read directly), so hooks on them do not fire."""
import torch
import torch.nn.functional as F

from ... import ops
from ..lora import LoraLinear as _FloatLora

try:  # the real peft layer, when the package is present
    from peft.tuners.lora import Linear as _PeftLora
except Exception:  # pragma: no cover - peft is not part of this image
    _PeftLora = None

__all__ = ["Linear"]


def _t(w, fan_in_fan_out):
    return w.T if fan_in_fan_out else w


class Linear(_FloatLora):
    _FLOAT_MODULE = _FloatLora
    _FLOAT_MODULES = tuple(c for c in (_FloatLora, _PeftLora) if c is not None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        in_dtype = x.dtype
        if self.disable_adapters or self.merged:
            return ops.linear(x, _t(self.weight, self.fan_in_fan_out), self.bias).to(in_dtype)
        merged = self.weight.data.clone()
        for name in self.active_adapters:
            if name in self.lora_A:
                a = self.weight_fake_quant(self.lora_A[name].weight)
                b = self.weight_fake_quant(self.lora_B[name].weight)
                merged = merged + _t(b @ a, self.fan_in_fan_out) * self.scaling[name]
        merged = self.weight_fake_quant(merged)
        return ops.linear(x, _t(merged, self.fan_in_fan_out), self.bias).to(in_dtype)

    @classmethod
    def from_float(cls, mod):
        assert isinstance(mod, cls._FLOAT_MODULES), f"qat.LoraLinear.from_float got {type(mod).__name__}"
        assert getattr(mod, "qconfig", None), "Input float module must have a valid qconfig"
        if getattr(mod, "merged", False) and hasattr(mod, "unmerge"):
            mod.unmerge()
        name = mod.active_adapter[0] if isinstance(mod.active_adapter, (list, tuple)) else mod.active_adapter
        qat = cls(mod.in_features, mod.out_features, bias=mod.bias is not None, r=mod.r[name],
                  lora_alpha=mod.lora_alpha[name], adapter_name=name,
                  fan_in_fan_out=getattr(mod, "fan_in_fan_out", False), device="meta")
        qat.qconfig = mod.qconfig
        qat.weight_fake_quant = mod.qconfig.weight()
        base = mod.base_layer if hasattr(mod, "base_layer") else mod  # peft >= 0.6 nests the float Linear
        qat.weight, qat.bias = base.weight, base.bias
        for attr in ("r", "lora_alpha", "scaling", "lora_dropout", "lora_A", "lora_B"):
            setattr(qat, attr, getattr(mod, attr))
        qat.active_adapter = [name]
        return qat
