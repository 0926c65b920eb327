"""A minimal LoRA linear layer (host model piece).

The reference fine-tunes with ``peft`` LoRA adapters (run_glue_no_trainer.py:353-362); ``peft`` is
not part of this image, so this is the float module that `qat.LoraLinear` is created from when
``peft.tuners.lora.Linear`` is unavailable.  Same attribute layout as peft's layer (``lora_A`` /
``lora_B`` ModuleDicts keyed by adapter name, ``scaling``, ``r``, ``lora_alpha``, ``merged``,
``fan_in_fan_out``), so the QAT wrapper treats both alike.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["LoraLinear", "apply_lora"]


class LoraLinear(nn.Linear):
    def __init__(self, in_features, out_features, bias=True, r=8, lora_alpha=8, lora_dropout=0.0,
                 adapter_name="default", fan_in_fan_out=False, device=None, dtype=None):
        super().__init__(in_features, out_features, bias, device=device, dtype=dtype)
        kw = dict(device=device, dtype=dtype)
        self.fan_in_fan_out = fan_in_fan_out
        self.merged = False
        self.disable_adapters = False
        self.active_adapter = [adapter_name]
        self.r = {adapter_name: r}
        self.lora_alpha = {adapter_name: lora_alpha}
        self.scaling = {adapter_name: lora_alpha / r}
        self.lora_dropout = nn.ModuleDict({adapter_name: nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(in_features, r, bias=False, **kw)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, out_features, bias=False, **kw)})
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.weight.requires_grad_(False)
        if self.bias is not None:
            self.bias.requires_grad_(False)

    @property
    def active_adapters(self):
        return self.active_adapter

    @classmethod
    def from_linear(cls, lin, r, lora_alpha, lora_dropout=0.0):
        m = cls(lin.in_features, lin.out_features, lin.bias is not None, r, lora_alpha, lora_dropout,
                device=lin.weight.device, dtype=lin.weight.dtype)
        m.weight, m.bias = lin.weight, lin.bias
        m.weight.requires_grad_(False)
        if m.bias is not None:
            m.bias.requires_grad_(False)
        return m

    def forward(self, x):
        out = F.linear(x, self.weight, self.bias)
        if self.disable_adapters or self.merged:
            return out
        for name in self.active_adapters:
            a, b = self.lora_A[name], self.lora_B[name]
            out = out + b(a(self.lora_dropout[name](x))) * self.scaling[name]
        return out


def apply_lora(model, target_modules, r, lora_alpha, lora_dropout=0.0):
    """Replace every ``nn.Linear`` whose attribute name is in `target_modules` by a LoraLinear
    and freeze everything else (what peft.get_peft_model does for the GLUE recipe)."""
    for p in model.parameters():
        p.requires_grad_(False)
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if type(child) is nn.Linear and name in target_modules:
                setattr(parent, name, LoraLinear.from_linear(child, r, lora_alpha, lora_dropout))
    return model
