"""Quantizable MobileBERT blocks (reference: modules/quantizable/modeling_mobilebert.py:38-206)."""
from torch import nn
from transformers.models.mobilebert import modeling_mobilebert as hf

from ... import fused
from ._common import attention_ops, hooked_attention, rebrand
from .functional_modules import AddFunctional

__all__ = ["MobileBertSelfAttention", "MobileBertSelfOutput", "MobileBertOutput", "FFNOutput", "OutputBottleneck",
           "MobileBertLayer"]


class MobileBertLayer(hf.MobileBertLayer):
    """The HF encoder layer, unchanged in structure and parameter names; in inference with observer-free
    fake-quantizers its forward runs as fused launches (fused.mobilebert_layer_forward) instead of module by module.
    (Not in the reference's mapping: the reference has no fused execution; module names and state-dict keys are those
    of the HF layer, so checkpoints and hooks are unaffected.)"""

    def forward(self, hidden_states, attention_mask=None, **kwargs):
        out = fused.mobilebert_layer_forward(self, hidden_states, attention_mask)
        if out is not None:
            return out
        return super().forward(hidden_states, attention_mask, **kwargs)

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, {})


class MobileBertSelfAttention(hf.MobileBertSelfAttention):
    def __init__(self, config):
        super().__init__(config)
        for name, mod in attention_ops().items():
            self.add_module(name, mod)

    def forward(self, query_tensor, key_tensor, value_tensor, attention_mask=None, **kwargs):
        lead = query_tensor.shape[:-1]
        split = (*lead, -1, self.attention_head_size)
        q = self.query(query_tensor).view(*split).transpose(1, 2)
        k = self.key(key_tensor).view(*split).transpose(1, 2)
        v = self.value(value_tensor).view(*split).transpose(1, 2)
        scaling = getattr(self, "scaling", self.attention_head_size ** -0.5)
        ctx, probs = hooked_attention(self, q, k, v, attention_mask, scaling, self.dropout.p)
        return ctx.reshape(*lead, -1).contiguous(), probs

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, attention_ops())


class _WithResidual:
    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, {"residual": AddFunctional()})


class MobileBertSelfOutput(_WithResidual, hf.MobileBertSelfOutput):
    def __init__(self, config):
        super().__init__(config)
        self.residual = AddFunctional()

    def forward(self, hidden_states, residual_tensor):
        out = self.dense(hidden_states)
        if not self.use_bottleneck:
            out = self.dropout(out)
        return self.LayerNorm(self.residual(out, residual_tensor))


class FFNOutput(_WithResidual, hf.FFNOutput):
    def __init__(self, config):
        super().__init__(config)
        self.residual = AddFunctional()

    def forward(self, hidden_states, residual_tensor):
        return self.LayerNorm(self.residual(self.dense(hidden_states), residual_tensor))


class OutputBottleneck(_WithResidual, hf.OutputBottleneck):
    def __init__(self, config):
        super().__init__(config)
        self.residual = AddFunctional()

    def forward(self, hidden_states, residual_tensor):
        out = self.dropout(self.dense(hidden_states))
        return self.LayerNorm(self.residual(out, residual_tensor))


class MobileBertOutput(_WithResidual, hf.MobileBertOutput):
    def __init__(self, config):
        super().__init__(config)
        self.residual = AddFunctional()
        if self.use_bottleneck:
            self.bottleneck = OutputBottleneck(config)

    def forward(self, intermediate_states, residual_tensor_1, residual_tensor_2):
        out = self.dense(intermediate_states)
        if not self.use_bottleneck:   # a bare `+` in the reference (modeling_mobilebert.py:165-166): never hooked
            return self.LayerNorm(self.dropout(out) + residual_tensor_1)
        out = self.LayerNorm(self.residual(out, residual_tensor_1))
        return self.bottleneck(out, residual_tensor_2)

    @classmethod
    def from_observed(cls, other):
        new = rebrand(other, cls, {"residual": AddFunctional()})
        if new.use_bottleneck and not isinstance(new.bottleneck, OutputBottleneck):
            new.bottleneck = OutputBottleneck.from_observed(new.bottleneck)
        return new
