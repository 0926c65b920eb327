"""Quantizable BERT / RoBERTa blocks (reference: modules/quantizable/modeling_bert.py:32-222),
written against the transformers version installed in this image (5.x block signatures)."""
import torch
from torch import nn
from transformers.models.bert import modeling_bert as hf

from ... import fused
from ._common import attention_ops, hooked_attention, rebrand
from .functional_modules import AddFunctional

__all__ = ["BertLayer", "BertSelfAttention", "BertSelfOutput", "BertOutput"]


class BertSelfAttention(hf.BertSelfAttention):
    """query/key/value projections + attention with hookable qk_matmul, attn_scaling, softmax, av_matmul."""

    def __init__(self, config, *args, **kwargs):
        super().__init__(config, *args, **kwargs)
        for name, mod in attention_ops().items():
            self.add_module(name, mod)

    def forward(self, hidden_states, attention_mask=None, past_key_values=None, **kwargs):
        lead = hidden_states.shape[:-1]
        split = (*lead, -1, self.attention_head_size)
        q = self.query(hidden_states).view(*split).transpose(1, 2)
        k = self.key(hidden_states).view(*split).transpose(1, 2)
        v = self.value(hidden_states).view(*split).transpose(1, 2)
        if past_key_values is not None:
            cache = getattr(past_key_values, "self_attention_cache", past_key_values)
            k, v = cache.update(k, v, self.layer_idx)
        scaling = getattr(self, "scaling", self.attention_head_size ** -0.5)
        ctx, probs = hooked_attention(self, q, k, v, attention_mask, scaling, self.dropout.p)
        return ctx.reshape(*lead, -1).contiguous(), probs

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, attention_ops())


class _ResidualNormOutput(nn.Module):
    """dense -> dropout -> LayerNorm(residual(dense_out, input)) with a hookable residual add."""

    def forward(self, hidden_states: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        hidden_states = self.dropout(self.dense(hidden_states))
        return self.LayerNorm(self.residual(hidden_states, input_tensor))

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, {"residual": AddFunctional()})


class BertSelfOutput(_ResidualNormOutput):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.residual = AddFunctional()


class BertOutput(_ResidualNormOutput):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.residual = AddFunctional()


class BertLayer(hf.BertLayer):
    """The HF encoder layer, unchanged in structure and parameter names; in inference with observer-free
    fake-quantizers its forward runs as ~14 fused launches (fused.py) instead of module by module.
    (Not in the reference's mapping: the reference has no fused execution; module names and state-dict keys are
    those of the HF layer, so checkpoints and hooks are unaffected.)"""

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                past_key_values=None, **kwargs):
        if encoder_hidden_states is None and past_key_values is None and not self.is_decoder:
            out = fused.bert_layer_forward(self, hidden_states, attention_mask)
            if out is not None:
                return out
        return super().forward(hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask,
                               past_key_values, **kwargs)

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, {})
