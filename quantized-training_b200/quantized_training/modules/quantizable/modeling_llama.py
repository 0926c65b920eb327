"""Quantizable Llama decoder layer (reference: modules/quantizable/modeling_llama.py:95-356; its
mapping entry is commented out at HEAD, quantization_mappings.py:33 -- enabled here so that the
five fusion levels of the README Llama table have their hook points).  The softmax runs in the
tensor dtype, as in the reference's block (no fp32 upcast, modeling_llama.py:244)."""
from transformers.models.llama import modeling_llama as hf

from ... import fused
from ._common import attention_ops, hooked_attention, rebrand
from .functional_modules import AddFunctional

__all__ = ["LlamaAttention", "LlamaDecoderLayer"]


class LlamaAttention(hf.LlamaAttention):
    def __init__(self, config, layer_idx):
        super().__init__(config, layer_idx)
        for name, mod in attention_ops().items():
            self.add_module(name, mod)

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        lead = hidden_states.shape[:-1]
        split = (*lead, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(split).transpose(1, 2)
        k = self.k_proj(hidden_states).view(split).transpose(1, 2)
        v = self.v_proj(hidden_states).view(split).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = hf.apply_rotary_pos_emb(q, k, cos, sin)
        if past_key_values is not None:
            k, v = past_key_values.update(k, v, self.layer_idx)
        ctx, probs = hooked_attention(self, q, k, v, attention_mask, self.scaling, self.attention_dropout,
                                      self.num_key_value_groups)
        return self.o_proj(ctx.reshape(*lead, -1).contiguous()), probs

    @classmethod
    def from_observed(cls, other):
        return rebrand(other, cls, attention_ops())


class LlamaDecoderLayer(hf.LlamaDecoderLayer):
    def __init__(self, config, layer_idx):
        super().__init__(config, layer_idx)
        self.self_attn = LlamaAttention(config, layer_idx)
        self.self_attn_residual = AddFunctional()
        self.mlp_residual = AddFunctional()

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_values=None,
                use_cache=False, position_embeddings=None, **kwargs):
        # inference with observer-free fake-quantizers: the whole layer as ~13 fused launches (fused.py)
        out = fused.llama_layer_forward(self, hidden_states, attention_mask, position_embeddings, past_key_values)
        if out is not None:
            return out
        attn_out, _ = self.self_attn(
            hidden_states=self.input_layernorm(hidden_states), attention_mask=attention_mask,
            position_ids=position_ids, past_key_values=past_key_values, use_cache=use_cache,
            position_embeddings=position_embeddings, **kwargs)
        hidden_states = self.self_attn_residual(hidden_states, attn_out)
        mlp_out = self.mlp(self.post_attention_layernorm(hidden_states))
        return self.mlp_residual(hidden_states, mlp_out)

    @classmethod
    def from_observed(cls, other):
        if not hasattr(other, "config"):
            other.config = other.self_attn.config
        new = rebrand(other, cls, {"self_attn_residual": AddFunctional(), "mlp_residual": AddFunctional()})
        if not isinstance(new.self_attn, LlamaAttention):
            new.self_attn = LlamaAttention.from_observed(new.self_attn)
        return new
