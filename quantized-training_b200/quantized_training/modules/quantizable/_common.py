"""Shared pieces of the quantizable transformer blocks.

A quantizable block is the installed Hugging Face block with its tensor ops re-expressed as
*modules* (`qk_matmul`, `attn_scaling`, `softmax`, `av_matmul`, `residual`) so that `prepare`
can hook fake-quantizers onto their inputs; which hooks exist decides the "fusion level"
(reference: modules/quantizable/modeling_*.py).
"""
from collections import OrderedDict

import torch
from torch import nn

from .functional_modules import AddFunctional, MatmulFunctional, MulFunctional


def rebrand(other: nn.Module, cls, extra_modules):
    """`other` re-typed as `cls`: same Parameters, buffers, children and plain attributes (nothing is
    reallocated), fresh hook tables (swap_module re-registers the old hooks), plus `extra_modules`."""
    assert hasattr(other, "config"), "The float module must have 'config'"
    new = cls.__new__(cls)
    new.__dict__.update(other.__dict__)
    for table in ("_parameters", "_buffers", "_modules"):
        new.__dict__[table] = OrderedDict(other.__dict__[table])
    for table, value in other.__dict__.items():
        if table.endswith("_hooks") and isinstance(value, dict):
            new.__dict__[table] = OrderedDict()
    for name, mod in extra_modules.items():
        new.add_module(name, mod)
    return new


def attention_ops():
    return OrderedDict(qk_matmul=MatmulFunctional(), av_matmul=MatmulFunctional(),
                       attn_scaling=MulFunctional(), softmax=nn.Softmax(dim=-1))


def repeat_kv(x, n_rep):
    if n_rep == 1:
        return x
    b, h, s, d = x.shape
    return x[:, :, None, :, :].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


_MASK_CACHE = {}


def additive_mask(mask, dtype):
    """The attention mask as an additive tensor of `dtype`.  HF builds a BOOLEAN mask (True = attend) when the model's
    attention implementation is sdpa (the default) and an additive float mask for "eager"; the quantizable blocks
    always compute attention explicitly (their matmul / softmax modules are the hook points), so a boolean mask
    is converted -- once per mask tensor: every layer of a forward receives the same object."""
    if mask is None or mask.dtype != torch.bool:
        return mask
    key = (mask.data_ptr(), mask._version, tuple(mask.shape), dtype, mask.device)
    hit = _MASK_CACHE.get("last")
    if hit is None or hit[0] != key:
        add = torch.zeros(mask.shape, dtype=dtype, device=mask.device).masked_fill_(~mask, torch.finfo(dtype).min)
        hit = (key, add, mask)   # keep `mask` alive so that its data_ptr cannot be reused by another tensor
        _MASK_CACHE["last"] = hit
    return hit[1]


def hooked_attention(block, query, key, value, attention_mask, scaling, dropout_p=0.0, kv_groups=1):
    """softmax(q k^T * scaling + mask) v through the block's hookable op modules.
    Shapes [B, H, S, D]; returns ([B, S, H, D] contiguous, probabilities)."""
    key, value = repeat_kv(key, kv_groups), repeat_kv(value, kv_groups)
    scores = block.attn_scaling(block.qk_matmul(query, key.transpose(-1, -2)), scaling)
    attention_mask = additive_mask(attention_mask, scores.dtype)
    if attention_mask is not None:
        scores = scores + attention_mask[..., : key.shape[-2]]
    probs = block.softmax(scores).to(query.dtype)
    if dropout_p > 0.0 and block.training:
        probs = nn.functional.dropout(probs, p=dropout_p, training=True)
    out = block.av_matmul(probs, value)
    return out.transpose(1, 2).contiguous(), probs
