"""Fused execution of quantize()-d transformer blocks (inference).

`quantize(model, args)` expresses a quantized block the way the reference does: QAT Linear modules and hookable
matmul / mul / add / softmax modules, with one fake-quantizer per hooked tensor (quantize.py:116-150).  Executed
module by module that is ~60 launches per Llama layer, most of them small bf16 ATen ops between the GEMMs.  This file
executes the SAME computation -- same weights, same fake-quantizers, same rounding points -- as ~13 launches:

    norm+fq -> QKV GEMM -> rope+fq / fq+transpose -> QK^T GEMM -> scale+mask+softmax+fq -> PV GEMM -> fq
            -> O GEMM (+residual) -> norm+fq -> gate|up GEMM -> silu*up+fq -> down GEMM (+residual)

Which fake-quant steps exist is read from the hooks `prepare` installed (the paper's fusion levels = which op groups
are hooked); a step that is absent simply is not applied, and the producing kernel's fp32 epilogue carries the value
(residual adds move into the GEMM epilogue when the `residual` group is not hooked).  A block is executed fused only
when that is unobservable: no autograd recording, every fake-quantizer involved is a bare or frozen per-tensor one
(no live observer, whose amax history must advance per call), and nobody else hooked the submodules.  Otherwise the
block runs module by module on the same kernels (fake_quantize.py, ops.py).
"""
import os
import traceback

import torch
from torch import nn

from . import _C
from .fake_quantize import FusedAmaxObsFakeQuantize

__all__ = ["llama_layer_forward", "bert_layer_forward", "set_enabled", "enabled"]

_ENABLED = True


def set_enabled(flag: bool):
    """Fused block execution on (default) / off (module-by-module, for A/B comparison and debugging)."""
    global _ENABLED
    _ENABLED = bool(flag)


def enabled():
    return _ENABLED


def _why(where):
    """QT_FUSED_DEBUG=1: say why a block fell back to module-by-module execution."""
    if os.environ.get("QT_FUSED_DEBUG"):
        print(f"[qt fused] {where}: module-by-module because of", traceback.format_exc(limit=-2).strip().splitlines()[-3:])


class _NotReady(Exception):
    """A lazily created fake-quantizer does not exist yet (first forward): run module by module."""


class _NotFusable(Exception):
    pass


_IDENTITY = {}


def _identity_fmt():
    if "fmt" not in _IDENTITY:
        _IDENTITY["fmt"] = _C.format_from_string("bfloat16")
    return _IDENTITY["fmt"]


def _only_our_hooks(module, hooked):
    if len(module._forward_hooks) or len(module._backward_hooks) or len(module._forward_pre_hooks) != (1 if hooked else 0):
        raise _NotFusable


def point(module, arg="0"):
    """The fake-quant step `prepare` put on positional input `arg` of `module`: a FusedAmaxObsFakeQuantize, or None
    when that input is not quantized."""
    hooks = module._modules.get("activation_pre_process")
    _only_our_hooks(module, hooks is not None)
    if hooks is None:
        return None
    if arg not in hooks:
        raise _NotReady
    fq = hooks[arg]
    if isinstance(fq, nn.Identity):
        return None
    if not isinstance(fq, FusedAmaxObsFakeQuantize):
        raise _NotFusable
    observe, quantize = fq._flags()
    if observe or fq.is_per_channel or fq.is_block_scaled or fq.record_histogram or fq.scale.numel() != 1:
        raise _NotFusable
    return fq if quantize else None


def same_points(*fqs):
    """Several consumers of one tensor (q/k/v projections, gate/up) quantize it identically -> one step."""
    first = fqs[0]
    for fq in fqs[1:]:
        if (fq is None) != (first is None):
            raise _NotFusable
        if fq is not None and (fq.dtype != first.dtype or fq.qscheme is not None or first.qscheme is not None):
            raise _NotFusable  # frozen scales could differ per consumer; bare specs cannot
    return first


def _spec(*fqs):
    """(fmt, lut, [scale or None per step]) for the steps of one kernel; they must share a format."""
    present = [f for f in fqs if f is not None]
    if not present:
        return _identity_fmt(), None, [None] * len(fqs)
    d = present[0].dtype
    if any(f.dtype != d for f in present):
        raise _NotFusable
    scales = [None if (f is None or f.qscheme is None) else f.scale.reshape(1) for f in fqs]
    return present[0]._fmt, present[0].lut, scales


def _flags(pre=None, mid=None, post=None):
    return (_C.FQ_PRE if pre is not None else 0) | (_C.FQ_MID if mid is not None else 0) | \
        (_C.FQ_POST if post is not None else 0)


# ---- op wrappers (allocate the output, resolve the fake-quant steps) ------------------------------------------

_FP8_OP = {("e4m3", "e4m3"): _C.GEMM_E4M3, ("e5m2", "e5m2"): _C.GEMM_E5M2,
           ("e4m3", "e5m2"): _C.GEMM_E4M3_E5M2, ("e5m2", "e4m3"): _C.GEMM_E5M2_E4M3}


def fp8_kind(fq):
    """"e4m3" / "e5m2" when `fq` is an UNSCALED fake-quantizer of that format: its output can be handed to the FP8
    tensor cores as one-byte codes without changing a single value."""
    if fq is None or not isinstance(fq, FusedAmaxObsFakeQuantize) or fq.qscheme is not None:
        return None
    return fq.fp8_kind


def gemm_operands(a_fq, b_fq, k):
    """(operand_type, use_codes) for a product whose operands leave fake-quantizers a_fq / b_fq."""
    ka, kb = fp8_kind(a_fq), fp8_kind(b_fq)
    if ka is not None and kb is not None and k % 16 == 0:
        return _FP8_OP[(ka, kb)], True
    return _C.GEMM_BF16, False


def _out_like(x, codes, shape=None):
    return torch.empty(x.shape if shape is None else shape, dtype=torch.uint8 if codes else torch.bfloat16,
                       device=x.device)


def norm(x2, weight, bias, eps, kind, pre, post, codes=False, want_raw=False):
    """fq_post(norm(fq_pre(x2))); with want_raw also the normalised tensor before the output step."""
    fmt, lut, (s_pre, s_post) = _spec(pre, post)
    y = _out_like(x2, codes)
    raw = torch.empty_like(x2) if want_raw and post is not None else None
    _C.norm_fq(x2, y, kind, weight, bias, eps, _flags(pre=pre, post=post), fmt, s_pre, s_post, lut, raw)
    if want_raw:
        return y, (raw if raw is not None else y)
    return y


def softmax(scores, alpha, mask, pre, mid, post, codes=False, causal=False, causal_flag=None):
    """scores [B, H, Sq, Sk] contiguous; mask None or additive [Bm, 1, Sq or 1, >=Sk] (Bm in {1, B}).
    causal=True: the caller has a device flag saying whether `mask` is the standard causal mask (see _causal_flag) -- masked scores are not
    read and probabilities beyond the row tile's diagonal block are not written."""
    B, H, Sq, Sk = scores.shape
    m3, mb, mrows = None, 1, Sq
    if mask is not None and mask.dtype == torch.bool:
        from .modules.quantizable._common import additive_mask
        mask = additive_mask(mask, torch.bfloat16)
    if mask is not None:
        if mask.dim() != 4 or mask.shape[1] != 1 or mask.shape[2] not in (1, Sq) or mask.shape[0] not in (1, B) \
                or not mask.dtype.is_floating_point or mask.shape[3] < Sk:
            raise _NotFusable
        m3 = mask[:, 0, :, :Sk]
        if m3.dtype != torch.bfloat16:
            m3 = m3.to(torch.bfloat16)
        m3 = m3.contiguous()
        mb, mrows = m3.shape[0], m3.shape[1]
    fmt, lut, (s_pre, s_mid, s_post) = _spec(pre, mid, post)
    probs = _out_like(scores, codes)
    flags = _flags(pre, mid, post) | (_C.SOFTMAX_CAUSAL if causal else 0)
    _C.softmax_fq(scores, probs, alpha, m3, H * Sq, mrows, mb, flags, fmt, s_pre, s_mid, s_post, lut,
                  causal_flag=causal_flag)
    return probs


def act_mul(gate, up, activation, post, codes=False, pre=None):
    """fq_post(act(fq_pre(gate)) * up); `pre` (the activation module's own input hook) must be a bare spec."""
    fmt, lut, (_, s_post) = _spec(pre, post)
    if pre is not None and pre.qscheme is not None:
        raise _NotFusable
    out = _out_like(gate, codes)
    _C.act_mul_fq(gate, up, out, activation, _flags(pre=pre, post=post), fmt, s_post, lut)
    return out


def add_norm(x2, res2, res_a, res_b, norm_mod, post, want_raw=True):
    """fq_post(norm(fq_pre(bf16(fq_a(x2) + fq_b(res2))))) in one pass: the hooked residual add (AddFunctional), the
    norm with its input hook and the consumer's input hook.  norm_mod None: the add alone."""
    if norm_mod is None:
        kind, pre, w, b, eps = _C.NORM_IDENTITY, None, None, None, 0.0
    else:
        kind = _C.NORM_LAYER if isinstance(norm_mod, nn.LayerNorm) else _C.NORM_NONE
        if kind == _C.NORM_NONE and type(norm_mod).__name__ != "NoNorm":
            raise _NotFusable
        pre, w, b, eps = point(norm_mod), norm_mod.weight, norm_mod.bias, getattr(norm_mod, "eps", 0.0)
    if any(f is not None and f.qscheme is not None for f in (res_a, res_b)):
        raise _NotFusable   # the add's hooks run on the bare path of the kernel
    fmt, lut, (_, _, s_pre, s_post) = _spec(res_a, res_b, pre, post)
    if not res2.is_contiguous():
        res2 = res2.contiguous()
    y = torch.empty_like(x2)
    raw = torch.empty_like(x2) if want_raw and post is not None else None
    flags = _flags(pre=pre, post=post) | (_C.FQ_RES_A if res_a is not None else 0) | (_C.FQ_RES_B if res_b is not None else 0)
    _C.add_norm_fq(x2, res2, y, kind, w, b, eps, flags, fmt, s_pre, s_post, lut, raw)
    if want_raw:
        return y, (raw if raw is not None else y)
    return y


def fake_quant(x2, post, codes=False):
    """Plain fake quant of a 2-D activation through the module's own kernel path (observer-free by construction)."""
    if codes:
        return post.quantize_to_codes(x2)
    return x2 if post is None else post(x2)


# ---- quantized weights of a block, concatenated once -----------------------------------------------------------

def _weight_fq(lin):
    fq = lin.weight_fake_quant
    if isinstance(fq, FusedAmaxObsFakeQuantize) and fq._flags() == (False, True):
        return fq
    return None


def _common_weight_fq(*linears):
    """The weight fake-quantizer shared by several Linears when all are unscaled fp8 of one format, else None."""
    fqs = [_weight_fq(lin) for lin in linears]
    kinds = {fp8_kind(f) for f in fqs}
    return fqs[0] if len(kinds) == 1 and None not in kinds else None


def _interleave(ts, block):
    """[t0 rows 0..block), [t1 rows 0..block), [t0 rows block..2 block), ...: the gate|up layout of the gated epilogue."""
    n = ts[0].shape[0]
    parts = [t.reshape(n // block, block, *t.shape[1:]) for t in ts]
    return torch.stack(parts, 1).reshape(len(ts) * n, *ts[0].shape[1:])


def _quantized_cat(owner, tag, linears, codes=False, interleave=0):
    """cat([fq(W) for each Linear]) along the output axis, cached on `owner` until a weight or a scale is written.
    The weight fake-quantizers must be observer-free (checked) so that skipping their per-forward re-run is
    unobservable."""
    key = []
    for lin in linears:
        fq, w = lin.weight_fake_quant, lin.weight
        if isinstance(fq, FusedAmaxObsFakeQuantize):
            observe, quantize = fq._flags()
            if observe:
                raise _NotFusable
            key += [w.data_ptr(), w._version, fq.state_key()]
        elif isinstance(fq, nn.Identity):
            key += [w.data_ptr(), w._version]
        else:
            raise _NotFusable
        _only_our_hooks(lin, lin._modules.get("activation_pre_process") is not None)
    key = tuple(key + [codes, interleave])
    cache = owner.__dict__.setdefault("_qt_wcache", {})
    hit = cache.get(tag)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            if codes:
                ws = [lin.weight_fake_quant.quantize_to_codes(lin.weight.detach()) for lin in linears]
            else:
                ws = [lin.weight_fake_quant(lin.weight).detach() for lin in linears]
            join = (lambda ts: _interleave(ts, interleave)) if interleave else (lambda ts: torch.cat(ts, 0))
            w = ws[0] if len(ws) == 1 else join(ws)
            bs = [lin.bias for lin in linears]
            b = None
            if any(x is not None for x in bs):
                b = join([x.detach() if x is not None else torch.zeros(lin.weight.shape[0], dtype=lin.weight.dtype,
                                                                         device=w.device)
                          for x, lin in zip(bs, linears)]).contiguous()
        hit = (key, w.contiguous(), b)
        cache[tag] = hit
    return hit[1], hit[2]


def _epilogue_fq(fq):
    """(fmt, lut) when `fq` can be applied by the producing GEMM's epilogue (bare spec), None when there is nothing to
    apply; raises _NotFusable for a scaled one (callers then keep the separate pass)."""
    if fq is None:
        return None
    if fq.qscheme is not None:
        raise _NotFusable
    return (fq._fmt, fq.lut)


def _usable(x):
    return _ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and not torch.is_grad_enabled()


def attention(q4, k4, vt, scaling, mask, sc_in, sm_in, p_in, o_in, t_qk, t_pv, c_o, B, S, H, D):
    """The attention core: QK^T GEMM -> scale+mask+softmax+fq -> PV GEMM, three launches.  (Round 1 also carried a
    single-kernel two-pass attention, qt_attention_fq; it recomputed the element-wise chain in both passes with 8
    softmax warps per SM and ran 2.2x slower than this chain -- 176 vs 80 us at the Llama-2-7B window -- so it was
    removed in round 2 rather than kept as dead weight.)"""
    Sk = k4.shape[2]
    # Causal schedule: with the standard causal mask and no fake quant between the mask and the softmax, everything
    # above the diagonal is exactly zero probability -- the score tiles there are not computed, the softmax neither
    # reads nor writes them and the P x V reduction of a row tile stops at its diagonal block.
    # Whether the mask IS the causal one is decided on the device (qt_causal_mask_check writes a flag the three kernels
    # read): no host round trip, and a captured graph stays correct if a later replay carries a padding mask.
    flag = None
    if mask is not None and sm_in is None and Sk == S and S % 128 == 0 and os.environ.get("QT_CAUSAL", "1") != "0":
        flag = _causal_flag(mask, S)
    causal = flag is not None
    scores = _C.gemm_nt(q4, k4, operand_type=t_qk, causal=_C.CAUSAL_OUT_LOWER if causal else 0, causal_flag=flag)
    probs = softmax(scores, scaling, mask, sc_in, sm_in, p_in, t_pv != _C.GEMM_BF16, causal=causal, causal_flag=flag)
    return _attention_context(probs, vt, t_pv, o_in, c_o, B, S, H, D, flag)


_FLAG_CACHE = {}


def _causal_flag(mask, S):
    """Device flag for `mask` ([Bm, 1, S, S] additive bf16), computed once per mask tensor (all layers of a forward
    see the same object); None when the mask cannot be the square causal one."""
    if mask.dim() != 4 or mask.shape[1] != 1 or mask.shape[2] != S or mask.shape[3] != S \
            or mask.dtype != torch.bfloat16 or not mask[:, 0].is_contiguous():
        return None
    # Under CUDA-graph capture the check kernel must be PART of the graph (a replay may carry a padding mask in the
    # same static buffer, and a flag computed at warm-up would then keep skipping tiles): the capture id is part of
    # the key, so each captured graph runs the check once, for all of its layers.
    key = (mask.data_ptr(), mask._version, tuple(mask.shape), _C.stream_capture_id(mask))
    hit = _FLAG_CACHE.get("last")
    if hit is None or hit[0] != key:
        hit = (key, _C.causal_mask_check(mask[:, 0]), mask)   # keeps the mask alive: its address is the key
        _FLAG_CACHE["last"] = hit
    return hit[1]


def _attention_context(probs, vt, t_pv, o_in, c_o, B, S, H, D, causal_flag=None):
    """probabilities x values written straight into [B*S, H*D].  The output projection's input fake quant is applied
    by the product's epilogue only for long reductions: at S = 1024 the product is epilogue-bound (128 x 128 tiles,
    16 k-blocks) and the table lookups of the re-quantization cost more there than the separate 7 us pass
    (measured: 26.6 vs 14 + 6.7 us per Llama-2-7B layer)."""
    causal = causal_flag is not None
    if o_in is None or (o_in.qscheme is None and S >= 4096):
        ctx = torch.empty(B, S, H * D, dtype=torch.uint8 if c_o else torch.bfloat16, device=probs.device)
        _C.gemm_nt(probs, vt, out=ctx.view(B, S, H, D).transpose(1, 2), operand_type=t_pv,
                   fq=_epilogue_fq(o_in), out_codes=c_o, causal=_C.CAUSAL_A_LOWER if causal else 0,
                   causal_flag=causal_flag)
        return ctx.view(B * S, H * D)
    ctx = torch.empty(B, S, H * D, dtype=torch.bfloat16, device=probs.device)
    _C.gemm_nt(probs, vt, out=ctx.view(B, S, H, D).transpose(1, 2), operand_type=t_pv,
               causal=_C.CAUSAL_A_LOWER if causal else 0, causal_flag=causal_flag)
    return fake_quant(ctx.view(B * S, H * D), o_in, c_o)


# ---- Llama decoder layer ------------------------------------------------------------------------------------

def llama_layer_forward(layer, hidden_states, attention_mask, position_embeddings, past_key_values=None):
    """Fused forward of a quantizable LlamaDecoderLayer, or None when the layer must run module by module."""
    if not _usable(hidden_states) or past_key_values is not None or position_embeddings is None:
        return None
    attn, mlp = layer.self_attn, layer.mlp
    try:
        if attn.num_key_value_groups != 1 or (attn.training and attn.attention_dropout > 0.0):
            return None
        if hidden_states.dim() != 3 or getattr(mlp.config, "pretraining_tp", 1) > 1:
            return None
        act_name = getattr(mlp.config, "hidden_act", "silu")
        if act_name not in ("silu", "gelu", "relu"):
            return None
        for m in (attn, mlp, attn.attn_scaling, attn.softmax, attn.qk_matmul, attn.av_matmul):
            if len(m._forward_hooks) or len(m._backward_hooks):
                return None
        # fake-quant steps, from the hooks
        ln1_in, ln2_in = point(layer.input_layernorm), point(layer.post_attention_layernorm)
        x_in = same_points(point(attn.q_proj), point(attn.k_proj), point(attn.v_proj))
        q_in, k_in = point(attn.qk_matmul, "0"), point(attn.qk_matmul, "1")
        p_in, v_in = point(attn.av_matmul, "0"), point(attn.av_matmul, "1")
        sc_in, sm_in = point(attn.attn_scaling), point(attn.softmax)
        o_in = point(attn.o_proj)
        gu_in = same_points(point(mlp.gate_proj), point(mlp.up_proj))
        d_in = point(mlp.down_proj)
        res1 = (point(layer.self_attn_residual, "0"), point(layer.self_attn_residual, "1"))
        res2 = (point(layer.mlp_residual, "0"), point(layer.mlp_residual, "1"))
        B, S, hidden = hidden_states.shape
        H, D = attn.config.num_attention_heads, attn.head_dim
        T = B * S
        inter = mlp.gate_proj.weight.shape[0]
        # which products run on the FP8 tensor cores (both operands leave unscaled e4m3 / e5m2 fake-quantizers)
        wq = _common_weight_fq(attn.q_proj, attn.k_proj, attn.v_proj)
        wgu = _common_weight_fq(mlp.gate_proj, mlp.up_proj)
        t_qkv, c_qkv = gemm_operands(x_in, wq, hidden)
        t_qk, c_qk = gemm_operands(q_in, k_in, D)
        t_pv, c_pv = gemm_operands(p_in, v_in, S)
        t_o, c_o = gemm_operands(o_in, _weight_fq(attn.o_proj), H * D)
        t_gu, c_gu = gemm_operands(gu_in, wgu, hidden)
        t_d, c_d = gemm_operands(d_in, _weight_fq(mlp.down_proj), inter)
        w_qkv, b_qkv = _quantized_cat(layer, "qkv", (attn.q_proj, attn.k_proj, attn.v_proj), c_qkv)
        w_o, b_o = _quantized_cat(layer, "o", (attn.o_proj,), c_o)
        glu = act_name == "silu" and inter % 64 == 0 and (d_in is None or d_in.qscheme is None)
        w_gu, b_gu = _quantized_cat(layer, "gu", (mlp.gate_proj, mlp.up_proj), c_gu, interleave=64 if glu else 0)
        w_d, b_d = _quantized_cat(layer, "d", (mlp.down_proj,), c_d)

        x = hidden_states.reshape(T, hidden)
        if not x.is_contiguous():
            x = x.contiguous()
        cos, sin = position_embeddings
        cos2, sin2 = cos.reshape(-1, D), sin.reshape(-1, D)
        if cos2.dtype != torch.bfloat16 or cos2.shape[0] not in (S, T) or not cos2.is_contiguous():
            return None
        if (q_in is None) != (k_in is None):
            raise _NotFusable

        # attention
        n1 = layer.input_layernorm
        xq = norm(x, n1.weight, None, n1.variance_epsilon, _C.NORM_RMS, ln1_in, x_in, c_qkv)
        qkv = _C.gemm_nt(xq, w_qkv, bias=b_qkv, operand_type=t_qkv)              # [T, 3 * H * D]
        q = qkv[:, :H * D].view(T, H, D)
        k = qkv[:, H * D:2 * H * D].view(T, H, D)
        v = qkv[:, 2 * H * D:].view(B, S, H, D)
        fmt, lut, (s_q, s_k) = _spec(q_in, k_in)
        qk = _out_like(x, c_qk, (2, T, H, D))
        _C.rope_fq(q, qk[0], k, qk[1], cos2, sin2, _C.FQ_POST if q_in is not None else 0, fmt, s_q, s_k, lut)
        fmt, lut, (s_v,) = _spec(v_in)
        vt = _out_like(x, c_pv, (B, H, D, S))
        _C.fq_transpose(v, vt, _flags(post=v_in), fmt, s_v, lut)
        q4 = qk[0].view(B, S, H, D).transpose(1, 2)
        k4 = qk[1].view(B, S, H, D).transpose(1, 2)
        ctx2 = attention(q4, k4, vt, attn.scaling, attention_mask, sc_in, sm_in, p_in, o_in, t_qk, t_pv, c_o, B, S, H, D)
        if res1 == (None, None):
            h1 = _C.gemm_nt(ctx2, w_o, bias=b_o, residual=x, operand_type=t_o)   # residual add in the epilogue
        else:
            h1 = layer.self_attn_residual(x, _C.gemm_nt(ctx2, w_o, bias=b_o, operand_type=t_o))

        # MLP
        n2 = layer.post_attention_layernorm
        x2 = norm(h1, n2.weight, None, n2.variance_epsilon, _C.NORM_RMS, ln2_in, gu_in, c_gu)
        if glu:   # silu(gate) * up and down_proj's input fake quant inside the gate|up GEMM's epilogue
            a = _C.gemm_nt(x2, w_gu, bias=b_gu, operand_type=t_gu, activation="silu", glu=True,
                           fq=_epilogue_fq(d_in), out_codes=c_d)
        else:
            gu = _C.gemm_nt(x2, w_gu, bias=b_gu, operand_type=t_gu)              # [T, 2 * I]
            a = act_mul(gu[:, :inter], gu[:, inter:], act_name, d_in, c_d)
        if res2 == (None, None):
            h2 = _C.gemm_nt(a, w_d, bias=b_d, residual=h1, operand_type=t_d)
        else:
            h2 = layer.mlp_residual(h1, _C.gemm_nt(a, w_d, bias=b_d, operand_type=t_d))
        return h2.view(B, S, hidden)
    except (_NotReady, _NotFusable, AttributeError):
        _why("layer")
        return None


# ---- BERT / RoBERTa encoder layer ------------------------------------------------------------------------------

def strided_fq(x2, post, codes=False):
    """fake quant of a 2-D view with a row stride (a column slice of a fused projection) into a contiguous tensor."""
    fmt, lut, (s_post,) = _spec(post)
    out = _out_like(x2, codes)
    _C.act_mul_fq(x2, None, out, None, _flags(post=post), fmt, s_post, lut)
    return out


def bert_layer_forward(layer, hidden_states, attention_mask):
    """Fused forward of a quantizable BERT / RoBERTa encoder layer (self-attention only), or None.
    Reference structure: modules/quantizable/modeling_bert.py:32-222 inside HF BertLayer."""
    if not _usable(hidden_states) or hidden_states.dim() != 3:
        return None
    att, so, inter, out = layer.attention.self, layer.attention.output, layer.intermediate, layer.output
    try:
        if layer.training and (att.dropout.p > 0.0 or so.dropout.p > 0.0 or out.dropout.p > 0.0):
            return None
        if getattr(att, "position_embedding_type", "absolute") not in (None, "absolute") or layer.chunk_size_feed_forward:
            return None
        for m in (layer.attention, att, so, inter, out, att.attn_scaling, att.softmax, att.qk_matmul, att.av_matmul):
            if len(m._forward_hooks) or len(m._backward_hooks):
                return None
        act_mod = inter.intermediate_act_fn
        act_name = {"GELUActivation": "gelu", "GELU": "gelu", "ReLU": "relu"}.get(type(act_mod).__name__)
        if act_name is None or (isinstance(act_mod, nn.GELU) and act_mod.approximate != "none"):
            return None
        x_in = same_points(point(att.query), point(att.key), point(att.value))
        q_in, k_in = point(att.qk_matmul, "0"), point(att.qk_matmul, "1")
        p_in, v_in = point(att.av_matmul, "0"), point(att.av_matmul, "1")
        sc_in, sm_in = point(att.attn_scaling), point(att.softmax)
        o_in, i_in, o2_in = point(so.dense), point(inter.dense), point(out.dense)
        ln1_in, ln2_in = point(so.LayerNorm), point(out.LayerNorm)
        act_in = point(act_mod) if isinstance(act_mod, nn.Module) else None
        res1 = (point(so.residual, "0"), point(so.residual, "1"))
        res2 = (point(out.residual, "0"), point(out.residual, "1"))
        if (q_in is None) != (k_in is None):
            raise _NotFusable

        B, S, hidden = hidden_states.shape
        H, D = att.num_attention_heads, att.attention_head_size
        T = B * S
        isz = inter.dense.weight.shape[0]
        wq = _common_weight_fq(att.query, att.key, att.value)
        t_qkv, c_qkv = gemm_operands(x_in, wq, hidden)
        t_qk, c_qk = gemm_operands(q_in, k_in, D)
        t_pv, c_pv = gemm_operands(p_in, v_in, S)
        t_o, c_o = gemm_operands(o_in, _weight_fq(so.dense), hidden)
        t_i, c_i = gemm_operands(i_in, _weight_fq(inter.dense), hidden)
        t_o2, c_o2 = gemm_operands(o2_in, _weight_fq(out.dense), isz)
        w_qkv, b_qkv = _quantized_cat(layer, "qkv", (att.query, att.key, att.value), c_qkv)
        w_o, b_o = _quantized_cat(layer, "o", (so.dense,), c_o)
        w_i, b_i = _quantized_cat(layer, "i", (inter.dense,), c_i)
        w_o2, b_o2 = _quantized_cat(layer, "o2", (out.dense,), c_o2)

        x = hidden_states.reshape(T, hidden)
        if not x.is_contiguous():
            x = x.contiguous()
        scaling = getattr(att, "scaling", D ** -0.5)

        # self-attention
        xq = x if x_in is None else (x_in.quantize_to_codes(x) if c_qkv else x_in(x))
        qkv = _C.gemm_nt(xq, w_qkv, bias=b_qkv, operand_type=t_qkv)             # [T, 3 * hidden]
        qk = strided_fq(qkv[:, :2 * hidden], q_in if q_in is not None else None, c_qk)   # [T, 2 * hidden]
        if q_in is not None and k_in is not None and (q_in.dtype != k_in.dtype or q_in.qscheme is not None
                                                      or k_in.qscheme is not None):
            raise _NotFusable  # one launch quantizes q and k: they must share a bare format
        fmt, lut, (s_v,) = _spec(v_in)
        vt = _out_like(x, c_pv, (B, H, D, S))
        _C.fq_transpose(qkv[:, 2 * hidden:].view(B, S, H, D), vt, _flags(post=v_in), fmt, s_v, lut)
        q4 = qk[:, :hidden].view(B, S, H, D).transpose(1, 2)
        k4 = qk[:, hidden:].view(B, S, H, D).transpose(1, 2)
        ctx2 = attention(q4, k4, vt, scaling, attention_mask, sc_in, sm_in, p_in, o_in, t_qk, t_pv, c_o, B, S, H, D)
        n1 = so.LayerNorm
        if res1 == (None, None):
            h1 = _C.gemm_nt(ctx2, w_o, bias=b_o, residual=x, operand_type=t_o)
            a_q, a_raw = norm(h1, n1.weight, n1.bias, n1.eps, _C.NORM_LAYER, ln1_in, i_in, c_i, want_raw=True)
        elif not c_i and all(f is None or f.qscheme is None for f in res1):
            # hooked residual add + LayerNorm + both hooks around it: one pass
            a_q, a_raw = add_norm(_C.gemm_nt(ctx2, w_o, bias=b_o, operand_type=t_o), x, res1[0], res1[1], n1, i_in)
        else:
            h1 = so.residual(_C.gemm_nt(ctx2, w_o, bias=b_o, operand_type=t_o), x)
            a_q, a_raw = norm(h1, n1.weight, n1.bias, n1.eps, _C.NORM_LAYER, ln1_in, i_in, c_i, want_raw=True)

        # feed-forward
        if act_in is None:
            mid = _C.gemm_nt(a_q, w_i, bias=b_i, activation=act_name, operand_type=t_i)   # activation in the epilogue
        else:
            mid = act_mod(_C.gemm_nt(a_q, w_i, bias=b_i, operand_type=t_i))                # hooked activation module
        mid_q = fake_quant(mid, o2_in, c_o2)
        n2 = out.LayerNorm
        if res2 == (None, None):
            h2 = _C.gemm_nt(mid_q, w_o2, bias=b_o2, residual=a_raw, operand_type=t_o2)
            y = norm(h2, n2.weight, n2.bias, n2.eps, _C.NORM_LAYER, ln2_in, None)
        elif all(f is None or f.qscheme is None for f in res2):
            y = add_norm(_C.gemm_nt(mid_q, w_o2, bias=b_o2, operand_type=t_o2), a_raw, res2[0], res2[1], n2, None,
                         want_raw=False)
        else:
            h2 = out.residual(_C.gemm_nt(mid_q, w_o2, bias=b_o2, operand_type=t_o2), a_raw)
            y = norm(h2, n2.weight, n2.bias, n2.eps, _C.NORM_LAYER, ln2_in, None)
        return y.view(B, S, hidden)
    except (_NotReady, _NotFusable, AttributeError):
        _why("layer")
        return None


# ---- MobileBERT encoder layer ---------------------------------------------------------------------------------

def _linear_block(owner, tag, x_q, in_fq, lin, res_mod=None, res=None, act=None, act_mod=None):
    """y = act(x_q W_q^T + b) [+ res]: one GEMM; the residual add and a plain activation ride in its epilogue when no
    fake-quant hook sits on them (their op group is "fused" in the paper's terms), otherwise the hooked module runs."""
    t, codes = gemm_operands(in_fq, _weight_fq(lin), lin.weight.shape[1])
    w, b = _quantized_cat(owner, tag, (lin,), codes)
    if codes and x_q.dtype != torch.uint8:
        x_q = x_q.to(_FP8_TORCH[in_fq.fp8_kind]).view(torch.uint8)   # exact: x_q already holds that format's values
    epi_act = act if (act is not None and (act_mod is None or point(act_mod) is None)) else None
    epi_res = res if (res is not None and (point(res_mod, "0"), point(res_mod, "1")) == (None, None)) else None
    y = _C.gemm_nt(x_q, w, bias=b, activation=epi_act, residual=epi_res, operand_type=t)
    # hooked activation / hooked residual add: left to the caller, which fuses them with what follows
    return y, (act is not None and epi_act is None), (res is not None and epi_res is None)


_FP8_TORCH = {"e4m3": torch.float8_e4m3fn, "e5m2": torch.float8_e5m2}


def _nonorm(norm_mod, y, post, want_raw=True):
    """NoNorm (x * weight + bias) with its input hook and the consumer's input hook in one pass; LayerNorm likewise."""
    kind = _C.NORM_LAYER if isinstance(norm_mod, nn.LayerNorm) else _C.NORM_NONE
    if kind == _C.NORM_NONE and type(norm_mod).__name__ != "NoNorm":
        raise _NotFusable
    eps = getattr(norm_mod, "eps", 0.0)
    codes = False
    return norm(y, norm_mod.weight, norm_mod.bias, eps, kind, point(norm_mod), post, codes, want_raw=want_raw)


def mobilebert_layer_forward(layer, hidden_states, attention_mask):
    """Fused forward of a MobileBERT encoder layer whose blocks are the quantizable ones, or None.
    Structure: transformers MobileBertLayer (bottleneck -> self-attention -> self-output -> (num_ffn - 1) x FFN ->
    intermediate -> output + output bottleneck) around the reference's blocks (modules/quantizable/
    modeling_mobilebert.py:38-206).  Same weights, same fake-quantizers, same rounding points as the module-by-module
    execution: the three consumers of the layer input share one fake quant, every NoNorm runs with its input hook and
    the next Linear's input hook in one pass, query | key are one GEMM, the attention core is the three-kernel chain,
    and the residual adds / ReLU ride in GEMM epilogues wherever their op group is not hooked."""
    if not _usable(hidden_states) or hidden_states.dim() != 3:
        return None
    try:
        if not getattr(layer, "use_bottleneck", False):
            return None
        bn, att, so, out = layer.bottleneck, layer.attention.self, layer.attention.output, layer.output
        if bn.use_bottleneck_attention or not bn.key_query_shared_bottleneck or not out.use_bottleneck:
            return None
        ffns = list(layer.ffn) if layer.num_feedforward_networks > 1 else []
        mods = [layer, layer.attention, bn, bn.input, bn.attention, att, so, out, out.bottleneck, layer.intermediate,
                att.attn_scaling, att.softmax, att.qk_matmul, att.av_matmul] + ffns + [f.intermediate for f in ffns] + \
               [f.output for f in ffns]
        for m in mods:
            if len(m._forward_hooks) or len(m._backward_hooks):
                return None
        if layer.training and (att.dropout.p > 0.0 or out.bottleneck.dropout.p > 0.0):
            return None

        def act_of(inter):
            am = inter.intermediate_act_fn
            name = {"ReLU": "relu", "GELUActivation": "gelu", "GELU": "gelu"}.get(type(am).__name__)
            if name is None or not isinstance(am, nn.Module):
                raise _NotFusable
            return name, am

        B, S, hidden = hidden_states.shape
        T = B * S
        H, D = att.num_attention_heads, att.attention_head_size
        th = att.query.weight.shape[1]                     # true hidden size
        x = hidden_states.reshape(T, hidden)
        if not x.is_contiguous():
            x = x.contiguous()

        # the layer input feeds three Linears (bottleneck.input, bottleneck.attention, value): one fake quant
        x_in = same_points(point(bn.input.dense), point(bn.attention.dense), point(att.value))
        xq = x if x_in is None else x_in(x)
        layer_in, _, _ = _linear_block(layer, "bn_in", xq, x_in, bn.input.dense)
        layer_in = _nonorm(bn.input.LayerNorm, layer_in, None, want_raw=False)              # residual of self-output
        qk_in = same_points(point(att.query), point(att.key))
        shared, _, _ = _linear_block(layer, "bn_att", xq, x_in, bn.attention.dense)
        shared_q = _nonorm(bn.attention.LayerNorm, shared, qk_in, want_raw=False)            # input of query and key

        # self-attention: query | key in one GEMM, value from the layer input
        t_qkp, c_qkp = gemm_operands(qk_in, _common_weight_fq(att.query, att.key), th)
        w_qk, b_qk = _quantized_cat(layer, "qk", (att.query, att.key), c_qkp)
        if c_qkp:
            shared_q = shared_q.to(_FP8_TORCH[qk_in.fp8_kind]).view(torch.uint8)
        qk = _C.gemm_nt(shared_q, w_qk, bias=b_qk, operand_type=t_qkp)                       # [T, 2 * H * D]
        v, _, _ = _linear_block(layer, "v", xq, x_in, att.value)                             # [T, H * D]
        q_in, k_in = point(att.qk_matmul, "0"), point(att.qk_matmul, "1")
        p_in, v_in = point(att.av_matmul, "0"), point(att.av_matmul, "1")
        sc_in, sm_in = point(att.attn_scaling), point(att.softmax)
        o_in = point(so.dense)
        if (q_in is None) != (k_in is None) or (q_in is not None and (q_in.dtype != k_in.dtype or q_in.qscheme is not None
                                                                      or k_in.qscheme is not None)):
            raise _NotFusable
        t_qk, c_qk = gemm_operands(q_in, k_in, D)
        t_pv, c_pv = gemm_operands(p_in, v_in, S)
        t_o, c_o = gemm_operands(o_in, _weight_fq(so.dense), H * D)
        qkq = strided_fq(qk, q_in, c_qk)
        fmt, lut, (s_v,) = _spec(v_in)
        vt = _out_like(x, c_pv, (B, H, D, S))
        _C.fq_transpose(v.view(B, S, H, D), vt, _flags(post=v_in), fmt, s_v, lut)
        q4 = qkq[:, :H * D].view(B, S, H, D).transpose(1, 2)
        k4 = qkq[:, H * D:].view(B, S, H, D).transpose(1, 2)
        scaling = getattr(att, "scaling", D ** -0.5)
        ctx2 = attention(q4, k4, vt, scaling, attention_mask, sc_in, sm_in, p_in, o_in, t_qk, t_pv, c_o, B, S, H, D)

        def dense_res_norm(tag, x_q, in_fq, lin, res_mod, res, norm_mod, post, want_raw=True):
            """dense (+ residual in its epilogue when the add is un-hooked) -> [hooked add +] NoNorm + hooks, one pass"""
            y, _, hooked_add = _linear_block(layer, tag, x_q, in_fq, lin, res_mod, res)
            if hooked_add:
                return add_norm(y, res, point(res_mod, "0"), point(res_mod, "1"), norm_mod, post, want_raw)
            return _nonorm(norm_mod, y, post, want_raw)

        # self-output: dense + residual(layer_in) + NoNorm
        first_ffn = ffns[0].intermediate.dense if ffns else layer.intermediate.dense
        a_q, a_raw = dense_res_norm("so", ctx2, o_in, so.dense, so.residual, layer_in, so.LayerNorm, point(first_ffn))

        # feed-forward stacks: intermediate (dense + act) -> output dense + residual + NoNorm
        stacks = [(f.intermediate, f.output, f"ffn{i}") for i, f in enumerate(ffns)] + [(layer.intermediate, out, "ffn_last")]
        for i, (inter, outp, tag) in enumerate(stacks):
            act_name, act_mod = act_of(inter)
            mid, hooked_act, _ = _linear_block(layer, tag + "_i", a_q, point(inter.dense), inter.dense, act=act_name,
                                               act_mod=act_mod)
            o_fq = point(outp.dense)
            o_codes = gemm_operands(o_fq, _weight_fq(outp.dense), mid.shape[1])[1]
            if hooked_act:   # act module's input hook + activation + the next Linear's input hook: one pass
                mid_q = act_mul(mid, None, act_name, o_fq, o_codes, pre=point(act_mod))
            else:
                mid_q = fake_quant(mid, o_fq, o_codes)
            nxt = point(stacks[i + 1][0].dense) if i + 1 < len(stacks) else point(out.bottleneck.dense)
            a_q, a_raw = dense_res_norm(tag + "_o", mid_q, o_fq, outp.dense, outp.residual, a_raw, outp.LayerNorm, nxt)

        # output bottleneck: dense (true hidden -> hidden) + residual(layer input) + NoNorm
        ob = out.bottleneck
        y = dense_res_norm("ob", a_q, point(ob.dense), ob.dense, ob.residual, x, ob.LayerNorm, None, want_raw=False)
        return y.view(B, S, hidden)
    except (_NotReady, _NotFusable, AttributeError):
        _why("layer")
        return None
