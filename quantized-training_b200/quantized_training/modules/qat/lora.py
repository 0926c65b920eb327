"""QAT LoRA linear: fake-quantize A, B and the merged weight W + s*B@A on every forward, then ONE
GEMM with the merged weight (reference: modules/qat/lora.py:34-55).  Note what that implies, and is
kept: LoRA dropout is not applied on this path and the ``lora_A`` / ``lora_B`` sub-Linears are never
called (their ``.weight`` is read directly), so hooks on them do not fire.

The merge -- clone, fq(A), fq(B), B @ A, * scaling, +=, fq: seven launches and four weight-sized temporaries
in the reference -- is ONE kernel here (`qt_lora_merge_fq`, bf16 roundings of the op chain reproduced in
registers) whenever the weight quantizer is stateless (bare spec, or a frozen per-tensor scale); with a live
observer each of the reference's three calls advances the amax history, so A, B and the merged weight go
through the module one by one and the kernel only merges.  Gradients of A and B (straight-through) are two
products on the tcgen05 GEMM."""
import torch

from ... import _C, ops
from ..lora import LoraLinear as _FloatLora

try:  # the real peft layer, when the package is present
    from peft.tuners.lora import Linear as _PeftLora
except Exception:  # pragma: no cover - peft is not part of this image
    _PeftLora = None

__all__ = ["Linear"]


def _t(w, fan_in_fan_out):
    return w.T if fan_in_fan_out else w


class _MergeFn(torch.autograd.Function):
    """merged = fq(W + (fq(B) @ fq(A)) * s) in one kernel; backward: straight-through estimators, W frozen (its
    `.data` is read, as in the reference), gA = s * fq(B)^T gM and gB = s * gM fq(A)^T on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, w, a, b, fq, scaling):
        observe, quantize = fq._flags()
        if fq.scale.device != w.device:
            fq.to(w.device)
        out = torch.empty_like(w)
        ac, bc = a.detach().contiguous(), b.detach().contiguous()
        if observe:   # stateful: the three calls of the reference, in its order; the kernel merges only
            aq, bq = fq(ac), fq(bc)
            _C.lora_merge_fq(w, aq, bq, out, scaling, 0, fq._fmt, None, fq.lut)
            out = fq(out)
        else:
            aq = bq = None
            bare = fq.qscheme is None
            points = (_C.FQ_PRE | _C.FQ_POST) if quantize else 0
            if quantize and (not bare or any(ctx.needs_input_grad[1:3])):
                # frozen per-tensor scale (A and B are quantized with it too), or training (backward needs fq(A), fq(B))
                aq, bq = fq(ac), fq(bc)
                points = _C.FQ_POST
            _C.lora_merge_fq(w, aq if aq is not None else ac, bq if bq is not None else bc, out, scaling, points,
                             fq._fmt, None if bare else fq.scale.reshape(1), fq.lut)
        if any(ctx.needs_input_grad[1:3]):
            ctx.save_for_backward(aq if aq is not None else ac, bq if bq is not None else bc)
            ctx.scaling = scaling
        return out

    @staticmethod
    def backward(ctx, gm):
        aq, bq = ctx.saved_tensors
        ga = gb = None
        if ctx.needs_input_grad[1]:   # [r, K] = s * Bq^T [r, N] @ gM [N, K]
            ga = ops.gemm(bq.t(), gm.t(), alpha=ctx.scaling)
        if ctx.needs_input_grad[2]:   # [N, r] = s * gM [N, K] @ Aq^T [K, r]
            gb = ops.gemm(gm, aq, alpha=ctx.scaling)
        return None, ga, gb, None, None


class Linear(_FloatLora):
    _FLOAT_MODULE = _FloatLora
    _FLOAT_MODULES = tuple(c for c in (_FloatLora, _PeftLora) if c is not None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        in_dtype = x.dtype
        if self.disable_adapters or self.merged:
            return ops.linear(x, _t(self.weight, self.fan_in_fan_out), self.bias).to(in_dtype)
        names = [n for n in self.active_adapters if n in self.lora_A]
        fq = self.weight_fake_quant
        w = self.weight.data
        if (len(names) == 1 and not self.fan_in_fan_out and w.is_cuda and w.dtype == torch.bfloat16
                and w.is_contiguous() and w.shape[1] % 8 == 0 and getattr(fq, "_fmt", None) is not None
                and not getattr(fq, "is_block_scaled", False) and not fq.is_per_channel and not fq.record_histogram):
            a, b = self.lora_A[names[0]].weight, self.lora_B[names[0]].weight
            merged = _MergeFn.apply(w, a, b, fq, float(self.scaling[names[0]]))
        else:   # several adapters / fan_in_fan_out / per-channel or block-scaled weight specs: op by op
            merged = w.clone()
            for name in names:
                a, b = fq(self.lora_A[name].weight), fq(self.lora_B[name].weight)
                merged = merged + _t(ops.gemm(b, a.t()), self.fan_in_fan_out) * self.scaling[name]
            merged = fq(merged)
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
