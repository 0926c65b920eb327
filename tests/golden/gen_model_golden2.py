#!/usr/bin/env python
"""Model-level golden fixtures, second set (round 2): the UNMODIFIED reference `quantize(model, args)` applied on CPU to

  * a 2-layer MobileBERT-tiny-shaped encoder made of the reference's own quantizable MobileBERT blocks
    (modules/quantizable/modeling_mobilebert.py:38-206) -- BASELINE configs[0] flags: e4m3, all five op groups;
  * a 2-layer Llama decoder made of the reference's own LlamaDecoderLayer / LlamaAttention
    (modules/quantizable/modeling_llama.py:95-356) -- configs[4] flags: posit8_1 / e4m3, `gemm` and all five groups --
    with the logits AND the causal-LM NLL (the quantity behind the reference's perplexity, wikitext.py:146-167);
  * a RoBERTa-style classifier whose query / value Linears carry LoRA adapters, swapped by the reference's
    quantize() into its qat.LoraLinear (modules/qat/lora.py:34-55) -- configs[3] flags: fp8_e4m3 forward, delayed-scaling
    fp8_e5m2 gradients -- forward, loss, backward: gradients of the LoRA factors, the head and the input;
  * for that training case, every call of a gradient fake-quantizer during the recorded backward (input, scale in
    use, output) so that the GPU test can replay each one in isolation, and the reference's OWN sensitivity: the same
    two steps re-run with the input perturbed by one bf16 ulp in 2 % of its elements.

Shims (test infrastructure, not reference code): peft and three transformers-4.36 symbols are absent from this image.
`PeftLoraLinear` below restates the peft 0.6 `lora.Linear` attribute layout the reference subclasses; the reference's
forward / from_float run unmodified on top of it.  `LlamaRotaryEmbedding436` restates transformers 4.36's rotary cache
(the reference's block calls `rotary_emb(x, seq_len=...)`).  Both are injected only while the reference files import.

    python tests/golden/gen_model_golden2.py        (container with /root/reference; never on the GPU box)
"""
import importlib
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import hosts  # noqa: E402
from gen_golden import bits16  # noqa: E402
from gen_model_golden import load_reference_quantize  # noqa: E402

ALL5 = "gemm,residual,layernorm,activation,scaling"


class PeftLoraLinear(nn.Linear):
    """Attribute layout of peft 0.6 `peft.tuners.lora.Linear` (what the reference's qat.LoraLinear expects)."""

    def __init__(self, adapter_name, in_features, out_features, r=0, lora_alpha=1, lora_dropout=0.0,
                 fan_in_fan_out=False, is_target_conv_1d_layer=False, **kwargs):
        kwargs.pop("init_lora_weights", None)
        super().__init__(in_features, out_features, **kwargs)
        self.fan_in_fan_out = fan_in_fan_out
        self.merged = False
        self.disable_adapters = False
        self.active_adapter = [adapter_name]
        self.r, self.lora_alpha, self.scaling = {adapter_name: r}, {adapter_name: lora_alpha}, {adapter_name: lora_alpha / r}
        self.lora_dropout = nn.ModuleDict({adapter_name: nn.Identity()})
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, out_features, bias=False)})
        self.weight.requires_grad_(False)
        if self.bias is not None:
            self.bias.requires_grad_(False)

    @property
    def active_adapters(self):
        return self.active_adapter

    def _linear(self, x):
        return F.linear(x, self.weight.T if self.fan_in_fan_out else self.weight, self.bias)

    def forward(self, x):   # float LoRA (not used after the swap)
        out = self._linear(x)
        for n in self.active_adapters:
            out = out + self.lora_B[n](self.lora_A[n](x)) * self.scaling[n]
        return out


class LlamaRotaryEmbedding436(nn.Module):
    """transformers 4.36 LlamaRotaryEmbedding: cos / sin cache, forward(x, seq_len) -> ([seq_len, dim], same)."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        cos, sin = hosts.rope_tables(dim, max_position_embeddings, float(base), torch.get_default_dtype())
        self.register_buffer("cos_cached", cos, persistent=False)
        self.register_buffer("sin_cached", sin, persistent=False)

    def forward(self, x, seq_len=None):
        return self.cos_cached[:seq_len].to(dtype=x.dtype), self.sin_cached[:seq_len].to(dtype=x.dtype)


def import_reference_llama():
    import transformers.models.llama.modeling_llama as hf
    saved = {n: getattr(hf, n, None) for n in ("LLAMA_ATTENTION_CLASSES", "LlamaRotaryEmbedding",
                                               "LlamaLinearScalingRotaryEmbedding", "LlamaDynamicNTKScalingRotaryEmbedding")}
    hf.LLAMA_ATTENTION_CLASSES = {}
    hf.LlamaRotaryEmbedding = LlamaRotaryEmbedding436
    hf.LlamaLinearScalingRotaryEmbedding = hf.LlamaDynamicNTKScalingRotaryEmbedding = LlamaRotaryEmbedding436
    try:
        mod = importlib.import_module("quantized_training.modules.quantizable.modeling_llama")
    finally:
        for n, v in saved.items():
            if v is None:
                delattr(hf, n)
            else:
                setattr(hf, n, v)
    return mod


def parse(ref, act, weight, fwd, bwd=None, error=None):
    argv = ["--activation", act, "--weight", weight, "--quantize_forward", fwd, "--bf16"]
    if bwd:
        argv += ["--quantize_backprop", bwd]
    args = ref.training_args.add_qspec_args().parse_args(argv)
    args.error = error      # as a STRING: the reference's CLI parses --error twice
    return args


def gen_mobilebert(ref, out):
    import types
    mm = sys.modules["quantized_training.modules.quantizable.modeling_mobilebert"]
    blocks = types.SimpleNamespace(MobileBertSelfAttention=mm.MobileBertSelfAttention,
                                   MobileBertSelfOutput=mm.MobileBertSelfOutput, FFNOutput=mm.FFNOutput,
                                   MobileBertOutput=mm.MobileBertOutput)
    cfg = hosts.mobilebert_config()
    g = torch.Generator().manual_seed(11)
    B, S = 3, 32
    x = torch.randn(B, S, cfg.hidden_size, generator=g)
    mask = torch.zeros(B, 1, 1, S)
    mask[2, ..., 24:] = -10000.0
    out["mobilebert/x"], out["mobilebert/mask"] = x.numpy(), mask.numpy()
    torch.manual_seed(5)
    proto = hosts.MobileBertHost(blocks, cfg)
    with torch.no_grad():   # NoNorm initialises to weight 1 / bias 0: give it something to do
        for n, p in proto.named_parameters():
            if "LayerNorm.weight" in n:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif "LayerNorm.bias" in n:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    out.update(hosts.state_to_numpy(proto, "mobilebert/w/"))
    for name, act, weight, fwd in [("mobilebert/e4m3_all", "e4m3", "e4m3", ALL5),
                                   ("mobilebert/e4m3_gemm", "e4m3", "e4m3", "gemm"),
                                   ("mobilebert/posit8_1_all", "posit8_1", "posit8_1", ALL5)]:
        model = hosts.MobileBertHost(blocks, cfg)
        model.load_state_dict(proto.state_dict())
        ref.quantize.quantize(model, parse(ref, act, weight, fwd))
        model.eval()
        with torch.no_grad():
            y = model(x.bfloat16(), mask.bfloat16())
        out[name + "/y"] = bits16(y)
        n_fq = sum(1 for m in model.modules() if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize))
        out[name + "/n_fq"] = np.array(n_fq)
        print(name, "fake-quant modules:", n_fq, "|y| mean", float(y.float().abs().mean()))


def gen_llama(ref, out):
    import types
    ml = import_reference_llama()
    blocks = types.SimpleNamespace(LlamaDecoderLayer=ml.LlamaDecoderLayer)
    cfg = hosts.llama_config()
    B, S = 2, 64
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(0, cfg.vocab_size, (B, S), generator=g)
    out["llama/ids"] = ids.numpy()
    mask = hosts.causal_mask(B, S)

    def call_layer(layer, x, m, pos):
        return layer(x, attention_mask=m, position_ids=pos)[0]

    torch.manual_seed(9)
    proto = hosts.LlamaHost(blocks, cfg, call_layer)
    with torch.no_grad():
        for n, p in proto.named_parameters():
            if "norm" in n:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
    out.update(hosts.state_to_numpy(proto, "llama/w/"))
    for name, act, weight, fwd in [("llama/posit8_1_gemm", "posit8_1", "posit8_1", "gemm"),
                                   ("llama/e4m3_gemm", "e4m3", "e4m3", "gemm"),
                                   ("llama/e4m3_all", "e4m3", "e4m3", ALL5),
                                   ("llama/posit8_1_all", "posit8_1", "posit8_1", ALL5)]:
        model = hosts.LlamaHost(blocks, cfg, call_layer)
        model.load_state_dict(proto.state_dict())
        ref.quantize.quantize(model, parse(ref, act, weight, fwd))
        model.eval()
        with torch.no_grad():
            logits = model(ids, mask)
        out[name + "/logits"] = bits16(logits)
        out[name + "/nll"] = np.array(float(hosts.nll(logits, ids)))
        n_fq = sum(1 for m in model.modules() if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize))
        out[name + "/n_fq"] = np.array(n_fq)
        print(name, "fake-quant modules:", n_fq, "nll", float(out[name + "/nll"]))
    # the un-quantized bf16 model, for scale: how far quantization itself moves the NLL
    model = hosts.LlamaHost(blocks, cfg, call_layer)
    model.load_state_dict(proto.state_dict())
    model.bfloat16().eval()
    with torch.no_grad():
        out["llama/bf16/nll"] = np.array(float(hosts.nll(model(ids, mask), ids)))
    print("llama bf16 nll", float(out["llama/bf16/nll"]))


def lora_inputs():
    g = torch.Generator().manual_seed(31)
    B, S, H = 4, 16, 64
    x = torch.randn(B, S, H, generator=g)
    mask = torch.zeros(B, 1, 1, S)
    mask[3, ..., 12:] = -10000.0
    labels = torch.tensor([0, 2, 1, 2])
    return x, mask, labels, g


def gen_lora(ref, out):
    import types
    mb = ref.blocks
    blocks = types.SimpleNamespace(BertSelfAttention=mb.BertSelfAttention, BertSelfOutput=mb.BertSelfOutput,
                                   BertOutput=mb.BertOutput)
    cfg = hosts.bert_config()
    x, mask, labels, g = lora_inputs()
    out["lora/x"], out["lora/mask"], out["lora/labels"] = x.numpy(), mask.numpy(), labels.numpy()

    def build():
        torch.manual_seed(3)
        model = hosts.BertHost(blocks, cfg)
        for p in model.parameters():
            p.requires_grad_(False)
        for layer in model.layers:     # what peft.get_peft_model does for target_modules = query, value (r = 8, alpha = 8)
            for name in ("query", "value"):
                lin = getattr(layer.attention, name)
                lo = PeftLoraLinear("default", lin.in_features, lin.out_features, r=8, lora_alpha=8, bias=True)
                lo.weight, lo.bias = lin.weight, lin.bias
                setattr(layer.attention, name, lo)
        for p in list(model.dense.parameters()) + list(model.out_proj.parameters()):   # modules_to_save: the head
            p.requires_grad_(True)
        return model

    proto = build()
    with torch.no_grad():   # peft initialises B to zero; give the adapters a non-trivial state
        for n, p in proto.named_parameters():
            if "lora_B" in n:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    out.update(hosts.state_to_numpy(proto, "lora/w/"))
    ERR = "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10"
    OPS = "gemm,residual,layernorm,activation"

    def run(xin, record):
        model = build()
        model.load_state_dict(proto.state_dict())
        ref.quantize.quantize(model, parse(ref, "fp8_e4m3", "fp8_e4m3", OPS, OPS, ERR))
        assert type(model.layers[0].attention.query).__module__.endswith("qat.lora")
        model.train()
        xb = xin.bfloat16().requires_grad_(True)
        trace = []
        handles = []
        for step in range(2):    # step 2 uses the delayed gradient scales of step 1
            for p in model.parameters():
                p.grad = None
            xb.grad = None
            if record and step == 1:
                for mn, m in model.named_modules():
                    if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize) and "error_" in mn:
                        def hook(mod, inp, outp, mn=mn):
                            trace.append((mn, bits16(inp[0].detach()), float(mod.scale), bits16(outp.detach())))
                        handles.append(m.register_forward_hook(hook))
            logits = model(xb, mask.bfloat16())
            loss = F.cross_entropy(logits.float(), labels)
            # scale saved BEFORE the call is what the call uses only if read before the observer update: record the
            # scale via a pre-hook instead
            loss.backward()
        for h in handles:
            h.remove()
        return model, xb, logits, loss, trace

    # scale in use: record with forward PRE-hooks (the observer updates `scale` inside forward, before quantizing)
    model, xb, logits, loss, trace = run(x, record=True)
    name = "lora/fp8_train"
    out[name + "/logits"] = bits16(logits.detach())
    out[name + "/loss"] = np.array(float(loss))
    out[name + "/gx"] = bits16(xb.grad)
    for pn, p in model.named_parameters():
        if p.grad is not None:
            out[f"{name}/grad/{pn}"] = bits16(p.grad) if p.grad.dtype == torch.bfloat16 else p.grad.float().numpy()
    for mn, m in model.named_modules():
        if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize) and "error_" in mn:
            out[f"{name}/scale/{mn}"] = m.scale.detach().float().reshape(-1).numpy().copy()
            out[f"{name}/hist/{mn}"] = m.amax_history.detach().float().reshape(-1).numpy().copy()
    # per-call replay records: the scale AFTER the call's own observer update is the one it quantized with
    for i, (mn, xin, scale_after, yout) in enumerate(trace):
        out[f"{name}/trace/{i:03d}/x"] = xin
        out[f"{name}/trace/{i:03d}/y"] = yout
        out[f"{name}/trace/{i:03d}/scale"] = np.array(scale_after, dtype=np.float32)
        out[f"{name}/trace/{i:03d}/name"] = np.array(mn)
    out[name + "/n_trace"] = np.array(len(trace))
    n_fq = sum(1 for m in model.modules() if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize))
    out[name + "/n_fq"] = np.array(n_fq)
    print(name, "fake-quant modules:", n_fq, "loss", float(loss), "trainable grads:",
          sum(1 for p in model.parameters() if p.grad is not None), "trace calls:", len(trace))

    # the reference's own sensitivity: perturb 2 % of the input elements by one bf16 ulp and re-run both steps
    xb16 = x.bfloat16()
    bits = xb16.view(torch.int16).clone()
    pick = torch.rand(bits.shape, generator=g) < 0.02
    bits[pick] += 1
    x2 = bits.view(torch.bfloat16).float()
    model2, xb2, logits2, loss2, _ = run(x2, record=False)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    sens = {"gx": rel(xb2.grad, xb.grad), "logits": rel(logits2, logits),
            "lora_A": rel(model2.layers[0].attention.query.lora_A["default"].weight.grad,
                          model.layers[0].attention.query.lora_A["default"].weight.grad),
            "out_proj": rel(model2.out_proj.weight.grad, model.out_proj.weight.grad)}
    for k, v in sens.items():
        out[f"{name}/ref_sensitivity/{k}"] = np.array(v)
    print("reference self-sensitivity to a 1-ulp perturbation of 2% of the input:", sens)


def main():
    ref = load_reference_quantize(lora_linear=PeftLoraLinear)
    out = {}
    gen_mobilebert(ref, out)
    gen_llama(ref, out)
    gen_lora(ref, out)
    path = os.path.join(HERE, "model_cases2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
