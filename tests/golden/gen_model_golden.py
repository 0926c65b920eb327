#!/usr/bin/env python
"""Model-level golden fixtures: the UNMODIFIED reference `quantize(model, args)` (quantize.py:52-101) applied, on CPU,
to a small HF-independent encoder assembled from the reference's OWN quantizable blocks
(modules/quantizable/modeling_bert.py: BertSelfAttention, BertSelfOutput, BertOutput) + nn.Linear / nn.GELU, then one
forward (and, for the training case, one backward).  Stored: the fp32 initial weights, the inputs, and the
reference's outputs / input gradients as bf16 bit patterns.  tests/test_model_golden_gpu.py rebuilds the same encoder
from THIS repo's blocks, loads the weights, calls this repo's quantize() with the same flags and compares.

What this pins: which tensors are fake-quantized for every --quantize_forward / --quantize_backprop
op group, in which order, weight re-quantization, the STE backward and the gradient hooks -- at model level, against
the reference itself rather than a restatement.

    python tests/golden/gen_model_golden.py        (container with /root/reference; never on the GPU box)
"""
import importlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import REF, bits16, load_reference  # noqa: E402

HID, HEADS, INTER, LAYERS, B, S = 64, 4, 128, 2, 3, 32
CASES = [
    # name, activation, weight, error, forward ops, backprop ops
    ("posit8_1_gemm", "posit8_1", "posit8_1", None, "gemm", None),
    ("posit8_1_all", "posit8_1", "posit8_1", None, "gemm,residual,layernorm,activation,scaling", None),
    ("e4m3_gemm", "e4m3", "e4m3", None, "gemm", None),
    ("e4m3_gemm_layernorm", "e4m3", "e4m3", None, "gemm,layernorm", None),
    ("int8_dyn_gemm", "int8,qs=per_tensor_symmetric", "int8,qs=per_channel_symmetric,ax=0", None, "gemm", None),
    ("fp8_train", "fp8_e4m3", "fp8_e4m3", "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10",
     "gemm,residual,layernorm,activation", "gemm,residual,layernorm,activation"),
]


def load_reference_quantize(lora_linear=None):
    """lora_linear: a functional stand-in for peft.tuners.lora.Linear (peft is absent from this image) that the
    reference's qat.LoraLinear subclasses; None keeps the inert placeholder."""
    ref = load_reference()
    if lora_linear is not None:
        sys.modules["peft.tuners.lora"].Linear = lora_linear
    pkg = sys.modules["quantized_training"]

    def synth(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    mods = synth("quantized_training.modules", os.path.join(REF, "modules"))
    pkg.modules = mods
    mods.Softmax = importlib.import_module("quantized_training.modules.softmax").Softmax
    for dummy in ("modeling_bert", "modeling_mobilebert"):       # legacy full-model copies: not on this path
        d = types.ModuleType(f"quantized_training.modules.{dummy}")
        sys.modules[d.__name__] = d
        setattr(mods, dummy, d)
    qz = synth("quantized_training.modules.quantizable", os.path.join(REF, "modules", "quantizable"))
    mods.quantizable = qz
    fm = importlib.import_module("quantized_training.modules.quantizable.functional_modules")
    mb = importlib.import_module("quantized_training.modules.quantizable.modeling_bert")
    mm = importlib.import_module("quantized_training.modules.quantizable.modeling_mobilebert")
    for n in ("AddFunctional", "MulFunctional", "MatmulFunctional"):
        setattr(qz, n, getattr(fm, n))
    for n in ("BertSelfAttention", "BertSelfOutput", "BertOutput"):
        setattr(qz, n, getattr(mb, n))
    for n in ("MobileBertSelfAttention", "MobileBertSelfOutput", "MobileBertOutput", "FFNOutput"):
        setattr(qz, n, getattr(mm, n))
    for n in ("TransformerBlock", "GPT2Block", "WhisperEncoderLayer", "WhisperDecoderLayer", "LlamaDecoderLayer"):
        setattr(qz, n, type(n, (nn.Module,), {}))                # other model families: placeholders
    mods.qat = importlib.import_module("quantized_training.modules.qat")
    ref.quantize = importlib.import_module("quantized_training.quantize")
    ref.training_args = importlib.import_module("quantized_training.training_args")
    ref.blocks = mb
    return ref


def config():
    return SimpleNamespace(hidden_size=HID, num_attention_heads=HEADS, intermediate_size=INTER,
                           attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0, layer_norm_eps=1e-12,
                           is_decoder=False, position_embedding_type="absolute", max_position_embeddings=64)


def build_host(blocks, cfg):
    """The encoder; `blocks` supplies BertSelfAttention / BertSelfOutput / BertOutput (reference's or this repo's)."""

    class Layer(nn.Module):
        def __init__(self):
            super().__init__()
            self.attention = blocks.BertSelfAttention(cfg)
            self.attn_out = blocks.BertSelfOutput(cfg)
            self.inter = nn.Linear(HID, INTER)
            self.act = nn.GELU()
            self.out = blocks.BertOutput(cfg)

        def forward(self, x, mask):
            a = self.attention(x, mask)[0]
            a = self.attn_out(a, x)
            return self.out(self.act(self.inter(a)), a)

    class Host(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList([Layer() for _ in range(LAYERS)])
            self.head = nn.Linear(HID, 2)

        def forward(self, x, mask):
            for layer in self.layers:
                x = layer(x, mask)
            return self.head(x)

    return Host()


def inputs():
    """Hidden states and an additive padding mask.  The fill value is -10000 (HF 3.x style), not finfo.min: with the
    `activation` group hooked the mask reaches a fake-quantizer (softmax input), and -3.39e38 lies in the band the
    reference's fpN_eXmY formats map to NaN (fp8.py:147-203 evaluated in bf16) -- a genuine reference behaviour that
    would turn the whole training case into NaNs and pin nothing."""
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, S, HID, generator=g)
    mask = torch.zeros(B, 1, 1, S)
    mask[1, ..., 20:] = -10000.0
    return x, mask


def main():
    ref = load_reference_quantize()
    out = {}
    x, mask = inputs()
    out["x"], out["mask"] = x.numpy(), mask.numpy()
    torch.manual_seed(7)
    proto = build_host(ref.blocks, config())
    for k, v in proto.state_dict().items():
        out["w/" + k] = v.numpy().copy()
    for name, act, weight, error, fwd, bwd in CASES:
        model = build_host(ref.blocks, config())
        model.load_state_dict(proto.state_dict())
        argv = ["--activation", act, "--weight", weight, "--quantize_forward", fwd, "--bf16"]
        if bwd:
            argv += ["--quantize_backprop", bwd]
        args = ref.training_args.add_qspec_args().parse_args(argv)
        args.error = error      # as a STRING: the reference's CLI parses --error twice
        ref.quantize.quantize(model, args)
        xb = x.bfloat16().requires_grad_(bwd is not None)
        mb = mask.bfloat16()
        if bwd:
            model.train()
            # two steps so that the delayed gradient scales (ahl = 10) are in use on the recorded one
            for step in range(2):
                xb.grad = None
                y = model(xb, mb)
                y.float().square().sum().backward()
            out[f"{name}/y"] = bits16(y.detach())
            out[f"{name}/gx"] = bits16(xb.grad)
            gw = model.head.weight.grad
            out[f"{name}/g_head"] = bits16(gw)
            # gradient fake-quantizer state after the two steps (delayed scaling: scale in use + amax history)
            for mn, m in model.named_modules():
                if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize) and "error_" in mn:
                    out[f"{name}/scale/{mn}"] = m.scale.detach().float().reshape(-1).numpy().copy()
                    out[f"{name}/hist/{mn}"] = m.amax_history.detach().float().reshape(-1).numpy().copy()
        else:
            model.eval()
            with torch.no_grad():
                for _ in range(2):                      # second call: delayed scales of the dynamic case are in use
                    y = model(xb, mb)
            out[f"{name}/y"] = bits16(y)
        n_fq = sum(1 for m in model.modules() if isinstance(m, ref.fq.FusedAmaxObsFakeQuantize))
        out[f"{name}/n_fq"] = np.array(n_fq)
        print(name, "fake-quant modules:", n_fq, "|y| mean", float(y.float().abs().mean()))
    np.savez_compressed(os.path.join(HERE, "model_cases.npz"), **out)
    print("wrote model_cases.npz", os.path.getsize(os.path.join(HERE, "model_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
