"""Model-level parity against the REFERENCE ITSELF, second set: MobileBERT (BASELINE configs[0]), Llama decoder
(configs[4], with a stated NLL delta) and the LoRA fine-tune step (configs[3]).  tests/golden/model_cases2.npz holds
the outputs of the unmodified reference `quantize(model, args)` run on CPU over host models built from the reference's
own quantizable blocks (generator: tests/golden/gen_model_golden2.py; hosts: tests/golden/hosts.py).  Here the same
hosts are built from this repo's blocks with the same weights, quantized with the same flags, and run on the kernels.

Tolerances (floating point, stated):
* forward outputs: relative Frobenius error <= 1 %, and >= 95 % of the elements bit-identical (CPU bf16 GEMM vs
  tcgen05: fp32 accumulation in a different order, one rounding; a flipped ulp can flip an 8-bit code downstream);
* Llama NLL (north star "end-to-end logits within a stated perplexity delta"): |NLL - NLL_ref| <= 1e-4 nats, i.e.
  perplexity within 0.01 % of the reference's run (measured on B200: logits bit-identical, |delta| <= 5e-7) -- for
  scale, quantization itself moves the NLL by 0.0002-0.006 nats on this model (golden `llama/bf16/nll`);
* fine-tune step: the chain forward -> loss -> E5M2 delayed-scaling backward is CHAOTIC at the percent level in the
  reference itself: re-running the reference with 2 % of the input elements moved by one bf16 ulp changes its own
  input gradient by 10.6 %, its LoRA-factor gradient by 21.6 % and its logits by 4.6 % (golden `ref_sensitivity/*`,
  produced by the generator).  The bars for this build are those self-sensitivities x 1.5.  What IS exact is every
  single step: each of the 38 gradient fake-quant calls of the reference's backward is replayed on its recorded
  input and scale and must reproduce the recorded output bit for bit (test_lora_backward_quantizers_replay_exactly),
  and dgrad / wgrad are checked against fp64 in tests/test_gemm_gpu.py.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import quantized_training as qt
from quantized_training import fused
from quantized_training.modules import quantizable as blocks
from quantized_training.modules.lora import LoraLinear

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import hosts  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_cases2.npz")
ALL5 = "gemm,residual,layernorm,activation,scaling"


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


def from_bits(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16).copy()).view(torch.bfloat16)


def weights(G, prefix):
    return {k[len(prefix):]: torch.from_numpy(G[k]) for k in G.files if k.startswith(prefix)}


def compare(got, want_bits, tol, min_identical, what=""):
    want = from_bits(want_bits).to(got.device).view(got.shape)
    g, w = got.double(), want.double()
    rel = float((g - w).norm() / w.norm())
    identical = float((got.contiguous().view(torch.int16) == want.contiguous().view(torch.int16)).float().mean())
    print(f"[golden2] {what}: rel {rel:.5f}, bit-identical {identical:.4f}")
    assert rel <= tol, f"{what}: relative error {rel:.4f} > {tol} (identical fraction {identical:.4f})"
    assert identical >= min_identical, f"{what}: only {identical:.4f} bit-identical (rel {rel:.5f})"
    return rel, identical


def parse(act, weight, fwd, bwd=None, error=None):
    argv = ["--activation", act, "--weight", weight, "--quantize_forward", fwd, "--bf16"]
    if bwd:
        argv += ["--quantize_backprop", bwd, "--error", error]
    return qt.add_qspec_args().parse_args(argv)


# ----------------------------------------------------------------------------------------------- MobileBERT
@pytest.mark.parametrize("case,act,fwd", [("e4m3_all", "e4m3", ALL5), ("e4m3_gemm", "e4m3", "gemm"),
                                          ("posit8_1_all", "posit8_1", ALL5)])
def test_mobilebert_encoder_matches_the_reference_run(G, case, act, fwd):
    cfg = hosts.mobilebert_config()
    model = hosts.MobileBertHost(blocks, cfg)
    model.load_state_dict(weights(G, "mobilebert/w/"), strict=True)
    model.to(DEV)
    qt.quantize(model, parse(act, act, fwd))
    model.eval()
    x = torch.from_numpy(G["mobilebert/x"]).to(DEV).bfloat16()
    mask = torch.from_numpy(G["mobilebert/mask"]).to(DEV).bfloat16()
    with torch.no_grad():
        y = model(x, mask)
    compare(y, G[f"mobilebert/{case}/y"], 1e-2, 0.95, f"mobilebert/{case}")
    n_fq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    assert n_fq == int(G[f"mobilebert/{case}/n_fq"])


# ----------------------------------------------------------------------------------------------- Llama
def llama_call_layer(cos, sin):
    def call(layer, x, mask, position_ids):
        b = x.shape[0]
        return layer(x, attention_mask=mask, position_ids=position_ids,
                     position_embeddings=(cos[None].expand(b, -1, -1).contiguous(), sin[None].expand(b, -1, -1).contiguous()))
    return call


@pytest.mark.parametrize("fused_on", [True, False])
@pytest.mark.parametrize("case,act,fwd", [("posit8_1_gemm", "posit8_1", "gemm"), ("e4m3_gemm", "e4m3", "gemm"),
                                          ("e4m3_all", "e4m3", ALL5), ("posit8_1_all", "posit8_1", ALL5)])
def test_llama_decoder_logits_and_nll_match_the_reference_run(G, case, act, fwd, fused_on):
    cfg = hosts.llama_config()
    ids = torch.from_numpy(G["llama/ids"]).to(DEV)
    B, S = ids.shape
    cos, sin = hosts.rope_tables(cfg.hidden_size // cfg.num_attention_heads, S, 10000.0, torch.bfloat16, DEV)
    model = hosts.LlamaHost(blocks, cfg, llama_call_layer(cos, sin))
    model.load_state_dict(weights(G, "llama/w/"), strict=True)
    model.to(DEV)
    qt.quantize(model, parse(act, act, fwd))
    model.eval()
    mask = hosts.causal_mask(B, S, torch.bfloat16, DEV)
    fused.set_enabled(fused_on)
    try:
        with torch.no_grad():
            logits = model(ids, mask)
    finally:
        fused.set_enabled(True)
    compare(logits, G[f"llama/{case}/logits"], 1e-2, 0.99, f"llama/{case} fused={fused_on}")   # measured: 1.0000
    nll, want = float(hosts.nll(logits, ids)), float(G[f"llama/{case}/nll"])
    print(f"[golden2] llama/{case} fused={fused_on}: NLL {nll:.6f} reference {want:.6f} delta {nll - want:+.2e} "
          f"(bf16 un-quantized {float(G['llama/bf16/nll']):.6f})")
    assert abs(nll - want) <= 1e-4, f"NLL {nll} vs reference {want}"   # measured: <= 5e-7
    n_fq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    assert n_fq == int(G[f"llama/{case}/n_fq"])


# ----------------------------------------------------------------------------------------------- LoRA fine-tune step
ERR = "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10"
OPS = "gemm,residual,layernorm,activation"


def build_lora_model(G):
    cfg = hosts.bert_config_hf()
    model = hosts.BertHost(blocks, cfg)
    for p in model.parameters():
        p.requires_grad_(False)
    for layer in model.layers:
        for name in ("query", "value"):
            setattr(layer.attention, name, LoraLinear.from_linear(getattr(layer.attention, name), r=8, lora_alpha=8))
    for p in list(model.dense.parameters()) + list(model.out_proj.parameters()):
        p.requires_grad_(True)
    model.load_state_dict(weights(G, "lora/w/"), strict=True)
    model.to(DEV)
    qt.quantize(model, parse("fp8_e4m3", "fp8_e4m3", OPS, OPS, ERR))
    assert isinstance(model.layers[0].attention.query, qt.modules.qat.LoraLinear)
    return model


def test_lora_finetune_step_matches_the_reference_run(G):
    name = "lora/fp8_train"
    model = build_lora_model(G)
    model.train()
    x = torch.from_numpy(G["lora/x"]).to(DEV).bfloat16().requires_grad_(True)
    mask = torch.from_numpy(G["lora/mask"]).to(DEV).bfloat16()
    labels = torch.from_numpy(G["lora/labels"]).to(DEV)
    for _ in range(2):
        for p in model.parameters():
            p.grad = None
        x.grad = None
        logits = model(x, mask)
        loss = F.cross_entropy(logits.float(), labels)
        loss.backward()
    sens = {k: float(G[f"{name}/ref_sensitivity/{k}"]) for k in ("gx", "logits", "lora_A", "out_proj")}
    compare(logits.detach(), G[f"{name}/logits"], 1.5 * sens["logits"], 0.0, "lora logits")
    assert abs(float(loss) - float(G[f"{name}/loss"])) <= 0.05 * abs(float(G[f"{name}/loss"]))
    # the same trainable set, gradients within the reference's own sensitivity x 1.5
    ref_grads = sorted(k[len(name) + 6:] for k in G.files if k.startswith(name + "/grad/"))
    ours = {n: p for n, p in model.named_parameters() if p.grad is not None}
    assert sorted(ours) == ref_grads, (sorted(ours), ref_grads)
    for n in ref_grads:
        bar = 1.5 * (sens["lora_A"] if "lora_" in n else sens["out_proj"])
        if n.endswith(".bias"):
            bar = max(bar, 0.15)
        compare(ours[n].grad, G[f"{name}/grad/{n}"], bar, 0.0, f"grad {n}")
    compare(x.grad, G[f"{name}/gx"], 1.5 * sens["gx"], 0.0, "lora gx")
    # same gradient fake-quantizers at the same hook points, delayed-scaling state tracking the reference's
    fqs = {n: m for n, m in model.named_modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize) and "error_" in n}
    ref_names = sorted(k[len(name) + 7:] for k in G.files if k.startswith(name + "/scale/"))
    assert sorted(fqs) == ref_names
    for n in ref_names:
        # a scale is ONE element (the previous step's largest |gradient|, itself an E5M2-quantized value upstream):
        # neighbouring E5M2 values are 14-25 % apart, so a single flipped code moves it by that much
        np.testing.assert_allclose(fqs[n].scale.detach().float().reshape(-1).cpu().numpy(), G[f"{name}/scale/{n}"], rtol=0.30)
    n_fq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    assert n_fq == int(G[f"{name}/n_fq"])


def test_lora_backward_quantizers_replay_exactly(G):
    """Every gradient fake-quant call of the reference's recorded backward, replayed in isolation on the reference's
    own input with the scale it used: bit-exact.  Together with the fp64 checks of dgrad / wgrad this isolates the
    percent-level end-to-end differences above to rounding-order noise amplified by the 2-bit-mantissa E5M2 grid."""
    name = "lora/fp8_train"
    n = int(G[f"{name}/n_trace"])
    assert n >= 30
    spec = qt.QuantizationSpec.from_str(ERR)
    for i in range(n):
        xin = from_bits(G[f"{name}/trace/{i:03d}/x"]).to(DEV)
        want = G[f"{name}/trace/{i:03d}/y"]
        m = qt.FusedAmaxObsFakeQuantize(**spec.fake_quant_kwargs(), device=DEV)
        m.disable_observer()
        m.scale.fill_(float(G[f"{name}/trace/{i:03d}/scale"]))
        got = m(xin).cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
        w = np.asarray(want).astype(np.uint16)
        same = (got == w) | (((got & 0x7FFF) > 0x7F80) & ((w & 0x7FFF) > 0x7F80))
        assert same.all(), f"trace call {i} ({str(G[f'{name}/trace/{i:03d}/name'])}): {int((~same).sum())} mismatches"


def test_qat_lora_linear_forward_backward_against_reference_formula(G):
    """qat.LoraLinear in isolation on the golden weights of layers.0.attention.query (modules/qat/lora.py:34-55):
    W' = fq(W + (fq(B) @ fq(A)) * scaling), y = x W'^T + b; gradients of A and B through the STE."""
    W = weights(G, "lora/w/")
    pre = "layers.0.attention.query."
    lin = torch.nn.Linear(64, 64)
    lo = LoraLinear.from_linear(lin, r=8, lora_alpha=8)
    lo.load_state_dict({k[len(pre):]: v for k, v in W.items() if k.startswith(pre)})
    lo.to(DEV).bfloat16()
    lo.qconfig = qt.get_qconfig("fp8_e4m3", "fp8_e4m3", None)
    q = qt.modules.qat.LoraLinear.from_float(lo)
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(32, 64, generator=gen).to(DEV).bfloat16()
    g = torch.randn(32, 64, generator=gen).to(DEV).bfloat16()
    y = q(x)
    y.backward(g)
    # the reference's op chain in torch, with this repo's bit-exact fake-quant kernel as `fq`
    fq = qt.FusedAmaxObsFakeQuantize("fp8_e4m3", device=DEV)
    A = q.lora_A["default"].weight.detach().clone().requires_grad_(True)
    Bm = q.lora_B["default"].weight.detach().clone().requires_grad_(True)
    ste = lambda t: t + (fq(t.detach()) - t.detach())
    merged = q.weight.data.clone() + (ste(Bm) @ ste(A)) * q.scaling["default"]
    yr = F.linear(x, ste(merged), q.bias)
    yr.backward(g)
    assert torch.equal(y.detach(), yr.detach()) or float((y.detach().double() - yr.detach().double()).norm() / yr.detach().double().norm()) < 4e-3
    for ours, ref in ((q.lora_A["default"].weight.grad, A.grad), (q.lora_B["default"].weight.grad, Bm.grad)):
        rel = float((ours.double() - ref.double()).norm() / ref.double().norm())
        assert rel < 1e-2, rel


@pytest.mark.parametrize("spec", ["fp8_e4m3", "posit8_1", "int8,qs=per_tensor_symmetric,ahl=4", "e4m3"])
@pytest.mark.parametrize("N,K,r", [(768, 768, 8), (64, 72, 4), (256, 1024, 16)])
def test_lora_merge_kernel_is_the_reference_op_chain(spec, N, K, r):
    """qt_lora_merge_fq vs clone -> fq(A) -> fq(B) -> B @ A -> * scaling -> += -> fq in torch bf16 ops (the reference's
    chain, modules/qat/lora.py:44-52) with this repo's bit-exact fake-quant as `fq`: BIT-EXACT (r <= 16 products of
    8-bit values are exact in fp32, so the accumulation order cannot matter), inference and observed-training modes."""
    gen = torch.Generator().manual_seed(N + K + r)
    w = (torch.randn(N, K, generator=gen) * 0.05).to(DEV).bfloat16()
    a = (torch.randn(r, K, generator=gen) * 0.1).to(DEV).bfloat16()
    b = (torch.randn(N, r, generator=gen) * 0.1).to(DEV).bfloat16()
    scaling = 8 / r
    qs = qt.QuantizationSpec.from_str(spec)
    for mode in ("infer", "train"):
        f1 = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), device=DEV)
        f2 = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), device=DEV)
        lin = torch.nn.Linear(K, N, bias=False)
        lo = LoraLinear.from_linear(lin, r=r, lora_alpha=8).to(DEV).bfloat16()
        lo.weight.data.copy_(w); lo.lora_A["default"].weight.data.copy_(a); lo.lora_B["default"].weight.data.copy_(b)
        lo.qconfig = qt.get_qconfig(spec, spec, None)
        q = qt.modules.qat.LoraLinear.from_float(lo)
        q.weight_fake_quant = f1
        x = torch.eye(K, device=DEV, dtype=torch.bfloat16)      # y = W'^T: reads the merged weight back exactly
        for call in range(2):
            if mode == "infer":
                with torch.no_grad():
                    y = q(x)
            else:
                y = q(x)
            merged = w.clone()
            aq = f2(a)                      # the reference quantizes A first, then B (lora.py:47-48)
            bq = f2(b)
            merged += (bq @ aq) * scaling
            want = f2(merged)
            assert torch.equal(y.detach().t().contiguous(), want), (spec, mode, call)
        assert torch.equal(f1.scale, f2.scale) and torch.equal(f1.amax_history, f2.amax_history)
