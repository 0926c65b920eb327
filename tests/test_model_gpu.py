"""GPU tests of quantize()-d models: the hooked blocks run on the kernels and agree with a plain-PyTorch
restatement of the same fake-quant placement (oracle tables applied with torch ops, cuBLAS matmuls).

Tolerance: the GEMMs accumulate in fp32 and round to bf16 once, like cuBLAS, but in a different summation
order, and every later fake-quant step can flip a rounding on such a difference.  So model outputs are compared
statistically: relative Frobenius error <= 2 % per quantized forward (measured ~0.3 %), and the structural
facts (which kernels ran, buffer names, STE gradients) exactly."""
import numpy as np
import pytest
import torch
from torch import nn
from transformers import BertConfig, BertForQuestionAnswering, LlamaConfig, LlamaForCausalLM

import quantized_training as qt
from quantized_training import _C, fused, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def parse(*argv):
    return qt.add_qspec_args().parse_args(list(argv))


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def table_fq(oracle, dtype):
    """fake-quant as a table lookup in torch (bare spec), the reference's own formulation."""
    table = torch.from_numpy(oracle.qmap(dtype).view(np.int16)).view(torch.bfloat16).to(DEV)
    return lambda t: table[(t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF)].view(t.shape)


def test_qat_linear_forward_backward(oracle):
    torch.manual_seed(0)
    lin = nn.Linear(256, 512).to(DEV).bfloat16()
    model = nn.Sequential(lin)
    qt.quantize(model, parse("--activation", "posit8_1", "--weight", "posit8_1", "--error", "posit8_1",
                             "--quantize_backprop", "gemm", "--bf16"))
    x = torch.randn(4, 96, 256, device=DEV).bfloat16().requires_grad_()
    y = model(x)
    fq = table_fq(oracle, "posit8_1")
    xq, wq = fq(x.detach()), fq(lin.weight.detach())
    ref = (xq.double() @ wq.double().t() + lin.bias.double())
    assert rel_err(y, ref) < 2e-3
    g = torch.randn_like(y)
    y.backward(g)
    gq = fq(g)                                            # error_pre_process quantizes grad_output
    assert rel_err(x.grad, gq.double() @ wq.double()) < 5e-3   # dgrad uses the quantized weight (STE through fq)
    qlin = model[0]
    assert rel_err(qlin.weight.grad, gq.reshape(-1, 512).double().t() @ xq.reshape(-1, 256).double()) < 5e-3
    assert "activation_pre_process" in dict(qlin.named_children()) and "0" in qlin.activation_pre_process
    assert "0" in qlin.error_pre_process


@pytest.mark.parametrize("ops_str", ["gemm", "gemm,residual,layernorm,activation,scaling"])
def test_bert_block_matches_torch_restatement(oracle, ops_str):
    torch.manual_seed(1)
    cfg = BertConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     vocab_size=500, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = BertForQuestionAnswering(cfg).to(DEV).eval()
    qt.quantize(model, parse("--activation", "posit8_1", "--weight", "posit8_1", "--quantize_forward", ops_str,
                             "--bf16", "--op_fusion", "qa_outputs"))
    ids = torch.randint(0, 500, (2, 64), device=DEV)
    with torch.no_grad():
        got = model(input_ids=ids)
        ops.set_enabled(False)          # same hooks and fake-quant kernels, cuBLAS matmuls (the reference's K5/K6)
        want = model(input_ids=ids)
        ops.set_enabled(True)
    assert rel_err(got.start_logits, want.start_logits) < 2e-2
    assert rel_err(got.end_logits, want.end_logits) < 2e-2
    names = [n for n, m in model.named_modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize)]
    assert any(n.endswith("qk_matmul.activation_pre_process.1") for n in names)
    kt = model.bert.encoder.layer[0].attention.self.qk_matmul.activation_pre_process["1"]
    assert kt.preserve_strides


def test_llama_tiny_forward_and_graph_capture():
    torch.manual_seed(2)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, attn_implementation="eager")
    model = LlamaForCausalLM(cfg).to(DEV).eval()
    qt.quantize(model, parse("--activation", "e4m3", "--weight", "e4m3", "--quantize_forward", "gemm", "--bf16"))
    ids = torch.randint(0, 512, (1, 128), device=DEV)
    with torch.no_grad():
        got = model(input_ids=ids, use_cache=False).logits
        ops.set_enabled(False)
        want = model(input_ids=ids, use_cache=False).logits
        ops.set_enabled(True)
    assert got.shape == (1, 128, 512) and rel_err(got, want) < 2e-2


def test_fp8_linear_route_matches_bf16_route():
    torch.manual_seed(4)
    model = nn.Sequential(nn.Linear(512, 768)).to(DEV)
    qt.quantize(model, parse("--activation", "e4m3", "--weight", "e4m3", "--bf16"))
    x = (torch.randn(3, 200, 512, device=DEV) * 2).bfloat16().requires_grad_()
    lin = model[0]
    y8 = model(x)                                   # bare e4m3 on both sides -> FP8 tensor cores
    assert ops.fp8_route(lin, lin.activation_pre_process["0"](x), lin.weight_fake_quant) == "e4m3"
    y8.sum().backward()
    g8 = x.grad.clone(); x.grad = None
    xq = lin.activation_pre_process["0"](x.detach())
    y16 = ops.linear(xq, lin.weight_fake_quant(lin.weight), lin.bias)
    assert rel_err(y8, y16) < 1e-3                  # same products, fp32 accumulation in both
    ref_gx = torch.ones_like(y16).reshape(-1, 768) @ lin.weight_fake_quant(lin.weight)
    assert rel_err(g8.reshape(-1, 512), ref_gx) < 1e-2


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3"])
@pytest.mark.parametrize("ops_str", ["gemm", "gemm,residual,layernorm,activation,scaling", "gemm,scaling,activation"])
def test_llama_fused_layer_matches_module_by_module(spec, ops_str, monkeypatch):
    """The fused block (fused.py, ~13 launches per layer) against the same model executed module by module through
    the hooks (the reference's structure) -- same kernels for every fake-quant step, same rounding points.
    Tolerance: relative Frobenius error of the logits <= 2 % (fp32 epilogue vs bf16 intermediate for the un-hooked
    residual adds, different reduction orders); the fused path must actually have run."""
    torch.manual_seed(5)
    cfg = LlamaConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, attn_implementation="eager")
    model = LlamaForCausalLM(cfg).to(DEV).eval()
    qt.quantize(model, parse("--activation", spec, "--weight", spec, "--quantize_forward", ops_str, "--bf16"))
    ids = torch.randint(0, 512, (2, 96), device=DEV)
    calls = {"n": 0}
    real = _C.norm_fq

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    monkeypatch.setattr(_C, "norm_fq", counting)
    with torch.no_grad():
        model(input_ids=ids, use_cache=False)           # first call creates the lazy fake-quantizers (module path)
        assert calls["n"] == 0
        got = model(input_ids=ids, use_cache=False).logits
        assert calls["n"] == 4                          # 2 layers x 2 norms went through the fused kernels
        fused.set_enabled(False)
        want = model(input_ids=ids, use_cache=False).logits
        fused.set_enabled(True)
    assert rel_err(got, want) < 2e-2
    if spec == "e4m3":     # every product of the fused layer took fp8 codes (FP8 tensor cores)
        cache = model.model.layers[0].__dict__["_qt_wcache"]
        assert all(v[1].dtype == torch.uint8 for v in cache.values())
    x = torch.randn(2, 96, 256, device=DEV).bfloat16().requires_grad_()
    assert fused.llama_layer_forward(model.model.layers[0], x, None, None) is None     # autograd on: module path


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3"])
@pytest.mark.parametrize("ops_str", ["gemm", "gemm,residual,layernorm,activation,scaling", "gemm,layernorm"])
def test_bert_fused_layer_matches_module_by_module(spec, ops_str, monkeypatch):
    """BERT encoder layer: fused execution (fused.bert_layer_forward) against the hooked module-by-module execution of
    the same quantize()-d model, with a padding mask.  Same tolerance statement as the Llama test."""
    torch.manual_seed(7)
    cfg = BertConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     vocab_size=500, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = BertForQuestionAnswering(cfg).to(DEV).eval()
    qt.quantize(model, parse("--activation", spec, "--weight", spec, "--quantize_forward", ops_str,
                             "--bf16", "--op_fusion", "qa_outputs"))
    ids = torch.randint(0, 500, (3, 64), device=DEV)
    am = torch.ones(3, 64, device=DEV, dtype=torch.long)
    am[1, 40:] = 0
    calls = {"n": 0}
    real, real_add = _C.norm_fq, _C.add_norm_fq

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    def counting_add(*a, **k):            # hooked residual add + LayerNorm + hooks in one pass (residual group on)
        calls["n"] += 1
        return real_add(*a, **k)

    monkeypatch.setattr(_C, "norm_fq", counting)
    monkeypatch.setattr(_C, "add_norm_fq", counting_add)
    with torch.no_grad():
        model(input_ids=ids, attention_mask=am)
        assert calls["n"] == 0
        got = model(input_ids=ids, attention_mask=am)
        assert calls["n"] == 4                          # 2 layers x 2 LayerNorms through the fused kernels
        fused.set_enabled(False)
        want = model(input_ids=ids, attention_mask=am)
        fused.set_enabled(True)
    assert rel_err(got.start_logits, want.start_logits) < 2e-2
    assert rel_err(got.end_logits, want.end_logits) < 2e-2
    same = float((got.start_logits.view(torch.int16) == want.start_logits.view(torch.int16)).float().mean())
    assert same >= 0.95, same


def test_quantized_weight_cache_follows_recalibration():
    """The eval-time cache of fq(W) must not outlive the scale it was computed with: calibrate -> freeze -> eval ->
    calibrate again (the kernel writes `scale` through a raw pointer, invisible to torch's version counter) -> freeze
    -> eval has to re-quantize with the NEW scale, as the reference does on every forward (modules/qat/linear.py:41)."""
    torch.manual_seed(0)
    lin = nn.Linear(64, 32).to(DEV).bfloat16()
    lin.qconfig = qt.get_qconfig(None, "int8,qs=per_tensor_symmetric,ahl=1", None)
    q = qt.modules.qat.Linear.from_float(lin)
    fq = q.weight_fake_quant
    x = torch.randn(8, 64, device=DEV).bfloat16()

    def frozen_eval():
        fq.disable_observer()
        with torch.no_grad():
            y = q(x)
            want = torch.nn.functional.linear(x, fq(q.weight), q.bias)
        return y, want

    with torch.no_grad():
        q(x); q(x)                          # calibration: scale <- amax(W) / 127
    y1, w1 = frozen_eval()
    assert torch.equal(y1, w1)
    s1 = float(fq.scale)
    fq.enable_observer()
    with torch.no_grad():
        q.weight.mul_(3.0)                  # in-place: version counter moves, cache invalid
        q(x); q(x)                          # re-calibration on the new weights: the scale moves
    y2, w2 = frozen_eval()
    assert float(fq.scale) != s1 and torch.equal(y2, w2)
    # scale rewritten by a further observed pass WITHOUT touching the weight: only the epoch tells
    fq.enable_observer()
    fq.amax_history.fill_(50.0)             # pretend an outlier batch went through
    with torch.no_grad():
        q(x)
    y3, w3 = frozen_eval()
    assert torch.equal(y3, w3) and not torch.equal(y3, y2)


@pytest.mark.parametrize("spec", ["e4m3", "posit8_1"])
@pytest.mark.parametrize("ops_str", ["gemm,residual,layernorm,activation,scaling", "gemm", "gemm,layernorm"])
def test_mobilebert_fused_layer_matches_module_by_module(spec, ops_str, monkeypatch):
    """BASELINE configs[0] host model: the fused MobileBERT layer (fused.mobilebert_layer_forward) against the same
    model executed module by module through the hooks (the reference's structure, itself pinned to the reference's own
    run by tests/test_model_golden2_gpu.py).  Same weights, same fake-quantizers, same rounding points: the logits
    must agree to <= 1 % relative Frobenius error with >= 98 % of the elements bit-identical, and the fused path must
    actually have run."""
    from transformers import MobileBertConfig, MobileBertForQuestionAnswering
    from quantized_training import fused
    torch.manual_seed(6)
    cfg = MobileBertConfig(vocab_size=600, hidden_size=512, intra_bottleneck_size=128, num_attention_heads=4,
                           intermediate_size=512, num_feedforward_networks=2, num_hidden_layers=2, embedding_size=128,
                           hidden_act="relu", normalization_type="no_norm", hidden_dropout_prob=0.0,
                           attention_probs_dropout_prob=0.0, attn_implementation="eager")
    model = MobileBertForQuestionAnswering(cfg).to(DEV).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "LayerNorm.weight" in n:
                p.add_(0.2 * torch.randn_like(p))
            elif "LayerNorm.bias" in n:
                p.add_(0.1 * torch.randn_like(p))
    qt.quantize(model, parse("--activation", spec, "--weight", spec, "--quantize_forward", ops_str, "--bf16",
                             "--op_fusion", "qa_outputs"))
    assert isinstance(model.mobilebert.encoder.layer[0], qt.modules.quantizable.MobileBertLayer)
    ids = torch.randint(0, 600, (2, 96), device=DEV)
    am = torch.ones(2, 96, dtype=torch.long, device=DEV)
    am[1, 80:] = 0
    calls = {"n": 0}
    real = fused.mobilebert_layer_forward

    def counting(*a, **k):
        out = real(*a, **k)
        calls["n"] += out is not None
        return out

    with torch.no_grad():
        model(input_ids=ids, attention_mask=am)     # first call: the hook fake-quantizers are created lazily
    monkeypatch.setattr(fused, "mobilebert_layer_forward", counting)
    with torch.no_grad():
        got = model(input_ids=ids, attention_mask=am)
        assert calls["n"] == 2, "the fused MobileBERT layer did not run"
        fused.set_enabled(False)
        try:
            want = model(input_ids=ids, attention_mask=am)
        finally:
            fused.set_enabled(True)
    for g, w in ((got.start_logits, want.start_logits), (got.end_logits, want.end_logits)):
        same = float((g.view(torch.int16) == w.view(torch.int16)).float().mean())
        assert rel_err(g, w) < 1e-2 and same >= 0.98, (rel_err(g, w), same)


def test_dtype_casts_after_quantize_do_not_touch_the_observer_state():
    """model.bfloat16() / .half() issued AFTER quantize() (HF bf16_full_eval does this) casts every floating-point
    buffer; the fake-quantizers keep `scale` / `amax_history` in fp32 and the rounding table as an integer buffer, so
    the next forward computes exactly what it did before the cast."""
    torch.manual_seed(3)
    model = nn.Sequential(nn.Linear(64, 64), nn.GELU(), nn.Linear(64, 32)).to(DEV)
    qt.quantize(model, parse("--activation", "posit8_1,qs=per_tensor_symmetric,qmax=64", "--weight", "posit8_1", "--bf16"))
    x = torch.randn(8, 64, device=DEV).bfloat16()
    with torch.no_grad():
        model(x); model(x)
        fqs = [m for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize)]
        for m in fqs:
            m.disable_observer()
        want = model(x)
        scales = [m.scale.clone() for m in fqs]
        model.half()
        model.bfloat16()
        for m, s in zip(fqs, scales):
            assert m.scale.dtype == torch.float32 and m.amax_history.dtype == torch.float32 and torch.equal(m.scale, s)
            assert m.lut is None or m.lut.dtype == torch.int32
        got = model(x)
    assert torch.equal(got, want)


def test_qat_conv2d_and_checkpoint_resume(oracle, tmp_path):
    """reference modules/qat/conv.py:41-42 (`_conv_forward(input, weight_fake_quant(weight), bias)`) with the input
    hook in front: bit-identical to cuDNN on table-quantized operands; STE gradients reach weight and input.  Then the
    calibrated model is checkpointed (run_qa_no_trainer.py:961-990 layout) and a second, freshly calibrated model takes
    the state and reproduces the outputs bit for bit."""
    from quantized_training import checkpoint as ck
    from quantized_training.modules import qat

    def make(seed):
        torch.manual_seed(seed)
        net = nn.Sequential(nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 8, 1, bias=False)).to(DEV).bfloat16()
        qt.quantize(net, parse("--activation", "posit8_1", "--weight", "posit8_1", "--error", "posit8_1",
                               "--quantize_forward", "gemm", "--quantize_backprop", "gemm", "--bf16"))
        return net

    net = make(0)
    assert isinstance(net[0], qat.Conv2d) and isinstance(net[2], qat.Conv2d)
    x = torch.randn(2, 16, 20, 20, device=DEV).bfloat16().requires_grad_()
    y = net(x)
    fq = table_fq(oracle, "posit8_1")
    h = torch.relu(torch.nn.functional.conv2d(fq(x.detach()), fq(net[0].weight.detach()), net[0].bias, padding=1))
    ref = torch.nn.functional.conv2d(fq(h), fq(net[2].weight.detach()))
    assert torch.equal(y, ref)
    y.backward(torch.randn_like(y))
    assert x.grad is not None and net[0].weight.grad is not None and net[2].weight.grad.abs().sum() > 0

    # per-tensor scaled spec: calibrate, checkpoint, restore into another model whose quantizers saw different data
    def make_scaled(seed, data_scale):
        torch.manual_seed(seed)
        m = nn.Sequential(nn.Conv2d(16, 8, 3), nn.Flatten(), nn.Linear(8 * 18 * 18, 4)).to(DEV).bfloat16()
        qt.quantize(m, parse("--activation", "int8,qs=per_tensor_symmetric,ahl=4", "--weight",
                             "int8,qs=per_channel_symmetric,ax=0", "--quantize_forward", "gemm", "--bf16"))
        with torch.no_grad():
            m(torch.randn(2, 16, 20, 20, device=DEV).bfloat16() * data_scale)
        return m

    a, b = make_scaled(1, 1.0), make_scaled(2, 7.0)
    xs = torch.randn(2, 16, 20, 20, device=DEV).bfloat16()
    a.eval(), b.eval()
    for m in (a, b):
        for mod in m.modules():
            if isinstance(mod, qt.FusedAmaxObsFakeQuantize):
                mod.disable_observer()
    with torch.no_grad():
        want = a(xs)
        assert not torch.equal(b(xs), want)
        ck.save_state(tmp_path / "epoch_0", a, best_metric={"f1": 1.0})
        assert ck.load_state(tmp_path / "epoch_0", b, map_location=DEV)["best_metric"] == {"f1": 1.0}
        assert torch.equal(b(xs), want)
