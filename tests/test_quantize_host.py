"""CPU tests of the plugin surface: add_qspec_args, quantize/convert/prepare, module swaps.
No fake-quant compute happens here (that needs CUDA); blocks are checked for structural and
numerical equivalence with the stock Hugging Face blocks they replace."""
import copy

import pytest
import torch
from torch import nn
from transformers import (BertConfig, BertForQuestionAnswering, LlamaConfig, LlamaForCausalLM, MobileBertConfig,
                          MobileBertForQuestionAnswering, RobertaConfig, RobertaForSequenceClassification)

import quantized_training as qt
from quantized_training.modules import apply_lora, qat, quantizable
from quantized_training.quantization_mappings import TRANSFORMER_MODULE_MAPPINGS

ALL_OPS = "gemm,residual,layernorm,activation,scaling"


def tiny(family):
    torch.manual_seed(0)
    if family == "bert":
        return BertForQuestionAnswering(BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                   intermediate_size=128, vocab_size=100)), 100
    if family == "roberta":
        return RobertaForSequenceClassification(RobertaConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                              intermediate_size=128, vocab_size=100, num_labels=3)), 100
    if family == "mobilebert":
        return MobileBertForQuestionAnswering(MobileBertConfig(
            hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64, embedding_size=32,
            intra_bottleneck_size=32, true_hidden_size=32, vocab_size=100, hidden_act="relu",
            num_feedforward_networks=2, normalization_type="no_norm")), 100
    if family == "llama":
        return LlamaForCausalLM(LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                            num_attention_heads=4, num_key_value_heads=2, vocab_size=100,
                                            attn_implementation="eager")), 100
    raise ValueError(family)


def parse(*argv):
    return qt.add_qspec_args().parse_args(list(argv))


def test_flags_and_defaults():
    a = parse()
    assert a.quantize_forward == "gemm" and a.quantize_backprop == "gemm" and a.activation is None
    assert a.lora_rank == 0 and a.lora_alpha == 8 and a.target_modules == ["query", "value"] and not a.bf16
    a = parse("--activation", "fp8_e4m3", "--weight", "fp8_e4m3", "--error",
              "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10", "--op_fusion", "classifier,qa_outputs",
              "--target_modules", "query,value", "--bf16", "--do_train", "--lora_rank", "8")
    assert a.op_fusion == ["classifier", "qa_outputs"] and a.target_modules == ["query", "value"]
    assert isinstance(a.error, str)  # stays a string; get_qconfig accepts both forms
    a = parse("--run_name", "x", "slurm", "--job-name", "j")
    assert a.action == "slurm" and a.job_name == "j"


@pytest.mark.parametrize("family", ["bert", "roberta", "mobilebert", "llama"])
def test_block_swap_preserves_the_float_model(family):
    model, vocab = tiny(family)
    model.eval()
    ids = torch.randint(0, vocab, (2, 16))
    with torch.no_grad():
        want = model(input_ids=ids)[0]
    qt.propagate_config(model, "config", model.config)
    qt.convert(model, inplace=True, custom_module_class_mapping=TRANSFORMER_MODULE_MAPPINGS)
    kinds = {type(m) for m in model.modules()}
    assert quantizable.MatmulFunctional in kinds and quantizable.AddFunctional in kinds
    assert not (kinds & set(TRANSFORMER_MODULE_MAPPINGS)), "a float block survived the swap"
    from quantized_training import ops
    ops.set_enabled(False)   # explicit A/B mode (stock torch products): this host-side structural check runs on CPU
    try:
        with torch.no_grad():
            got = model(input_ids=ids)[0]
    finally:
        ops.set_enabled(True)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


def hooked(model, kind="activation_pre_process"):
    return sorted(n for n, m in model.named_modules() if hasattr(m, kind))


def test_hook_placement_follows_op_groups():
    model, _ = tiny("bert")
    args = parse("--activation", "posit8_1", "--weight", "posit8_1", "--quantize_forward", ALL_OPS,
                 "--op_fusion", "qa_outputs")
    qt.quantize(model, args)
    names = hooked(model)
    L0 = "bert.encoder.layer.0."
    for leaf in ["attention.self.query", "attention.self.key", "attention.self.value", "attention.self.qk_matmul",
                 "attention.self.av_matmul", "attention.self.attn_scaling", "attention.self.softmax",
                 "attention.output.dense", "attention.output.residual", "attention.output.LayerNorm",
                 "intermediate.dense", "intermediate.intermediate_act_fn", "output.dense", "output.residual",
                 "output.LayerNorm"]:
        assert L0 + leaf in names, leaf
    assert "bert.embeddings.LayerNorm" in names and "qa_outputs" not in names
    assert isinstance(model.bert.encoder.layer[0].attention.self.query, qat.Linear)
    assert model.bert.encoder.layer[0].attention.self.query.weight.dtype == torch.float32  # no --bf16

    # "+residual fusion" level: gemm only
    model, _ = tiny("bert")
    args = parse("--activation", "posit8_1", "--weight", "posit8_1", "--bf16")
    qt.quantize(model, args)
    names = hooked(model)
    assert all(n.rsplit(".", 1)[-1] in {"query", "key", "value", "dense", "qk_matmul", "av_matmul", "qa_outputs"}
               for n in names), names
    assert model.bert.encoder.layer[0].output.dense.weight.dtype == torch.bfloat16
    assert args.quantize_backprop is None and args.quantize_forward == "gemm"  # written back, as the reference does


def test_weight_only_leaves_blocks_alone():
    model, _ = tiny("bert")
    args = parse("--weight", "int8,qs=per_channel_symmetric,ax=0")
    qt.quantize(model, args)
    assert args.quantize_forward is None
    assert type(model.bert.encoder.layer[0].attention.self).__module__.startswith("transformers.")
    assert hooked(model) == []
    lin = model.bert.encoder.layer[0].output.dense
    assert isinstance(lin, qat.Linear) and lin.weight_fake_quant.is_per_channel and lin.weight_fake_quant.ch_axis == 0


def test_backward_hooks_and_lora():
    model, _ = tiny("roberta")
    apply_lora(model, ["query", "value"], r=4, lora_alpha=8)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    args = parse("--activation", "fp8_e4m3", "--weight", "fp8_e4m3", "--error",
                 "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10", "--quantize_forward",
                 "gemm,residual,layernorm,activation", "--quantize_backprop", "gemm,residual,layernorm,activation",
                 "--bf16")
    qt.quantize(model, args)
    layer = model.roberta.encoder.layer[0]
    assert isinstance(layer.attention.self.query, qat.LoraLinear) and isinstance(layer.attention.self.key, qat.Linear)
    assert layer.attention.self.query.weight is not None and layer.attention.self.query.lora_A["default"].weight.requires_grad
    pre, post = hooked(model, "error_pre_process"), hooked(model, "error_post_process")
    L0 = "roberta.encoder.layer.0."
    assert L0 + "attention.self.query" in pre and L0 + "attention.output.residual" in pre
    assert L0 + "attention.self.query" in post and L0 + "intermediate.dense" in post
    assert L0 + "attention.output.residual" in post and L0 + "attention.output.dense" not in post
    assert "classifier.dense" in pre  # RoBERTa recipe has no --op_fusion: the head is quantized too
    fq = layer.attention.self.query.qconfig.error()
    assert fq.quant_max == 57344.0 and fq.amax_history_len == 10 and fq.qscheme == qt.per_tensor_symmetric


def test_invalid_op_group():
    model, _ = tiny("bert")
    with pytest.raises(AssertionError, match="Invalid operation"):
        qt.quantize(model, parse("--activation", "int8", "--quantize_forward", "gemm,conv"))


def test_llama_and_mobilebert_hook_points():
    model, _ = tiny("llama")
    qt.quantize(model, parse("--activation", "posit8_1", "--weight", "posit8_1", "--quantize_forward", ALL_OPS, "--bf16"))
    names = hooked(model)
    L0 = "model.layers.0."
    for leaf in ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "self_attn.qk_matmul",
                 "self_attn.av_matmul", "self_attn.attn_scaling", "self_attn.softmax", "self_attn_residual",
                 "mlp_residual", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj", "input_layernorm",
                 "post_attention_layernorm"]:
        assert L0 + leaf in names, leaf
    assert "model.norm" in names and "lm_head" in names and L0 + "mlp.act_fn" not in names  # SiLU is never quantized
    model, _ = tiny("mobilebert")
    qt.quantize(model, parse("--activation", "e4m3", "--weight", "e4m3", "--quantize_forward", ALL_OPS, "--bf16"))
    names = hooked(model)
    L0 = "mobilebert.encoder.layer.0."
    for leaf in ["attention.self.qk_matmul", "attention.self.attn_scaling", "attention.output.residual",
                 "attention.output.LayerNorm", "ffn.0.output.residual", "output.residual", "output.bottleneck.residual",
                 "output.bottleneck.dense", "bottleneck.input.dense", "bottleneck.input.LayerNorm"]:
        assert L0 + leaf in names, leaf


class _ConvNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.conv3 = nn.Conv3d(2, 4, (1, 3, 3), bias=False)
        self.head = nn.Linear(8, 4)


def test_conv_modules_become_qat_and_back():
    """reference quantization_mappings.py:17-18 + modules/qat/conv.py: Conv2d / Conv3d are swapped for weight-quantized
    subclasses that share the float Parameters, and `to_float` undoes it."""
    net = _ConvNet()
    w = net.conv.weight
    qt.quantize(net, parse("--weight", "int8,qs=per_channel_symmetric,ax=0", "--activation", "int8,qs=per_tensor_symmetric",
                           "--quantize_forward", "gemm"))
    assert isinstance(net.conv, qat.Conv2d) and isinstance(net.conv, nn.Conv2d)
    assert isinstance(net.conv3, qat.Conv3d) and net.conv3.bias is None
    assert net.conv.weight is w and net.conv.stride == (2, 2) and net.conv.padding == (1, 1)
    assert net.conv.weight_fake_quant.is_per_channel and net.conv.weight_fake_quant.ch_axis == 0
    assert "conv" in hooked(net) and "conv3" in hooked(net)  # inputs are hooked like any gemm-group module
    back = net.conv.to_float()
    assert type(back) is nn.Conv2d and torch.equal(back.weight, w) and back.kernel_size == (3, 3)
    with pytest.raises(AssertionError, match="only works for Conv2d"):
        qat.Conv2d.from_float(nn.Conv1d(1, 1, 1))


def test_checkpoint_tar_round_trip(tmp_path):
    """reference run_qa_no_trainer.py:961-990 / :1021-1056: checkpoint.tar layout, resume bookkeeping, and a freshly
    prepared model taking quantizer buffers that were shaped after the checkpointed run's calibration."""
    from quantized_training import checkpoint as ck

    def make():
        torch.manual_seed(0)
        m = nn.Sequential(nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 2))
        qt.quantize(m, parse("--weight", "int8,qs=per_channel_symmetric,ax=0", "--activation",
                             "int8,qs=per_tensor_symmetric,ahl=4", "--quantize_forward", "gemm"))
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
        return m, opt, torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))

    m, opt, sch = make()
    with torch.no_grad():  # stand-in for a calibrated + trained state (the kernels themselves need a GPU)
        m[0].weight_fake_quant.scale = torch.rand(8, 1) + 0.5
        m[0].weight_fake_quant.amax_history = torch.full((4, 8), 3.0)
        m[0].weight.add_(1.0)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step(), sch.step()
    path = ck.save_state(tmp_path / "step_30", m, opt, sch, best_metric={"f1": 88.5}, run_id="abc")
    assert path.endswith("step_30/checkpoint.tar")
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_metric", "run_id"}

    m2, opt2, sch2 = make()
    got = ck.load_state(tmp_path / "step_30", m2, opt2, sch2)
    assert got["best_metric"] == {"f1": 88.5} and got["run_id"] == "abc"
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert a.shape == b.shape and torch.equal(a, b), k
    assert m2[0].weight_fake_quant.scale.shape == (8, 1) and sch2.last_epoch == 1
    assert opt2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]

    ck.save_state(tmp_path / "epoch_0", m, opt, sch)
    assert ck.find_latest(tmp_path).endswith("epoch_0")
    # 100 batches per epoch, 4-batch accumulation: step_30 = 120 batches in = epoch 1, 20 batches to skip
    assert ck.parse_resume(tmp_path / "step_30", 100, 4) == (1, 20, 30)
    assert ck.parse_resume(tmp_path / "epoch_2", 100, 4) == (3, None, 75)
    with pytest.raises(ValueError):
        ck.parse_resume(tmp_path / "final", 100)
