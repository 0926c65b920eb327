"""Quantized encoder forward on synthetic SQuAD-shape batches (BASELINE configs[0] / configs[1]):
  --model bert-base        BERT-base, posit(8,1) act + weight, --quantize_forward gemm (config 1: every other op group
                           fused: attention scaling, activation, LayerNorm, residual), batch 16 x seq 384
  --model mobilebert-tiny  models/mobilebert_tiny_squad shape (config 0's model), e4m3, all five op groups
Random-init weights (no network), bf16.  Prints one JSON line: ms per batch, sequences/s, tokens/s, TFLOP/s."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import fused

def build(name, dev):
    from transformers import BertConfig, BertForQuestionAnswering, MobileBertConfig, MobileBertForQuestionAnswering
    if name == "bert-base":
        cfg = BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = BertForQuestionAnswering(cfg)
        h, i, L, v = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, None
        flops = lambda B, S: B * (2 * S * L * (4 * h * h + 2 * h * i) + L * 4 * S * S * h)
    else:  # MobileBERT-tiny: reference models/mobilebert_tiny_squad/config.json
        cfg = MobileBertConfig(hidden_size=512, embedding_size=128, intra_bottleneck_size=128, true_hidden_size=128,
                               intermediate_size=512, num_attention_heads=4, num_feedforward_networks=2,
                               num_hidden_layers=21, hidden_act="relu", normalization_type="no_norm",
                               hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, trigram_input=True,
                               use_bottleneck=True, use_bottleneck_attention=False, key_query_shared_bottleneck=True,
                               classifier_activation=False)
        model = MobileBertForQuestionAnswering(cfg)
        flops = lambda B, S: 0
    return model.to(dev).eval(), cfg, flops

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="bert-base")
    ap.add_argument("--spec", default="posit8_1")
    ap.add_argument("--ops", default="gemm")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seq", type=int, default=384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0"); torch.cuda.set_device(dev); torch.manual_seed(0)
    model, cfg, flops = build(a.model, dev)
    qt.quantize(model, qt.add_qspec_args().parse_args(["--activation", a.spec, "--weight", a.spec, "--quantize_forward", a.ops,
                                                       "--bf16", "--op_fusion", "qa_outputs"]))
    if a.no_fused:
        fused.set_enabled(False)
    ids = torch.randint(0, cfg.vocab_size, (a.batch, a.seq), device=dev)
    tt = torch.zeros_like(ids)
    def fwd():
        with torch.no_grad():
            return model(input_ids=ids, token_type_ids=tt).start_logits
    for _ in range(3): out = fwd()
    torch.cuda.synchronize()
    run = fwd
    if not a.no_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            go = fwd()
        run = lambda: (g.replay(), go)[1]
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(a.steps): out = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    fl = flops(a.batch, a.seq)
    print(json.dumps({"workload": f"{a.model} quantized forward, batch {a.batch} x seq {a.seq}, random-init bf16", "spec": a.spec,
                      "quantize_forward": a.ops, "fused_blocks": not a.no_fused, "graph": not a.no_graph, "ms_per_batch": ms,
                      "sequences_per_s": a.batch / ms * 1e3, "tokens_per_s": a.batch * a.seq / ms * 1e3,
                      "TFLOPs": fl / ms / 1e9 if fl else None, "finite": bool(torch.isfinite(out.float()).all())}), flush=True)

if __name__ == "__main__":
    main()
