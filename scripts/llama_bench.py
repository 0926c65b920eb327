"""Llama-2-7B-shape quantized forward (BASELINE configs[4]): random-init bf16 weights, synthetic tokens,
windows [1, 1024], `quantize(model, args)` with --quantize_forward gemm (the "+residual fusion" level).
Prints one JSON line per variant: tokens/s, ms per window, kernel mix.  Usage:
    python scripts/llama_bench.py [--spec posit8_1] [--layers 32] [--steps 10] [--graph] [--torch-gemm]"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import ops
from transformers import LlamaConfig, LlamaForCausalLM


def build(layers, dev):
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=layers, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096, rms_norm_eps=1e-5,
                      attn_implementation="eager", tie_word_embeddings=False)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg)
    torch.set_default_dtype(torch.float32)
    return model.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", default="posit8_1")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--torch-gemm", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--ops", default="gemm")
    a = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    t0 = time.time()
    model = build(a.layers, dev)
    args = qt.add_qspec_args().parse_args(["--activation", a.spec, "--weight", a.spec, "--quantize_forward", a.ops, "--bf16"])
    qt.quantize(model, args)
    build_s = time.time() - t0
    ids = torch.randint(0, 32000, (a.batch, a.seq), device=dev)
    # prebuilt additive causal mask [1, 1, S, S]: HF then skips its own mask construction (which copies a CPU
    # scalar to the device and cannot be captured in a CUDA graph)
    mask = torch.full((a.seq, a.seq), torch.finfo(torch.bfloat16).min, device=dev, dtype=torch.bfloat16).triu(1)[None, None]
    pos = torch.arange(a.seq, device=dev)[None].expand(a.batch, -1)
    if a.torch_gemm:
        ops.set_enabled(False)
    if a.no_fused:
        from quantized_training import fused
        fused.set_enabled(False)

    def fwd():
        with torch.no_grad():
            out = model(input_ids=ids, labels=ids, use_cache=False, attention_mask=mask, position_ids=pos)
        return out.loss

    for _ in range(2):
        loss = fwd()
    torch.cuda.synchronize()
    run = fwd
    if a.graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            gl = fwd()
        run = lambda: (g.replay(), gl)[1]
    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1) / a.steps
    nfq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    print(json.dumps({"workload": f"Llama-2-7B-shape quantized forward, {a.layers} layers, window [1,{a.seq}]",
                      "spec": a.spec, "quantize_forward": a.ops, "graph": a.graph, "fused_blocks": not a.no_fused, "gemm": "torch/cuBLAS" if a.torch_gemm else "qt_gemm_nt",
                      "ms_per_window": ms, "batch": a.batch, "tokens_per_s": a.batch * a.seq / ms * 1e3, "wall_ms_per_window": wall / a.steps * 1e3,
                      "loss": float(loss), "fake_quant_modules": nfq, "build_s": build_s,
                      "flops_per_window_T": (2 * a.seq * (a.layers * (4 * 4096 * 4096 + 3 * 4096 * 11008) + 32000 * 4096) + a.layers * 4 * 32 * a.seq * a.seq * 128) / 1e12}), flush=True)


if __name__ == "__main__":
    main()
