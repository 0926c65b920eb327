"""Llama-2-7B-shape quantized forward (BASELINE configs[4]): random-init bf16 weights, synthetic tokens,
windows [1, 1024], `quantize(model, args)` with --quantize_forward gemm (the "+residual fusion" level).
Prints one JSON line per variant: tokens/s, ms per window, kernel mix.  Usage:
    python scripts/llama_bench.py [--spec posit8_1] [--layers 32] [--steps 10] [--graph] [--torch-gemm] [--no-fused]
bench.py imports `setup` / `measure` from here for the tokens/s part of its line."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))

LLAMA2_7B = dict(hidden_size=4096, intermediate_size=11008, num_attention_heads=32, num_key_value_heads=32,
                 vocab_size=32000, max_position_embeddings=4096, rms_norm_eps=1e-5)


def flops_per_window(layers, seq, batch=1, c=LLAMA2_7B):
    h, i, v = c["hidden_size"], c["intermediate_size"], c["vocab_size"]
    dense = 2 * seq * (layers * (4 * h * h + 3 * h * i) + v * h)
    attn = layers * 4 * seq * seq * h
    return batch * (dense + attn)


def build(layers, dev):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(num_hidden_layers=layers, attn_implementation="eager", tie_word_embeddings=False, **LLAMA2_7B)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg)
    torch.set_default_dtype(torch.float32)
    return model.eval()


def setup(spec, dev, layers=32, seq=1024, batch=1, ops_str="gemm", seed=0):
    """Random-init Llama-2-7B-shape model, quantize()-d; returns (model, step) where step(ids_host_pinned or None)
    runs one window and returns the loss tensor (on the device)."""
    import quantized_training as qt
    torch.manual_seed(seed)
    model = build(layers, dev)
    args = qt.add_qspec_args().parse_args(["--activation", spec, "--weight", spec, "--quantize_forward", ops_str, "--bf16"])
    qt.quantize(model, args)
    ids = torch.randint(0, LLAMA2_7B["vocab_size"], (batch, seq), device=dev)
    # prebuilt additive causal mask [1, 1, S, S]: HF then skips its own mask construction (which copies a CPU
    # scalar to the device and cannot be captured in a CUDA graph)
    mask = torch.full((seq, seq), torch.finfo(torch.bfloat16).min, device=dev, dtype=torch.bfloat16).triu(1)[None, None]
    pos = torch.arange(seq, device=dev)[None].expand(batch, -1)

    def fwd():
        with torch.no_grad():
            out = model(input_ids=ids, labels=ids, use_cache=False, attention_mask=mask, position_ids=pos)
        return out.loss

    return model, fwd, ids


def measure(fwd, steps, graph=True, warmup=3, ids=None, barrier=None):
    """(ms per window on the device, e2e seconds per window or None, loss).  With `graph` the whole forward is one
    CUDA graph.  e2e: per step the token ids are copied from pinned host memory and the loss is read back."""
    for _ in range(2):          # first call creates the lazy fake-quantizers, second takes the fused path
        loss = fwd()
    torch.cuda.synchronize()
    run = fwd
    if graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            gl = fwd()
        run = lambda: (g.replay(), gl)[1]
    for _ in range(max(warmup, 1)):
        loss = run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = run()
    e1.record()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    e2e = None
    if ids is not None:
        host_ids = torch.randint(0, LLAMA2_7B["vocab_size"], ids.shape).pin_memory()
        host_loss = torch.empty((), dtype=torch.float32).pin_memory()
        n = max(1, min(steps, 10))
        run(); torch.cuda.synchronize()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            ids.copy_(host_ids, non_blocking=True)      # H2D of the step's input from pinned memory
            l = run()
            host_loss.copy_(l.float(), non_blocking=True)   # D2H of the step's result
            torch.cuda.synchronize()
        if barrier:
            barrier()
        e2e = (time.perf_counter() - t0) / n
    return ms, e2e, float(loss)


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
    import quantized_training as qt
    from quantized_training import fused, ops
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    t0 = time.time()
    model, fwd, ids = setup(a.spec, dev, a.layers, a.seq, a.batch, a.ops)
    build_s = time.time() - t0
    if a.torch_gemm:
        ops.set_enabled(False)
    if a.no_fused:
        fused.set_enabled(False)
    ms, e2e, loss = measure(fwd, a.steps, a.graph, ids=ids)
    nfq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    fl = flops_per_window(a.layers, a.seq, a.batch)
    print(json.dumps({"workload": f"Llama-2-7B-shape quantized forward, {a.layers} layers, window [{a.batch},{a.seq}]",
                      "spec": a.spec, "quantize_forward": a.ops, "graph": a.graph, "fused_blocks": not a.no_fused,
                      "gemm": "torch/cuBLAS" if a.torch_gemm else "qt_gemm_nt",
                      "ms_per_window": ms, "batch": a.batch, "tokens_per_s": a.batch * a.seq / ms * 1e3,
                      "e2e_tokens_per_s": a.batch * a.seq / e2e if e2e else None, "TFLOPs": fl / ms / 1e9,
                      "loss": loss, "fake_quant_modules": nfq, "build_s": build_s,
                      "flops_per_window_T": fl / 1e12}), flush=True)


if __name__ == "__main__":
    main()
