import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import fused, ops, _C
from transformers import LlamaConfig, LlamaForCausalLM
DEV = "cuda:0"
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
torch.manual_seed(2)
for spec, ops_str in [("e4m3", "gemm"), ("posit8_1", "gemm"), ("e4m3", "gemm,residual,layernorm,activation,scaling")]:
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, attn_implementation="eager")
    model = LlamaForCausalLM(cfg).to(DEV).eval()
    qt.quantize(model, qt.add_qspec_args().parse_args(["--activation", spec, "--weight", spec, "--quantize_forward", ops_str, "--bf16"]))
    ids = torch.randint(0, 512, (1, 128), device=DEV)
    caps = {}
    def mk(name):
        def hook(m, a, kw, out):
            caps.setdefault(name, []).append((kw.get("attention_mask"), out.detach().clone()))
        return hook
    for i, l in enumerate(model.model.layers):
        l.register_forward_hook(mk(f"layer{i}"), with_kwargs=True)
    with torch.no_grad():
        model(input_ids=ids, use_cache=False)
        a = model(input_ids=ids, use_cache=False).logits      # layer fwd hooks present -> are they blocking fusion? (layer-level hooks are allowed)
        fused.set_enabled(False)
        b = model(input_ids=ids, use_cache=False).logits
        ops.set_enabled(False)
        c = model(input_ids=ids, use_cache=False).logits
        ops.set_enabled(True); fused.set_enabled(True)
    m = caps["layer0"][0][0]
    print(spec, ops_str, "mask:", None if m is None else (m.dtype, tuple(m.shape)))
    for i in range(2):
        o = [x[1] for x in caps[f"layer{i}"]]
        print(f"  layer{i}: fused vs module {rel(o[1], o[2]):.4f}   module(qt gemm) vs module(cuBLAS) {rel(o[2], o[3]):.4f}  fused vs cuBLAS {rel(o[1], o[3]):.4f}")
    print(f"  logits: fused vs module {rel(a, b):.4f}  module vs cublas {rel(b, c):.4f}")
