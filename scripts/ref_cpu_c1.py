"""BASELINE configs[0]: the REFERENCE's own CPU path, timed -- MobileBERT-tiny-shaped encoder layers built from the
reference's quantizable blocks (tests/golden/hosts.py), `quantize(model, args)` of the unmodified reference with
--activation e4m3 --weight e4m3 --quantize_forward gemm,residual,layernorm,activation,scaling --bf16, input
[16, 384, 512] on the host cores of THIS container (the reference is pure Python and cannot travel to the GPU box;
BASELINE.md §3).  Two layers are timed and the 21-layer encoder of models/mobilebert_tiny_squad is extrapolated.

    python scripts/ref_cpu_c1.py > profiles/ref_cpu_c1_r02.json      (container with /root/reference)"""
import json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import hosts
from gen_model_golden import load_reference_quantize

ref = load_reference_quantize()
mm = sys.modules["quantized_training.modules.quantizable.modeling_mobilebert"]
blocks = types.SimpleNamespace(MobileBertSelfAttention=mm.MobileBertSelfAttention, MobileBertSelfOutput=mm.MobileBertSelfOutput,
                               FFNOutput=mm.FFNOutput, MobileBertOutput=mm.MobileBertOutput)
cfg = hosts.mobilebert_config(hidden=512, true_hidden=128, heads=4, inter=512, ffn=2)
LAYERS = 2
torch.manual_seed(0)
model = hosts.MobileBertHost(blocks, cfg, layers=LAYERS)
args = ref.training_args.add_qspec_args().parse_args(
    ["--activation", "e4m3", "--weight", "e4m3", "--quantize_forward", "gemm,residual,layernorm,activation,scaling", "--bf16"])
args.error = None
ref.quantize.quantize(model, args)
model.eval()
B, S = 16, 384
x = torch.randn(B, S, cfg.hidden_size).bfloat16()
mask = torch.zeros(B, 1, 1, S).bfloat16()
out = {}
for threads in (os.cpu_count(), 1):
    torch.set_num_threads(threads)
    with torch.no_grad():
        model(x, mask)
        reps = 2 if threads > 1 else 1
        t0 = time.perf_counter()
        for _ in range(reps):
            model(x, mask)
        dt = (time.perf_counter() - t0) / reps
    per_layer = dt / LAYERS
    out[f"threads_{threads}"] = {"seconds_per_layer": per_layer, "extrapolated_21_layer_forward_s": per_layer * 21,
                                 "tokens_per_s_21_layers": B * S / (per_layer * 21)}
print(json.dumps({"what": "reference quantize()-d MobileBERT-tiny encoder layers on CPU (this container), e4m3, all five op groups, "
                          "batch 16 x seq 384", "cpu_count": os.cpu_count(), "layers_timed": LAYERS, "results": out}, indent=1))
