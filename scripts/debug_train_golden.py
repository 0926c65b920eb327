import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import quantized_training as qt
if os.environ.get("NO_RPR"): torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
import test_model_golden_gpu as T
G = np.load(T.GOLDEN); name = "fp8_train"
act, weight, error, fwd, bwd = T.CASES[name]
model = T.build_host(); model.load_state_dict({k[2:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("w/")}); model.to(T.DEV)
qt.quantize(model, qt.add_qspec_args().parse_args(["--activation", act, "--weight", weight, "--quantize_forward", fwd, "--bf16", "--quantize_backprop", bwd, "--error", error]))
x = torch.from_numpy(G["x"]).to(T.DEV).bfloat16().requires_grad_(True); mask = torch.from_numpy(G["mask"]).to(T.DEV).bfloat16()
model.train()
for _ in range(2):
    x.grad = None; y = model(x, mask); y.float().square().sum().backward()
ours = {n: m for n, m in model.named_modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize) and "error_" in n}
ref_names = sorted(k[len(name) + 7:] for k in G.files if k.startswith(name + "/scale/"))
print("ours", len(ours), "ref", len(ref_names), "only ours:", sorted(set(ours) - set(ref_names))[:5], "only ref:", sorted(set(ref_names) - set(ours))[:5])
for n in ref_names:
    if n in ours:
        rs, rh = G[f"{name}/scale/{n}"], G[f"{name}/hist/{n}"]
        os_, oh = ours[n].scale.detach().float().reshape(-1).cpu().numpy(), ours[n].amax_history.detach().float().reshape(-1).cpu().numpy()
        flag = "" if np.allclose(rs, os_, rtol=1e-3) and np.allclose(rh[:2], oh[:2], rtol=1e-3) else "   <-- differs"
        print(f"{n:60s} scale ref {rs[0]:.4e} ours {os_[0]:.4e} | hist ref {rh[:2]} ours {oh[:2]}{flag}")
