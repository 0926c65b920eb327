"""Per-kernel device times INSIDE the captured forward (torch.profiler / CUPTI on graph replays), aggregated by kernel.
Usage: python scripts/llama_profile.py --spec e4m3 [--layers 8]"""
import argparse, collections, json, os, sys, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import llama_bench as LB
ap = argparse.ArgumentParser(); ap.add_argument("--spec", default="e4m3"); ap.add_argument("--layers", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
model, fwd, ids = LB.setup(a.spec, dev, layers=a.layers)
for _ in range(2): fwd()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fwd()
for _ in range(3): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict(); t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
for e in ev:
    k = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))[:80]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += e.time_range.elapsed_us()
busy = sum(v[1] for v in agg.values())
print(f"spec {a.spec} layers {a.layers}: span {(t1-t0)/3:.1f} us per replay, kernel busy {busy/3:.1f} us per replay")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v[0]//3:4d} x {v[1]/v[0]:8.1f} us = {v[1]/3:9.1f} us  {k}")
