#!/usr/bin/env python
"""bench.py -- standalone quantize sweep (BASELINE.json configs[2]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2-numel L]

A step = one pass of the fake-quant hot path over one batch: a resident bf16 tensor of 2^L
elements per GPU (default 2^32 = 8 GiB in + 8 GiB out, the top of BASELINE's 1M-4G range) is
quantized once with every spec of the sweep (int4, int8, e4m3, e5m2, fp6_e3m2, fp4_e2m1,
posit8_1, posit8_2: 8 launches of the same kernel family).  Algorithmic bytes per launch
= 2 * 2 B * 2^L.  `value` is whole-job algorithmic GB/s with inputs resident
in HBM; `e2e` runs the same sweep through the public module API from pinned HOST buffers
(H2D + kernels + D2H inside the timed region); `roofline` compares the kernel with the
measured HBM copy peak; `cpu_baseline` times the CPU oracle port on the box's host cores.
`--impl reference` times that CPU port alone (the reference is pure Python and cannot travel).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
sys.path.insert(0, ROOT)

SWEEP = ["int4", "int8", "e4m3", "e5m2", "fp6_e3m2", "fp4_e2m1", "posit8_1", "posit8_2"]
METRIC = "quantize GB/s vs HBM peak (standalone fake-quant sweep, algorithmic read+write bytes)"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


FALLBACK_BF16_TFLOPS = 1590.0  # burst; ~1400 sustained (B200_PROFILING.md)


def _measured():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def peaks():
    m = _measured()
    if m:
        return float(m["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 TFLOP/s for a kernel timed inside a long step (sustained figure)."""
    m = _measured()
    if m and "bf16_tflops_sustained" in m:
        return float(m["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md, sustained)"


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture, scaled to this run's
    launch size (the capture's own size is recorded beside it)."""
    p = os.path.join(ROOT, "profiles", "fq_flat_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def run_llama(args, dev, rank, world, dist, barrier):
    """Llama-2-7B-shape quantized forward (BASELINE configs[4]): windows [1, 1024], posit8_1 and e4m3, --quantize_forward
    gemm, data-parallel over ranks (each rank its own windows, no collective on the data path)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import llama_bench as LB

    out = {"workload": "Llama-2-7B shape (32 layers, hidden 4096, 32 heads, inter 11008, vocab 32000), random-init "
                       "bf16 weights, synthetic tokens, windows [1,1024], quantize_forward=gemm, whole forward "
                       "captured in one CUDA graph", "windows_per_rank": args.llama_steps, "n_gpus": world,
           "flops_per_window_T": LB.flops_per_window(32, 1024) / 1e12, "specs": {}}
    peak, peak_src = tensor_peak()
    for spec in ("posit8_1", "e4m3"):
        model, fwd, ids = LB.setup(spec, dev, seed=rank)
        ms, e2e_s, loss = LB.measure(fwd, args.llama_steps, graph=True, warmup=max(args.warmup, 3), ids=ids,
                                     barrier=barrier)
        if dist is not None:
            t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, e2e_s = float(t[0]), float(t[1])
        tf = LB.flops_per_window(32, 1024) / ms / 1e9
        fp8 = spec == "e4m3"
        out["specs"][spec] = {
            "tokens_per_s": world * 1024 / ms * 1e3, "ms_per_window": ms, "loss": loss,
            "e2e_tokens_per_s": world * 1024 / e2e_s, "tflops_per_gpu": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "unit": "TFLOP/s",
                         "peak": peak * (2.0 if fp8 else 1.0), "frac": tf / (peak * (2.0 if fp8 else 1.0)),
                         "peak_source": peak_src + (" x 2 (fp8 MMA rate)" if fp8 else ""),
                         "operands": "e4m3 codes, kind::f8f6f4" if fp8 else "posit8 values held in bf16, kind::f16"}}
        del model, fwd
        torch.cuda.empty_cache()
    return out



def run_size_sweep(dev, qt):
    """BASELINE configs[2] size axis: 2^20 .. 2^32 bf16 elements on one GPU, L2 defeated by cycling through a pool of
    distinct buffers larger than the 126 MB L2.  `device`: launches replayed from a CUDA graph (device time, what a
    captured model step sees); `api`: the public module call issued eagerly from Python, wall clock (host overhead of
    the ctypes boundary included).  GB/s of algorithmic read + write bytes."""
    import torch
    out = {"unit": "GB/s", "l2": "pool of distinct in/out buffers > 256 MB cycled between launches (2^26 and up: a "
                                 "single launch already exceeds L2)", "sizes": {}}
    unit = torch.ones(1, device=dev)
    sc = torch.full((1,), 0.0123, device=dev)
    free_b = torch.cuda.mem_get_info(dev)[0]
    for L in range(20, 33, 2):
        n = 1 << L
        if 4 * n * 2 + (1 << 30) > free_b:
            out["sizes"][str(L)] = {"skipped": "not enough free memory"}
            continue
        k = max(1, min(64, (256 << 20) // (4 * n) + 1)) if L < 26 else 1
        xs = [(torch.randn(min(n, 1 << 24), device=dev) * 3).to(torch.bfloat16) for _ in range(min(k, 4))]
        xs = [x.repeat(n // x.numel()) if n > x.numel() else x for x in xs]
        xs = [xs[i % len(xs)].clone() if i >= len(xs) else xs[i] for i in range(k)]
        ys = [torch.empty_like(x) for x in xs]
        row = {"buffers": k}
        for label, spec, scaled in (("e4m3 bare", "e4m3", False), ("posit8_1 bare", "posit8_1", False),
                                    ("e4m3 per-tensor scale + amax", "e4m3", True)):
            m = qt.FusedAmaxObsFakeQuantize(spec, device=dev)
            hist = torch.zeros(1, device=dev)

            def launch(i):
                qt._C.fq_forward(xs[i], ys[i], 1, 1, n, m._fmt, sc if scaled else unit, hist if scaled else None, m.lut)

            for i in range(k):
                launch(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(k):
                    launch(i)
            g.replay()
            reps = max(2, min(20, (1 << 30) // (n * k) + 1))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            row[label + " (device)"] = 4.0 * n * k * reps / (a.elapsed_time(b) * 1e-3) / 1e9
            del g
        qs = qt.QuantizationSpec.from_str("fp8_e4m3,qs=per_tensor_symmetric")
        mod = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), device=dev)
        for i in range(k):
            mod(xs[i])
        torch.cuda.synchronize()
        reps = max(2, min(20, (1 << 30) // (n * k) + 1))
        t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(k):
                mod(xs[i])
        torch.cuda.synchronize()
        row["fp8_e4m3 per-tensor scale + amax (api, eager module call)"] = 4.0 * n * k * reps / (time.perf_counter() - t0) / 1e9
        out["sizes"][str(L)] = row
        del xs, ys
        torch.cuda.empty_cache()
    return out


def run_gemm(dev):
    """BASELINE metric (ii): quantized GEMM / BMM TFLOP/s on the configs' shapes, bf16 operands and e4m3 codes, with
    cuBLAS (torch.matmul) on the same box beside it, against the measured dense bf16 peak (burst: kernels timed alone)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gemm_bench as GB
    m = _measured()
    peak = float(m["bf16_tflops"]) if m and "bf16_tflops" in m else FALLBACK_BF16_TFLOPS
    src = "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if m and "bf16_tflops" in m else "fallback (B200_PROFILING.md)"
    res = GB.run(str(dev), verbose=False, quick=True)
    shapes = {}
    for name, r in res.items():
        if name.startswith("bwd "):
            shapes[name] = {k: round(v, 1) for k, v in r.items()}
            continue
        shapes[name] = {"qt_bf16_TFLOPs": round(r["qt_bf16_TF"], 1), "cublas_bf16_TFLOPs": round(r["cublas_bf16_TF"], 1),
                        "qt_fp8_TFLOPs": round(r["qt_fp8_TF"], 1), "frac_bf16": r["qt_bf16_TF"] / peak,
                        "frac_fp8_of_2x": r["qt_fp8_TF"] / (2 * peak), "vs_cublas_bf16": r["qt_bf16_TF"] / r["cublas_bf16_TF"]}
        if "qt_mxfp8_TF" in r:   # block-scaled fp8 (kind::mxf8f6f4.block_scale), the linear_mx product
            shapes[name]["qt_mxfp8_TFLOPs"] = round(r["qt_mxfp8_TF"], 1)
    return {"bound": "tensor", "unit": "TFLOP/s", "peak_bf16": peak, "peak_source": src,
            "timing": "20 launches per CUDA-graph replay, CUDA events, operands rotate through L2-resident buffers "
                      "(GEMM operands are re-read from L2 by design)", "shapes": shapes}


def run_finetune(args, dev, rank, world, dist):
    """BASELINE configs[3] at N GPUs: RoBERTa-base LoRA fine-tune step, NCCL all-reduce of the trainable gradients
    overlapped with backward, whole step in one CUDA graph; plus the same step without the all-reduce (its cost)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import finetune_step as FT
    rec = FT.run(dev, rank, world, dist, steps=args.finetune_steps, warmup=3, graph=True, allreduce=True)
    if world > 1:
        no_ar = FT.run(dev, rank, world, dist, steps=args.finetune_steps, warmup=3, graph=True, allreduce=False)
        rec["ms_per_step_without_allreduce"] = no_ar["ms_per_step"]
        rec["allreduce_cost_ms"] = rec["ms_per_step"] - no_ar["ms_per_step"]
    return rec


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_gbs(log2_numel, reps, specs=SWEEP):
    """The CPU oracle port (table build once, then divide/index/lookup/multiply per call) on all host threads."""
    import numpy as np
    from oracle import oracle as O

    # all host threads, whatever the launcher exported: torch.distributed.run sets OMP_NUM_THREADS=1 for its workers
    O.set_num_threads(os.cpu_count() or 1)
    n = 1 << log2_numel
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(n, dtype=np.float32) * 4.0)
    xb = (x.view(np.uint32) >> 16).astype(np.uint16)  # bf16 bits (truncation is fine for a timing input)
    mods = [O.FakeQuant(s) for s in specs]
    for m in mods[:1]:
        m(xb, (n,))
    t0 = time.perf_counter()
    for _ in range(reps):
        for m in mods:
            m(xb, (n,))
    dt = time.perf_counter() - t0
    return 4.0 * n * len(mods) * reps / dt / 1e9, O.num_threads(), dt


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path, as ported in oracle/ (OpenMP, all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = min(args.log2_numel, 26)   # 2^26 elements x 8 specs per step: ~0.1-0.2 s of CPU work per step
    for _ in range(args.warmup):
        cpu_port_gbs(L, 1, SWEEP[:1])
    t0 = time.perf_counter()
    gbs, threads, dt = cpu_port_gbs(L, args.steps)
    sample = f"{args.steps} steps x {len(SWEEP)} specs x 2^{L} bf16 elements (bounded sample of the GPU workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "standalone quantize sweep (configs[2]) -- CPU port of the reference path",
                   "specs": SWEEP, "log2_numel_per_step": L},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(local):
    """Pin this rank to the host cores of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated
    (first touch decides where the pages live): with one process per GPU the host <-> device legs then stay on the
    GPU's own socket instead of all ranks sharing node 0's memory controllers.  Returns a short description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"gpu {local} ({bdf}): no NUMA affinity reported"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"gpu {local} ({bdf}) -> NUMA node {node}, {len(allowed)} cores"
    except Exception as e:  # sysfs layout differs, container without the files, ...: run unbound
        return f"unbound ({type(e).__name__})"


def run_gpu(args):
    import torch
    import quantized_training as qt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa = bind_to_gpu_numa(local) if world > 1 else "single process: unbound"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = 1 << args.log2_numel
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.empty(n, device=dev, dtype=torch.bfloat16)
    chunk = 1 << 26
    for i in range(0, n, chunk):  # randn * 4: exercises normal, subnormal and (for fp4/int4) saturating branches
        x[i:i + chunk] = (torch.randn(min(chunk, n - i), device=dev, generator=gen) * 4.0).to(torch.bfloat16)
    mods = [qt.FusedAmaxObsFakeQuantize(s, device=dev) for s in SWEEP]
    y = torch.empty_like(x)
    fmts = [(m._fmt, m.lut) for m in mods]
    unit = torch.ones(1, device=dev)

    def step():
        for f, lut in fmts:  # straight through the C ABI: y = fq(x), output buffer reused
            qt._C.fq_forward(x, y, 1, 1, n, f, unit, None, lut)

    # the clock sampler starts BEFORE warm-up: nvidia-smi's start-up (NVML init) stalls the GPU for a few ms and
    # must not land inside the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            step()
            torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in SWEEP]
          for _ in range(args.steps)]
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        for i, (f, lut) in enumerate(fmts):
            ev[k][i][0].record()
            qt._C.fq_forward(x, y, 1, 1, n, f, unit, None, lut)
            ev[k][i][1].record()
    t1.record()
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    bytes_per_launch = 4.0 * n  # 2 B read + 2 B written per element
    launches = args.steps * len(SWEEP)
    value = bytes_per_launch * launches * world / (elapsed_ms * 1e-3) / 1e9
    per_spec_ms = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(args.steps)) / args.steps
                   for i in range(len(SWEEP))]  # mean launch duration per spec over the timed region
    kernel_ms = sum(per_spec_ms) / len(SWEEP)
    slowest = max(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(args.steps) for i in range(len(SWEEP)))
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9

    # ---- side measurements (not part of `value`): other call shapes of the same kernels, same tensor
    def timed(fn, reps=5):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    extra = {}
    if rank == 0 and not args.no_extras:
        nn_ = min(n, 1 << 28)
        xs, ys = x[:nn_], y[:nn_]
        sc = torch.full((1,), 0.0123, device=dev)
        hist = torch.zeros(1, device=dev)
        for name in ("e4m3", "fp8_e4m3", "posit8_1", "int8"):
            m = qt.FusedAmaxObsFakeQuantize(name, device=dev)
            ms = timed(lambda: qt._C.fq_forward(xs, ys, 1, 1, nn_, m._fmt, sc, hist, m.lut))
            extra[f"{name} bf16 per-tensor scale + amax"] = 4.0 * nn_ / (ms * 1e-3) / 1e9
            ms = timed(lambda: qt._C.fq_forward(xs, ys, 1, 1, nn_, m._fmt, unit, None, None))
            extra[f"{name} bf16 bare, direct bitwise path"] = 4.0 * nn_ / (ms * 1e-3) / 1e9
        m = qt.FusedAmaxObsFakeQuantize("e4m3", device=dev)
        codes = torch.empty(nn_, dtype=torch.uint8, device=dev)
        ms = timed(lambda: qt._C.quantize_codes(xs, codes, m._fmt, unit, None, m.lut))
        extra["e4m3 bf16 -> fp8 codes (3 B/element)"] = 3.0 * nn_ / (ms * 1e-3) / 1e9
        del codes
        rows = nn_ // 4096
        scr = torch.rand(rows, device=dev) * 0.05 + 0.01
        scc = torch.rand(4096, device=dev) * 0.05 + 0.01
        m = qt.FusedAmaxObsFakeQuantize("posit8_1", device=dev)
        ms = timed(lambda: qt._C.fq_forward(xs, ys, 1, rows, 4096, m._fmt, scr, torch.zeros(rows, device=dev), m.lut))
        extra["posit8_1 bf16 per-channel ax=0 [N/4096,4096] + amax"] = 4.0 * nn_ / (ms * 1e-3) / 1e9
        ms = timed(lambda: qt._C.fq_forward(xs, ys, rows, 4096, 1, m._fmt, scc, torch.zeros(4096, device=dev), m.lut))
        extra["posit8_1 bf16 per-channel ax=-1 [N/4096,4096] + amax"] = 4.0 * nn_ / (ms * 1e-3) / 1e9
        # block-scaled qschemes through the module API (scale buffer written per block: + 4 / bs bytes per element)
        x2 = xs.view(rows, 4096)
        for spec, ax in (("fp4_e2m1,qs=microscaling,bs=32", -1), ("fp8_e4m3,qs=microscaling,bs=32", -1),
                         ("int6,qs=microscaling,bs=64,scale=fp8_e5m3", -1), ("int6,qs=microscaling,bs=64,scale=fp8_e5m3", 0)):
            for pow2 in ((True,) if "fp" in spec.split(",")[0] else (False,)):
                qs = qt.QuantizationSpec.from_str(spec + f",ax={ax}")
                m = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), force_scale_power_of_two=pow2, device=dev)
                ms = timed(lambda: m(x2))
                extra[f"{spec} ax={ax}{' pow2' if pow2 else ''} bf16 [N/4096,4096]"] = \
                    (4.0 + 4.0 / qs.block_size) * nn_ / (ms * 1e-3) / 1e9
        qs = qt.QuantizationSpec.from_str("uint4,qs=group_wise_affine,bs=64,ax=-1")
        m = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), device=dev)
        ms = timed(lambda: m(x2))
        extra["uint4,qs=group_wise_affine,bs=64 ax=-1 bf16 [N/4096,4096] (two-kernel generic path)"] = \
            (4.0 + 8.0 / 64) * nn_ / (ms * 1e-3) / 1e9
        del x2
        ms = timed(lambda: qt._C.amax(xs, 1, 1, nn_, hist))
        extra["amax only bf16 (read bytes)"] = 2.0 * nn_ / (ms * 1e-3) / 1e9
        x32 = xs[: nn_ // 2].float()
        y32 = torch.empty_like(x32)
        for name in ("posit8_1", "int8"):
            m = qt.FusedAmaxObsFakeQuantize(name, device=dev)
            ms = timed(lambda: qt._C.fq_forward(x32, y32, 1, 1, x32.numel(), m._fmt, unit, None, m.lut))
            extra[f"{name} fp32 bare"] = 8.0 * x32.numel() / (ms * 1e-3) / 1e9
            ms = timed(lambda: qt._C.fq_forward(x32, y32, 1, 1, x32.numel(), m._fmt, sc, hist, m.lut))
            extra[f"{name} fp32 per-tensor scale + amax"] = 8.0 * x32.numel() / (ms * 1e-3) / 1e9
        del x32, y32

    # ---- end to end through the public module API, host buffers, copies inside the timed region
    ne = 1 << min(args.log2_numel, args.e2e_log2_numel)
    xh = torch.empty(ne, dtype=torch.bfloat16).pin_memory()
    xh.copy_(x[:ne])
    yhs = [torch.empty(ne, dtype=torch.bfloat16).pin_memory() for _ in SWEEP]  # one result buffer per spec
    e2e_steps = max(1, min(args.steps, 5))

    from quantized_training.host_io import HostPipeline
    pipe = HostPipeline(dev, torch.bfloat16, chunk_elems=1 << args.e2e_log2_chunk, depth=4)

    def e2e_step():
        # the sweep applies 8 specs to ONE host tensor: each chunk is uploaded once, quantized 8 times and the 8
        # results are downloaded (H2D | kernels | D2H overlapped on 4 streams, pinned host memory both ways)
        pipe.run_many(mods, xh, yhs)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - w0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = 4.0 * ne * len(SWEEP) * e2e_steps * world / e2e_s / 1e9

    del pipe, xh, yhs, x, y
    torch.cuda.empty_cache()
    llama = None
    if not args.no_llama:
        llama = run_llama(args, dev, rank, world, dist, barrier)
    finetune = None
    if not args.no_finetune:
        finetune = run_finetune(args, dev, rank, world, dist)
    sizes = gemm = None
    if rank == 0 and not args.no_extras:
        sizes = run_size_sweep(dev, qt)
        gemm = run_gemm(dev)
    barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            gbs, threads, dt = cpu_port_gbs(24, 2)
            reps = max(1, min(2000, int(12.0 / max(dt / 2, 1e-3))))  # about 12 s of CPU work
            gbs, threads, dt = cpu_port_gbs(24, reps)
            cpu = {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{reps} x {len(SWEEP)} specs x 2^24 bf16 elements, {dt:.1f} s"}
        tr = measured_traffic()
        traffic_per_launch, traffic_src = None, None
        if tr:  # measured DRAM bytes / algorithmic bytes of the captured launch, applied to this launch size
            traffic_per_launch = bytes_per_launch * tr["dram_bytes_per_launch"] / tr["algorithmic_bytes_per_launch"]
            traffic_src = tr["source"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "standalone quantize sweep (BASELINE configs[2]): bf16 tensor resident in HBM, "
                                   "bare specs " + "/".join(SWEEP),
                       "log2_numel_per_gpu": args.log2_numel, "launches_per_step": len(SWEEP),
                       "l2": "input+output 4*2^L bytes per launch >> 126 MB L2 (no flush needed)",
                       "parallelism": f"dp{world} (independent shards, no collective)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_per_launch, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "fq_flat_kernel",
                         "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": kernel_ms,
                         "max_launch_ms": slowest,
                         "per_spec_GBps": {s: bytes_per_launch / (ms * 1e-3) / 1e9 for s, ms in zip(SWEEP, per_spec_ms)}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * ne,
                    "d2h_bytes_per_step": 2 * ne * len(SWEEP), "log2_numel": ne.bit_length() - 1, "host_affinity_rank0": numa,
                    "api": "quantized_training.host_io.HostPipeline.run_many(8 modules, pinned host tensor, 8 pinned "
                           f"host results): {2 << (args.e2e_log2_chunk - 20)} MB chunks, each uploaded once, H2D / "
                           "8 kernels / 8 D2H overlapped on 4 streams; value counts the algorithmic read+write bytes "
                           "of the 8 specs like the device-timed figure"},
            "gpu_launches": launches, "clocks": clocks, "other_shapes_GBps": extra, "size_sweep": sizes,
            "gemm": gemm, "llama_forward": llama, "finetune": finetune,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-numel", type=int, default=32)
    ap.add_argument("--e2e-log2-numel", type=int, default=26)
    ap.add_argument("--e2e-log2-chunk", type=int, default=22, help="log2 elements per staged chunk of the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-llama", action="store_true")
    ap.add_argument("--llama-steps", type=int, default=10)
    ap.add_argument("--no-finetune", action="store_true")
    ap.add_argument("--finetune-steps", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
