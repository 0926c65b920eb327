#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (jeffreyyu0602/quantized-training, mounted read-only at
/root/reference) on CPU.

The reference package cannot be imported as shipped in this image (graphviz,
accelerate, peft, torch.ao.quantization.quantizer are absent), so its numerics
files are imported one by one through a synthetic package.
Nothing is copied: the reference code is *executed* and only its inputs and
outputs are stored.

Run (container with /root/reference only; never on the GPU box):

    python tests/golden/gen_golden.py

Outputs (all small, committed):
    qmaps.npz      get_quantization_map(dtype) for every dtype string below,
                   the total function bf16-bits -> bf16-bits  (fake_quantize.py:31-95)
    pbits.npz      quantize_to_posit(..., return_pbits=True) codes  (posit.py:60-65)
    vmap32.npz     decomposed.vmap on fp32 inputs covering every (top16, sticky)
                   pair  (decomposed.py:146-163)
    fq_cases.npz   FusedAmaxObsFakeQuantize call sequences (delayed scaling,
                   per-tensor / per-channel, bf16 / fp32) with the buffers after
                   every call  (fake_quantize.py:202-248)
    manifest.json  what is in the files above
"""
import abc
import importlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/quantized_training"

QMAP_DTYPES = [
    # north-star list
    "int4", "int8", "e4m3", "e5m2", "fp8_e4m3", "fp8_e5m2", "fp6_e3m2", "fp6_e2m3",
    "fp4_e2m1", "posit8_0", "posit8_1", "posit8_2", "posit16_1",
    # grammar coverage / generic-parameter stress
    "int2", "int3", "int5", "int6", "int7", "int16", "uint4", "uint8", "E4M3", "fp8.e5m2",
    "fp8_e3m4", "fp8_e2m5", "fp7_e3m3", "fp5_e2m2", "fp5_e3m1", "fp8_e5m3", "fp8_e4m4",
    "posit4_0", "posit6_1", "posit8_3", "posit10_1", "posit16_2", "posit12_0",
]
PBITS_DTYPES = [(8, 0), (8, 1), (8, 2), (6, 1), (16, 1)]
VMAP32_DTYPES = ["int8", "e4m3", "posit8_1", "fp6_e3m2", "fp8_e4m3"]


def load_reference():
    import transformers  # noqa: F401  (must be imported before accelerate is stubbed)

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class QuantizationSpecBase(abc.ABC):
        ...

    stub("torch.ao.quantization.quantizer", QuantizationSpecBase=QuantizationSpecBase, EdgeOrNode=object)
    stub("torch.ao.quantization.quantizer.quantizer", QuantizationSpecBase=QuantizationSpecBase)
    stub("accelerate", dispatch_model=lambda *a, **k: None)

    class _L(nn.Module):
        ...

    p, t = stub("peft"), stub("peft.tuners")
    p.tuners = t
    t.lora = stub("peft.tuners.lora", Linear=_L)
    stub("peft.utils")
    stub("peft.utils.other", transpose=lambda w, f: w.T if f else w)
    pkg = types.ModuleType("quantized_training")
    pkg.__path__ = [REF]
    sys.modules["quantized_training"] = pkg
    Q = importlib.import_module("quantized_training.quantizer.quantizer").QScheme
    pkg.per_tensor_symmetric, pkg.per_channel_symmetric = Q.PER_TENSOR_SYMMETRIC, Q.PER_CHANNEL_SYMMETRIC
    pkg.microscaling, pkg.group_wise_affine = Q.MICROSCALING, Q.GROUP_WISE_AFFINE
    ref = types.SimpleNamespace()
    ref.fq = importlib.import_module("quantized_training.fake_quantize")
    ref.posit = importlib.import_module("quantized_training.posit")
    ref.decomposed = importlib.import_module("quantized_training.decomposed")
    ref.qconfig = importlib.import_module("quantized_training.qconfig")
    ref.quantizer = importlib.import_module("quantized_training.quantizer.quantizer")
    return ref


def bits16(t):
    """bf16 tensor -> uint16 numpy."""
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def bits32(t):
    return t.contiguous().view(torch.int32).numpy().view(np.uint32).copy()


def tensor_bits(t):
    if t.dtype == torch.bfloat16:
        return bits16(t)
    assert t.dtype == torch.float32, t.dtype
    return bits32(t)


def gen_qmaps(ref):
    out = {}
    for d in QMAP_DTYPES:
        m = ref.fq.get_quantization_map(d)
        assert m.dtype == torch.bfloat16 and m.numel() == 65536, (d, m.dtype, m.shape)
        out[d] = bits16(m)
    return out


def gen_pbits(ref):
    vals = torch.arange(2 ** 16, dtype=torch.int16).view(torch.bfloat16)
    out = {}
    for n, es in PBITS_DTYPES:
        _, pb = ref.posit.quantize_to_posit(vals, n, es, round_to_even=True, return_pbits=True)
        out[f"posit{n}_{es}"] = pb.numpy().astype(np.int32)
    return out


def adversarial_fp32_bits():
    """Every bf16 top-half with low halves that exercise the sticky bit."""
    top = np.arange(65536, dtype=np.uint32) << 16
    lows = np.array([0x0000, 0x0001, 0x4000, 0x7FFF, 0x8000, 0x8001, 0xC000, 0xFFFF], dtype=np.uint32)
    return (top[:, None] | lows[None, :]).reshape(-1)


def gen_vmap32(ref):
    xb = adversarial_fp32_bits()
    x = torch.from_numpy(xb.view(np.int32).copy()).view(torch.float32)
    out = {"x_bits": xb}
    for d in VMAP32_DTYPES:
        qmap = ref.fq.get_quantization_map(d)
        y = ref.decomposed.vmap(x, qmap)
        assert y.dtype == torch.float32
        yb = bits32(y)
        assert np.all((yb & 0xFFFF) == 0), "vmap output on fp32 must be bf16-exact"
        out[d] = (yb >> 16).astype(np.uint16)
    return out


# (name, spec string, input dtype, shape, force_pow2, per-call input recipe)
FQ_CASES = [
    ("e4m3_bare_bf16", "e4m3", "bf16", (37, 129), False),
    ("posit8_1_bare_bf16", "posit8_1", "bf16", (5, 33, 65), False),
    ("posit8_1_bare_fp32", "posit8_1", "fp32", (37, 129), False),
    ("int8_bare_fp32", "int8", "fp32", (1000,), False),
    ("fp4_e2m1_bare_bf16", "fp4_e2m1", "bf16", (1023,), False),
    ("int8_pt_bf16", "int8,qs=per_tensor_symmetric", "bf16", (37, 129), False),
    ("int8_pt_fp32", "int8,qs=per_tensor_symmetric,ahl=3", "fp32", (37, 129), False),
    ("int4_pt_bf16", "int4,qs=per_tensor_symmetric,ahl=2", "bf16", (64, 64), False),
    ("fp8e4m3_pt_bf16", "fp8_e4m3,qs=per_tensor_symmetric,ahl=4", "bf16", (3, 50, 70), False),
    ("e5m2_err_bf16", "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10", "bf16", (16, 128, 24), False),
    ("posit8_1_pt64_bf16", "posit8_1,qs=per_tensor_symmetric,qmax=64,ahl=10", "bf16", (33, 77), False),
    ("e4m3_pt_ahl1_bf16", "fp8_e4m3,qs=per_tensor_symmetric,qmax=448,ahl=1", "bf16", (129,), False),
    ("int8_pc0_bf16", "int8,qs=per_channel_symmetric,ax=0", "bf16", (24, 130), False),
    ("int4_pc0_fp32", "int4,qs=per_channel_symmetric,ax=0,ahl=3", "fp32", (24, 130), False),
    ("posit8_1_pclast_bf16", "posit8_1,qs=per_channel_symmetric,ax=-1,qmax=64,ahl=2", "bf16", (7, 19, 40), False),
    ("fp6_e3m2_pcmid_bf16", "fp6_e3m2,qs=per_channel_symmetric,ax=1,ahl=4", "bf16", (6, 12, 34), False),
    ("int8_pt_pow2_bf16", "int8,qs=per_tensor_symmetric,ahl=4", "bf16", (37, 129), True),
    ("e4m3_pt_pow2_fp32", "fp8_e4m3,qs=per_tensor_symmetric,ahl=2", "fp32", (37, 129), True),
]
N_CALLS = 7


def make_input(gen, shape, dtype, call):
    """Inputs whose magnitude changes call to call so the delayed scale moves;
    call 3 carries NaN/Inf, call 4 is all zeros, call 5 has subnormal-range values."""
    scales = [1.0, 37.5, 0.004, 900.0, 0.0, 3.0e-5, 5.0]
    x = torch.randn(shape, generator=gen, dtype=torch.float32) * scales[call]
    flat = x.view(-1)
    if call == 3 and flat.numel() >= 8:
        flat[1] = float("inf")
        flat[5] = float("nan")
        flat[7] = -float("inf")
    if call == 5:
        flat[0] = 1e-40  # fp32 subnormal (flushes to 0 in bf16 conversion or stays for fp32)
        flat[2] = -0.0
    return x.to(torch.bfloat16 if dtype == "bf16" else torch.float32)


def gen_fq_cases(ref):
    out, manifest = {}, []
    for ci, (name, spec, dtype, shape, pow2) in enumerate(FQ_CASES):
        gen = torch.Generator().manual_seed(1234 + ci)
        ctr = ref.qconfig._create_fake_quant(spec, False, pow2)
        mod = ctr()
        entry = {"name": name, "spec": spec, "dtype": dtype, "shape": list(shape),
                 "force_scale_power_of_two": pow2, "calls": N_CALLS,
                 "quant_max": mod.quant_max, "amax_history_len": mod.amax_history_len,
                 "ch_axis": mod.ch_axis,
                 "qscheme": None if mod.qscheme is None else mod.qscheme.value}
        for k in range(N_CALLS):
            x = make_input(gen, shape, dtype, k)
            y = mod(x.clone())
            assert y.dtype == x.dtype and y.shape == x.shape and y.is_contiguous()
            out[f"{name}/x{k}"] = tensor_bits(x)
            out[f"{name}/y{k}"] = tensor_bits(y)
            out[f"{name}/scale{k}"] = bits32(mod.scale.reshape(-1))
            out[f"{name}/hist{k}"] = bits32(mod.amax_history.reshape(-1))
            entry.setdefault("scale_shape", []).append(list(mod.scale.shape))
            entry.setdefault("hist_shape", []).append(list(mod.amax_history.shape))
        # observer off / fake-quant off behaviour on the last state
        mod.disable_observer()
        x = make_input(gen, shape, dtype, 1)
        y = mod(x.clone())
        out[f"{name}/x_obsoff"] = tensor_bits(x)
        out[f"{name}/y_obsoff"] = tensor_bits(y)
        out[f"{name}/scale_obsoff"] = bits32(mod.scale.reshape(-1))
        manifest.append(entry)
    return out, manifest


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = load_reference()
    qmaps = gen_qmaps(ref)
    np.savez_compressed(os.path.join(HERE, "qmaps.npz"), **qmaps)
    pbits = gen_pbits(ref)
    np.savez_compressed(os.path.join(HERE, "pbits.npz"), **pbits)
    vm = gen_vmap32(ref)
    np.savez_compressed(os.path.join(HERE, "vmap32.npz"), **vm)
    fq, fq_manifest = gen_fq_cases(ref)
    np.savez_compressed(os.path.join(HERE, "fq_cases.npz"), **fq)
    manifest = {
        "generator": "tests/golden/gen_golden.py",
        "reference": "jeffreyyu0602/quantized-training @ /root/reference (CPU, torch %s)" % torch.__version__,
        "qmaps": QMAP_DTYPES,
        "pbits": [f"posit{n}_{es}" for n, es in PBITS_DTYPES],
        "vmap32": VMAP32_DTYPES,
        "fq_cases": fq_manifest,
    }
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for fn in ["qmaps.npz", "pbits.npz", "vmap32.npz", "fq_cases.npz", "manifest.json"]:
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
