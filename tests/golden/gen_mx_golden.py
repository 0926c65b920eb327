#!/usr/bin/env python
"""Golden fixtures for the block-scaled qschemes (microscaling, group_wise_affine), produced by
running the UNMODIFIED reference on CPU through the import shim of gen_golden.py.

    python tests/golden/gen_mx_golden.py        (build container only: needs /root/reference)

Outputs:
    mx_cases.npz      FusedAmaxObsFakeQuantize(qscheme=microscaling / group_wise_affine) forward on
                      seeded inputs: x, y, scale (and zero_point) bit patterns per case
                      (fake_quantize.py:98-194, decomposed.py:366-448, mx_utils.py:62-121)
    mx_scale.npz      calculate_mx_qparam as a function of the block amax alone: every positive bf16
                      value, and fp32 values around every power of two, with and without
                      force_scale_power_of_two (the floor(log2()) of the reference runs in the tensor's
                      dtype, so the result is not simply the exponent field)
    mx_manifest.json  what is in the files above
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402

# name, tensor dtype, shape, ch_axis, block_size, element dtype, force_pow2, scale_dtype, magnitude
MX_CASES = [
    ("mx_int6_last_bs32_e5m3", "bf16", (16, 256), -1, 32, "int6", False, "fp8_e5m3", 1.0),
    ("mx_int6_last_bs64_e5m3", "bf16", (4, 6, 128), -1, 64, "int6", False, "fp8_e5m3", 4.0),
    ("mx_int6_ax2_bs64_e5m3", "bf16", (3, 160, 64), -2, 64, "int6", False, "fp8_e5m3", 1.0),
    ("mx_int6_2d_bs16_e5m3", "bf16", (2, 3, 48, 64), (-2, -1), 16, "int6", False, "fp8_e5m3", 1.0),
    ("mx_int4_2d_ragged_pow2", "bf16", (2, 40, 50), (-2, -1), 16, "int4", True, None, 1e-3),
    ("mx_fp4_last_bs32_pow2", "bf16", (32, 512), -1, 32, "fp4_e2m1", True, None, 1.0),
    ("mx_fp6_last_bs32_pow2", "bf16", (8, 512), -1, 32, "fp6_e3m2", True, None, 30.0),
    ("mx_fp8_last_bs32_pow2", "bf16", (8, 512), -1, 32, "fp8_e4m3", True, None, 1e4),
    ("mx_fp4_last_ragged", "bf16", (5, 70), -1, 32, "fp4_e2m1", False, None, 1.0),
    ("mx_fp4_mid_ragged_e5m3", "bf16", (3, 70, 24), 1, 32, "fp4_e2m1", False, "fp8_e5m3", 3.0),
    ("mx_posit8_ax0_bs4", "bf16", (10, 33), 0, 4, "posit8_1", False, None, 30.0),
    ("mx_int8_short_axis", "bf16", (7,), 0, 32, "int8", False, None, 1.0),
    ("mx_int8_bs8_odd_inner", "bf16", (6, 40, 3), 1, 8, "int8", False, None, 2.0),
    ("mx_int6_last_bs32_e5m3_f32", "f32", (16, 256), -1, 32, "int6", False, "fp8_e5m3", 1.0),
    ("mx_fp4_ax2_pow2_f32", "f32", (3, 70, 16), -2, 32, "fp4_e2m1", True, None, 3.0),
    ("mx_fp6_last_pow2_f32", "f32", (8, 512), -1, 64, "fp6_e3m2", True, None, 1.0),
    ("mx_posit8_ax0_bs4_f32", "f32", (10, 33), 0, 4, "posit8_1", False, None, 30.0),
    ("mx_int6_2d_bs16_f32", "f32", (2, 40, 50), (-2, -1), 16, "int6", False, "fp8_e5m3", 1.0),
]
# name, tensor dtype, shape, ch_axis, block_size, element dtype, scale_dtype, magnitude
GWA_CASES = [
    ("gwa_uint2_last_bs64_e5m3", "bf16", (8, 256), -1, 64, "uint2", "fp8_e5m3", 1.0),
    ("gwa_uint2_ax2_bs64_e5m3", "bf16", (2, 160, 32), -2, 64, "uint2", "fp8_e5m3", 1.0),
    ("gwa_uint4_last_ragged", "bf16", (4, 70), -1, 32, "uint4", None, 10.0),
    ("gwa_uint4_2d_ragged", "bf16", (2, 40, 50), (-2, -1), 16, "uint4", None, 1.0),
    ("gwa_int4_last_bs32", "bf16", (16, 128), -1, 32, "int4", None, 5.0),
    ("gwa_uint2_ax2_bs64_e5m3_f32", "f32", (2, 160, 32), -2, 64, "uint2", "fp8_e5m3", 1.0),
    ("gwa_int4_last_ragged_f32", "f32", (4, 70), -1, 32, "int4", None, 10.0),
]


def make_x(gen, shape, dtype, mag, affine=False):
    """Heavy-tailed values (block maxima differ by orders of magnitude), an all-zero block, NaN / Inf,
    a constant block (affine range 0) and bf16-subnormal entries."""
    td = torch.bfloat16 if dtype == "bf16" else torch.float32
    x = torch.randn(shape, generator=gen) * mag * torch.exp2(torch.randint(-6, 6, shape, generator=gen).float())
    if affine:
        x = x + 0.3 * mag
    flat = x.view(-1)
    n = flat.numel()
    if n >= 512:
        flat[64:128] = 0.0
        flat[130] = float("nan")
        flat[200] = float("inf")
        flat[260] = -float("inf")
        flat[320:384] = 1.5
        flat[400] = 1e-39
        flat[401] = -3e-40
    return x.to(td)


def gen_cases(ref):
    FQ = ref.fq.FusedAmaxObsFakeQuantize
    QS = ref.quantizer.QScheme
    out, manifest = {}, []
    for ci, (name, dtype, shape, ax, bs, el, pow2, sdt, mag) in enumerate(MX_CASES):
        gen = torch.Generator().manual_seed(4000 + ci)
        qmin, qmax = ref.quantizer.get_quant_min_max(el)
        mod = FQ(el, qscheme=QS.MICROSCALING, quant_min=float(qmin), quant_max=float(qmax), ch_axis=ax,
                 block_size=bs, scale_dtype=sdt, force_scale_power_of_two=pow2)
        x = make_x(gen, shape, dtype, mag)
        y = mod(x.clone())
        assert y.shape == x.shape and y.dtype == x.dtype
        out[f"{name}/x"] = G.tensor_bits(x)
        out[f"{name}/y"] = G.tensor_bits(y)
        out[f"{name}/scale"] = G.bits32(mod.scale.reshape(-1))
        manifest.append({"name": name, "qscheme": "microscaling", "dtype": dtype, "shape": list(shape),
                         "ch_axis": ax, "block_size": bs, "element": el, "force_scale_power_of_two": pow2,
                         "scale_dtype": sdt, "quant_min": float(qmin), "quant_max": float(qmax),
                         "scale_shape": list(mod.scale.shape)})
    for ci, (name, dtype, shape, ax, bs, el, sdt, mag) in enumerate(GWA_CASES):
        gen = torch.Generator().manual_seed(5000 + ci)
        qmin, qmax = ref.quantizer.get_quant_min_max(el)
        mod = FQ(el, qscheme=QS.GROUP_WISE_AFFINE, quant_min=float(qmin), quant_max=float(qmax), ch_axis=ax,
                 block_size=bs, scale_dtype=sdt)
        x = make_x(gen, shape, dtype, mag, affine=True)
        y = mod(x.clone())
        out[f"{name}/x"] = G.tensor_bits(x)
        out[f"{name}/y"] = G.tensor_bits(y)
        out[f"{name}/scale"] = G.bits32(mod.scale.reshape(-1))
        out[f"{name}/zero_point"] = G.bits32(mod.zero_point.reshape(-1))
        manifest.append({"name": name, "qscheme": "group_wise_affine", "dtype": dtype, "shape": list(shape),
                         "ch_axis": ax, "block_size": bs, "element": el, "scale_dtype": sdt,
                         "quant_min": float(qmin), "quant_max": float(qmax),
                         "scale_shape": list(mod.scale.shape)})
    return out, manifest


def f32_probe_bits():
    """fp32 amax values around every power of two, plus subnormals and specials."""
    tails = [0, 1, 2, 0x10, 0x100, 0x1000, 0x10000, 0x100000, 0x200000, 0x400000, 0x600000, 0x700000,
             0x780000, 0x7F0000, 0x7FF000, 0x7FFF00, 0x7FFFF0, 0x7FFFF8, 0x7FFFFC, 0x7FFFFE, 0x7FFFFF]
    v = [(e << 23) | t for e in range(0, 255) for t in tails]
    v += [0x7F800000, 0x7FC00000]
    return np.array(sorted(set(v)), dtype=np.uint32)


def gen_scale_fn(ref):
    D = ref.decomposed
    out = {}
    allb = torch.arange(0, 0x8000, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).reshape(-1, 1)
    fb = f32_probe_bits()
    allf = torch.from_numpy(fb.view(np.int32).copy()).view(torch.float32).reshape(-1, 1)
    out["f32_amax_bits"] = fb
    sq = ref.fq.get_quantization_map("fp8_e5m3")
    for qmax in (448.0, 31.0, 6.0, 7.5, 32767.0):
        tag = str(qmax)
        out[f"bf16/pow2/{tag}"] = G.bits16(D.calculate_mx_qparam(allb, [-1], 1, qmax, True).reshape(-1))
        out[f"bf16/amax/{tag}"] = G.bits16(D.calculate_mx_qparam(allb, [-1], 1, qmax, False).reshape(-1))
        out[f"bf16/e5m3/{tag}"] = G.bits16(D.calculate_mx_qparam(allb, [-1], 1, qmax, False, sq).reshape(-1))
        out[f"f32/pow2/{tag}"] = G.bits32(D.calculate_mx_qparam(allf, [-1], 1, qmax, True).reshape(-1))
        out[f"f32/amax/{tag}"] = G.bits32(D.calculate_mx_qparam(allf, [-1], 1, qmax, False).reshape(-1))
        out[f"f32/e5m3/{tag}"] = G.bits32(D.calculate_mx_qparam(allf, [-1], 1, qmax, False, sq).reshape(-1))
    return out


def gen_ops(ref):
    """torch.ops.quantized_ops.{vmap, quantize, dequantize} of the reference on CPU (decomposed.py:143-262):
    inputs, parameters, tables and outputs of a few calls that cover block grids, type promotion, zero points,
    both tables of dequantize and codebooks this library has no bitwise rounder for (NF4, a random table)."""
    D = ref.decomposed
    gen = torch.Generator().manual_seed(77)
    out, manifest = {}, []

    def table(name):
        t = ref.fq.get_quantization_map(name)
        if isinstance(t, tuple):
            t = t[1][t[0]]
        return t.to(torch.bfloat16)

    rnd_table = torch.randint(0, 65536, (65536,), generator=gen, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    tabs = {"int6": table("int6"), "nf4": table("nf4"), "fp8_e4m3": table("fp8_e4m3"), "random": rnd_table}
    for k, t in tabs.items():
        out[f"table/{k}"] = G.bits16(t)

    def rec(name, op, x, scale, zp, axes, bs, ta, tb, y):
        out[f"{name}/x"] = G.tensor_bits(x)
        out[f"{name}/scale"] = G.tensor_bits(scale)
        if zp is not None:
            out[f"{name}/zp"] = G.tensor_bits(zp)
        out[f"{name}/y"] = G.tensor_bits(y)
        dn = lambda t: "bf16" if t.dtype == torch.bfloat16 else "f32"
        manifest.append({"name": name, "op": op, "x_dtype": dn(x), "x_shape": list(x.shape), "scale_dtype": dn(scale),
                         "scale_shape": list(scale.shape), "zp": zp is not None, "zp_dtype": None if zp is None else dn(zp),
                         "axes": axes, "block_size": bs, "table_a": ta, "table_b": tb, "y_dtype": dn(y)})

    x = (torch.randn(8, 128, generator=gen) * 3).to(torch.bfloat16)
    x[0, :5] = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.0, -0.0]).to(torch.bfloat16)
    s = (torch.rand(8, 4, generator=gen) * 0.2 + 0.01).to(torch.bfloat16)
    rec("q_bf16_blocks_int6", "quantize", x, s, None, [-1], 32, "int6", None, D.quantize(x, s, None, [-1], 32, tabs["int6"]))
    rec("q_bf16_blocks_nf4", "quantize", x, s * 20, None, [-1], 32, "nf4", None, D.quantize(x, s * 20, None, [-1], 32, tabs["nf4"]))
    s32 = s.float() * 1.00123
    zp32 = torch.rand(8, 4, generator=gen) * 4 - 2
    rec("q_bf16_f32scale_zp", "quantize", x, s32, zp32, [-1], 32, "int6", None, D.quantize(x, s32, zp32, [-1], 32, tabs["int6"]))
    xf = torch.randn(5, 70, generator=gen) * 40
    s0 = torch.tensor(0.37)
    rec("q_f32_scalar_fp8", "quantize", xf, s0, None, None, None, "fp8_e4m3", None, D.quantize(xf, s0, None, None, None, tabs["fp8_e4m3"]))
    x3 = (torch.randn(2, 70, 16, generator=gen) * 2).to(torch.bfloat16)
    sg = (torch.rand(2, 3, 16, generator=gen) * 0.5 + 0.05).to(torch.bfloat16)
    zg = (torch.rand(2, 3, 16, generator=gen) * 3).to(torch.bfloat16)
    rec("dq_bf16_ax1_tables", "dequantize", x3, sg, zg, [-2], 32, "int6", "fp8_e4m3",
        D.dequantize(x3, sg, zg, [-2], 32, tabs["int6"], tabs["fp8_e4m3"]))
    rec("dq_bf16_scalar", "dequantize", x3, torch.tensor(0.25, dtype=torch.bfloat16), None, None, None, None, None,
        D.dequantize(x3, torch.tensor(0.25, dtype=torch.bfloat16)))
    x2 = (torch.randn(40, 48, generator=gen)).to(torch.bfloat16)
    s2 = (torch.rand(3, 3, generator=gen) * 0.3 + 0.02).to(torch.bfloat16)
    rec("q_bf16_2d_blocks", "quantize", x2, s2, None, [-2, -1], 16, "int6", None, D.quantize(x2, s2, None, [-2, -1], 16, tabs["int6"]))
    allb = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    rec("vmap_bf16_random", "vmap", allb, torch.ones(1, dtype=torch.bfloat16), None, None, None, "random", None, D.vmap(allb, tabs["random"]))
    xr = torch.randn(3001, generator=gen) * 100
    rec("vmap_f32_random", "vmap", xr, torch.ones(1), None, None, None, "random", None, D.vmap(xr, tabs["random"]))
    return out, manifest


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = G.load_reference()
    cases, manifest = gen_cases(ref)
    np.savez_compressed(os.path.join(HERE, "mx_cases.npz"), **cases)
    sc = gen_scale_fn(ref)
    np.savez_compressed(os.path.join(HERE, "mx_scale.npz"), **sc)
    ops, ops_manifest = gen_ops(ref)
    np.savez_compressed(os.path.join(HERE, "ops_cases.npz"), **ops)
    with open(os.path.join(HERE, "mx_manifest.json"), "w") as f:
        json.dump({"ops": ops_manifest, "generator": "tests/golden/gen_mx_golden.py",
                   "reference": "jeffreyyu0602/quantized-training @ /root/reference (CPU, torch %s)" % torch.__version__,
                   "cases": manifest, "scale_fn_quant_max": [448.0, 31.0, 6.0, 7.5, 32767.0]}, f, indent=1)
    for fn in ["mx_cases.npz", "mx_scale.npz", "ops_cases.npz", "mx_manifest.json"]:
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
