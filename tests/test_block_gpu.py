"""GPU parity tests of the block-scaled qschemes, microscaling and group_wise_affine (pytest -m gpu).

Bit-exact (NaN == NaN) against the reference's own outputs (tests/golden/mx_cases.npz, mx_scale.npz) and against
the CPU oracle on seeded inputs that steer every kernel (flat / cols / generic).  All calls go through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import nan_eq, nan_eq32

import quantized_training as qt
from quantized_training import _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bits_of(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).reshape(-1)
    return t.view(torch.int32).numpy().view(np.uint32).reshape(-1)


def tensor_from_bits(a, dtype, shape):
    if dtype == "bf16":
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).view(torch.float32)
    return t.reshape(shape).to(DEV)


def make_module(qscheme, element, ax, bs, qmin, qmax, scale_dtype=None, pow2=False):
    return qt.FusedAmaxObsFakeQuantize(element, qscheme=qscheme, quant_min=qmin, quant_max=qmax, ch_axis=ax,
                                       block_size=bs, scale_dtype=scale_dtype, force_scale_power_of_two=pow2,
                                       device=DEV)


def test_reference_goldens(golden):
    """Every reference-generated case: output, scale buffer (shape and bits), zero_point."""
    for case in golden.mx_manifest["cases"]:
        name = case["name"]
        x = tensor_from_bits(golden.mx[f"{name}/x"], case["dtype"], case["shape"])
        ax = case["ch_axis"]
        ax = tuple(ax) if isinstance(ax, list) else ax
        mod = make_module(case["qscheme"], case["element"], ax, case["block_size"], case["quant_min"],
                          case["quant_max"], case["scale_dtype"], case.get("force_scale_power_of_two", False))
        y = mod(x)
        assert y.shape == x.shape and y.dtype == x.dtype and y.is_contiguous()
        assert list(mod.scale.shape) == case["scale_shape"], name
        assert nan_eq32(bits_of(mod.scale), golden.mx[f"{name}/scale"]).all(), (name, "scale")
        if case["qscheme"] == "group_wise_affine":
            assert list(mod.zero_point.shape) == case["scale_shape"], name
            assert nan_eq32(bits_of(mod.zero_point), golden.mx[f"{name}/zero_point"]).all(), (name, "zero_point")
        bad = np.nonzero(~nan_eq(bits_of(y), golden.mx[f"{name}/y"].reshape(-1)))[0]
        assert bad.size == 0, (name, bad[:8])


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_scale_function_exhaustive(golden, dtype):
    """amax -> scale on the device for every positive bf16 amax / fp32 values around every power of two, as a
    one-element block (generic kernels) and as the maximum of an 8-wide block (flat kernel)."""
    if dtype == "bf16":
        amax_bits = np.arange(0x8000, dtype=np.uint16)
    else:
        amax_bits = golden.mx_scale["f32_amax_bits"]
    x1 = tensor_from_bits(amax_bits, dtype, (amax_bits.size, 1))
    x8 = torch.zeros(amax_bits.size, 8, dtype=x1.dtype, device=DEV)
    x8[:, 5] = -x1[:, 0]
    for qmax in golden.mx_manifest["scale_fn_quant_max"]:
        for mode, pow2, sdt in (("pow2", True, None), ("amax", False, None), ("e5m3", False, "fp8_e5m3")):
            want = golden.mx_scale[f"{dtype}/{mode}/{qmax}"]
            want = want.astype(np.uint32) << 16 if dtype == "bf16" else want
            for x, bs in ((x1, 1), (x8, 8)):
                mod = make_module("microscaling", "bfloat16", -1, bs, -qmax, qmax, sdt, pow2)
                mod(x)
                got = bits_of(mod.scale)
                bad = np.nonzero(~nan_eq32(got, want))[0]
                assert bad.size == 0, (dtype, mode, qmax, bs, [(hex(amax_bits[i]), hex(got[i]), hex(want[i]))
                                                              for i in bad[:6]])


def seeded_input(shape, dtype, seed, mag=1.0, affine=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g) * mag * torch.exp2(torch.randint(-8, 8, shape, generator=g).float())
    if affine:
        x = x + 0.4 * mag
    flat = x.view(-1)
    n = flat.numel()
    if n >= 4096:
        flat[100:300] = 0.0
        flat[1000] = float("nan")
        flat[2000] = float("inf")
        flat[3000] = -float("inf")
        flat[3500:3600] = 2.5
        flat[3700] = 1e-39
    return x.to(torch.bfloat16 if dtype == "bf16" else torch.float32)


MX_SWEEP = [
    # shape, ax, bs -- flat kernel: lanes 1..32 per block
    ((64, 512), -1, 8), ((64, 512), -1, 16), ((64, 512), -1, 32), ((64, 512), -1, 64), ((64, 512), -1, 128),
    ((64, 512), -1, 256), ((3, 5, 96), -1, 32), ((1000, 64), -1, 64),
    # cols kernel: RPT 1..16, ragged n, inner not a multiple of 32 vectors
    ((4, 64, 256), -2, 8), ((4, 64, 256), 1, 16), ((4, 96, 256), 1, 32), ((2, 200, 64), 1, 64),
    ((2, 300, 40), 1, 128), ((160, 1024), 0, 32), ((7, 70, 8), 1, 32),
    # tile kernel: two tiled axes (the last two), ragged rows / columns
    ((2, 3, 128, 256), (-2, -1), 64), ((3, 100, 72), (-2, -1), 32), ((1, 256, 512), (-2, -1), 128), ((4, 40, 64), (-2, -1), 8),
    # generic kernels: ragged last axis, odd block sizes, tiny inner, non-adjacent tiled axes
    ((2, 3, 48, 64), (-2, -1), 16), ((2, 40, 50), (-2, -1), 16), ((5, 70), -1, 32), ((6, 40, 3), 1, 8),
    ((9, 33), -1, 5), ((3, 50, 6), (0, 2), 4), ((31,), 0, 64),
]


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("element,sdt,pow2", [("int6", "fp8_e5m3", False), ("fp4_e2m1", None, True),
                                              ("fp8_e4m3", None, False), ("posit8_1", None, True),
                                              ("int8", "fp8_e4m3", False)])
def test_microscaling_vs_oracle(oracle, dtype, element, sdt, pow2):
    from quantized_training.quantizer import get_quant_min_max
    qmin, qmax = (float(v) for v in get_quant_min_max(element))
    tab = oracle.qmap(element)
    stab = oracle.qmap(sdt) if sdt else None
    for ci, (shape, ax, bs) in enumerate(MX_SWEEP):
        for mag in (1.0, 3e-4 if ci % 2 else 2e3):
            x = seeded_input(shape, dtype, 77 + ci, mag)
            xb = bits_of(x)
            xin = xb.view(np.float32) if dtype == "f32" else xb
            want_y, want_s = oracle.mx_fake_quant(xin, shape, ax, bs, qmax, tab, pow2, stab)
            mod = make_module("microscaling", element, ax, bs, qmin, qmax, sdt, pow2)
            y = mod(x.to(DEV))
            assert tuple(mod.scale.shape) == want_s.shape, (shape, ax, bs)
            assert nan_eq32(bits_of(mod.scale), want_s.reshape(-1).view(np.uint32)).all(), (shape, ax, bs, "scale")
            wy = want_y.view(np.uint32) if dtype == "f32" else want_y
            bad = np.nonzero(~nan_eq(bits_of(y), wy.reshape(-1)))[0]
            assert bad.size == 0, (element, dtype, shape, ax, bs, mag, bad[:8])


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("element,sdt", [("uint2", "fp8_e5m3"), ("uint4", None), ("int4", None), ("int8", "fp8_e5m3")])
def test_group_wise_affine_vs_oracle(oracle, dtype, element, sdt):
    from quantized_training.quantizer import get_quant_min_max
    qmin, qmax = (float(v) for v in get_quant_min_max(element))
    stab = oracle.qmap(sdt) if sdt else None
    for ci, (shape, ax, bs) in enumerate(MX_SWEEP):
        x = seeded_input(shape, dtype, 177 + ci, 1.0 if ci % 3 else 50.0, affine=True)
        xb = bits_of(x)
        xin = xb.view(np.float32) if dtype == "f32" else xb
        want_y, want_s, want_z = oracle.gwa_fake_quant(xin, shape, ax, bs, qmin, qmax, stab)
        mod = make_module("group_wise_affine", element, ax, bs, qmin, qmax, sdt)
        y = mod(x.to(DEV))
        assert tuple(mod.scale.shape) == want_s.shape == tuple(mod.zero_point.shape)
        assert nan_eq32(bits_of(mod.scale), want_s.reshape(-1).view(np.uint32)).all(), (shape, ax, bs, "scale")
        assert nan_eq32(bits_of(mod.zero_point), want_z.reshape(-1).view(np.uint32)).all(), (shape, ax, bs, "zp")
        wy = want_y.view(np.uint32) if dtype == "f32" else want_y
        bad = np.nonzero(~nan_eq(bits_of(y), wy.reshape(-1)))[0]
        assert bad.size == 0, (element, dtype, shape, ax, bs, bad[:8])


def test_large_tensor_properties():
    """2^27 elements (no oracle at this size): the result is a function of the block alone, so a row permutation
    commutes with it; the flat, cols and generic kernels agree on layouts all three can express; scales are
    positive and, with force_scale_power_of_two, powers of two."""
    g = torch.Generator(device=DEV).manual_seed(3)
    rows, cols = 1 << 15, 1 << 12
    x = torch.randn(rows, cols, generator=g, device=DEV, dtype=torch.bfloat16) * 3
    mod = make_module("microscaling", "fp4_e2m1", -1, 32, -6.0, 6.0, None, True)
    y = mod(x)
    s = mod.scale.clone()
    assert s.shape == (rows, cols // 32) and bool((s > 0).all())
    assert bool((torch.frexp(s)[0] == 0.5).all())
    perm = torch.randperm(rows, generator=g, device=DEV)
    assert torch.equal(mod(x[perm].contiguous()), y[perm])
    # the same blocks seen as an inner-axis tiling of the transposed tensor (cols kernel) ...
    xt = x[:4096].t().contiguous()  # [cols, 4096]: blocks along axis 0
    mt = make_module("microscaling", "fp4_e2m1", 0, 32, -6.0, 6.0, None, True)
    assert torch.equal(mt(xt).t(), y[:4096])
    assert torch.equal(mt.scale.t(), s[:4096])
    # ... and through the generic kernels (a misaligned view defeats the vector paths)
    buf = torch.empty(4096 * cols + 1, dtype=torch.bfloat16, device=DEV)
    xm = buf[1:].view(4096, cols)
    xm.copy_(x[:4096])
    assert torch.equal(mod(xm), y[:4096])
    # idempotent on its own output for power-of-two scales (block maxima survive quantization)
    assert torch.equal(mod(y), y)


def test_flags_gradients_and_state_dict():
    x = torch.randn(8, 128, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    mod = make_module("microscaling", "int6", -1, 64, -32.0, 31.0, "fp8_e5m3")
    y = mod(x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))  # straight-through
    mod.disable_fake_quant()
    assert torch.equal(mod(x), x)
    mod.enable_fake_quant()
    sd = mod.state_dict()
    assert sd["scale"].shape == (8, 2)
    fresh = make_module("microscaling", "int6", -1, 64, -32.0, 31.0, "fp8_e5m3")
    fresh.load_state_dict(sd)
    assert torch.equal(fresh.scale, mod.scale)
    # empty input, and quantize() wiring through the qspec grammar
    assert mod(torch.empty(0, 64, device=DEV, dtype=torch.bfloat16)).shape == (0, 64)
    lin = torch.nn.Sequential(torch.nn.Linear(128, 64)).to(DEV).bfloat16()
    args = qt.add_qspec_args().parse_args(["--activation", "int6,qs=microscaling,bs=64,ax=-1,scale=fp8_e5m3",
                                           "--weight", "int6,qs=microscaling,bs=64,ax=-1,scale=fp8_e5m3", "--bf16"])
    qt.quantize(lin, args)
    out = lin(x.detach())
    assert out.shape == (8, 64) and bool(torch.isfinite(out.float()).all())
    w = lin[0].weight_fake_quant
    assert w.scale.shape == (64, 2)
