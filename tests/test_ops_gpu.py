"""GPU parity of the operator surface torch.ops.quantized_ops.{vmap, quantize, dequantize} (pytest -m gpu):
bit-exact against outputs of the reference's own ops (decomposed.py:143-262) run on CPU (tests/golden/ops_cases.npz,
made by tests/golden/gen_mx_golden.py), including codebooks the bitwise rounders do not cover (NF4, a random table)."""
import numpy as np
import pytest
import torch

from conftest import nan_eq

import quantized_training as qt  # noqa: F401  (registers the ops)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_tensor(bits, dtype, shape):
    if dtype == "bf16":
        t = torch.from_numpy(np.ascontiguousarray(bits).view(np.int16)).view(torch.bfloat16)
    else:
        t = torch.from_numpy(np.ascontiguousarray(bits).view(np.int32)).view(torch.float32)
    return t.reshape(shape).to(DEV)


def bits_of(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).reshape(-1)
    return t.view(torch.int32).numpy().view(np.uint32).reshape(-1)


def test_reference_op_goldens(golden):
    cases = golden.mx_manifest["ops"]
    assert len(cases) >= 9
    tables = {k[len("table/"):]: to_tensor(golden.ops[k], "bf16", (65536,)) for k in golden.ops.files
              if k.startswith("table/")}
    for c in cases:
        n = c["name"]
        x = to_tensor(golden.ops[f"{n}/x"], c["x_dtype"], c["x_shape"])
        s = to_tensor(golden.ops[f"{n}/scale"], c["scale_dtype"], c["scale_shape"])
        zp = to_tensor(golden.ops[f"{n}/zp"], c["zp_dtype"], c["scale_shape"]) if c["zp"] else None
        ta = tables[c["table_a"]] if c["table_a"] else None
        tb = tables[c["table_b"]] if c["table_b"] else None
        if c["op"] == "vmap":
            y = torch.ops.quantized_ops.vmap(x, ta)
        elif c["op"] == "quantize":
            y = torch.ops.quantized_ops.quantize(x, s, zp, c["axes"], c["block_size"], ta)
        else:
            y = torch.ops.quantized_ops.dequantize(x, s, zp, c["axes"], c["block_size"], ta, tb)
        assert y.shape == x.shape and y.is_contiguous()
        assert ("bf16" if y.dtype == torch.bfloat16 else "f32") == c["y_dtype"], (n, y.dtype)
        bad = np.nonzero(~nan_eq(bits_of(y), golden.ops[f"{n}/y"].reshape(-1)))[0]
        assert bad.size == 0, (n, bad[:8])


def test_vmap_equals_the_module_tables(golden):
    """vmap with the reference's table == the module's bitwise rounding, on every bf16 pattern."""
    allb = to_tensor(np.arange(65536, dtype=np.uint16), "bf16", (65536,))
    for d in ("posit8_1", "fp6_e3m2", "int4", "e4m3"):
        table = to_tensor(golden.qmaps[d], "bf16", (65536,))
        a = torch.ops.quantized_ops.vmap(allb, table)
        b = qt.FusedAmaxObsFakeQuantize(d, device=DEV)(allb)
        assert nan_eq(bits_of(a), bits_of(b)).all(), d


def test_cpu_tensors_are_rejected():
    x = torch.zeros(8, dtype=torch.bfloat16)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.quantized_ops.vmap(x, torch.zeros(65536, dtype=torch.bfloat16))


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_calculate_mx_qparam_op(golden, dtype):
    """torch.ops.quantized_ops.calculate_mx_qparam against the reference's own function on every positive bf16 amax /
    fp32 values around every power of two (tests/golden/mx_scale.npz), the codebook passed as a TABLE like the
    reference does."""
    amax_bits = np.arange(0x8000, dtype=np.uint16) if dtype == "bf16" else golden.mx_scale["f32_amax_bits"]
    x = to_tensor(amax_bits, dtype, (amax_bits.size, 1))
    e5m3 = to_tensor(golden.qmaps["fp8_e5m3"], "bf16", (65536,))
    for qmax in golden.mx_manifest["scale_fn_quant_max"]:
        for mode, pow2, tab in (("pow2", True, None), ("amax", False, None), ("e5m3", False, e5m3)):
            s = torch.ops.quantized_ops.calculate_mx_qparam(x, [-1], 1, qmax, pow2, tab)
            assert s.dtype == x.dtype and s.shape == x.shape
            bad = np.nonzero(~nan_eq(bits_of(s), golden.mx_scale[f"{dtype}/{mode}/{qmax}"]))[0]
            assert bad.size == 0, (dtype, mode, qmax, bad[:6])


@pytest.mark.parametrize("element,sdt,pow2,shape,axes,bs", [
    ("int6", "fp8_e5m3", False, (16, 256), [-1], 32), ("fp4_e2m1", None, True, (4, 96, 64), [-2], 32),
    ("int6", "fp8_e5m3", False, (2, 48, 64), [-2, -1], 16), ("posit8_1", None, False, (5, 70), [-1], 32)])
def test_quantize_mx_op_agrees_with_the_module(golden, element, sdt, pow2, shape, axes, bs):
    """(scale, q) = quantize_mx(x, qmap, ...): scale is the module's `scale` buffer and q * expand(scale) its output
    (MXFakeQuantFunction.forward = quantize_mx + one multiply, fake_quantize.py:118-129)."""
    from quantized_training.decomposed import expand
    from quantized_training.quantizer import get_quant_min_max
    torch.manual_seed(5)
    x = (torch.randn(shape, device=DEV) * 3 * torch.exp2(torch.randint(-5, 5, shape, device=DEV).float())).bfloat16()
    qmin, qmax = (float(v) for v in get_quant_min_max(element))
    qmap = qt.get_quantization_map(element, DEV)
    smap = qt.get_quantization_map(sdt, DEV) if sdt else None
    scale, q = torch.ops.quantized_ops.quantize_mx(x, qmap, axes, bs, qmax, pow2, smap)
    mod = qt.FusedAmaxObsFakeQuantize(element, qscheme="microscaling", quant_min=qmin, quant_max=qmax,
                                      ch_axis=tuple(axes) if len(axes) > 1 else axes[0], block_size=bs, scale_dtype=sdt,
                                      force_scale_power_of_two=pow2, device=DEV)
    y = mod(x)
    assert scale.dtype == x.dtype and tuple(scale.shape) == tuple(mod.scale.shape)
    assert torch.equal(scale.float(), mod.scale)
    assert nan_eq(bits_of(q * expand(scale, x.shape, bs)), bits_of(y)).all()


def test_linear_and_matmul_mx_ops():
    """linear_mx / matmul_mx = dequantize the operands (codebook decode, block scales) and multiply on the tcgen05 GEMM:
    against the reference's formula evaluated with torch on the same device (decomposed.py:311-363), relative error
    <= 2^-7 (bf16 output, fp32 accumulation; the accumulation order differs from cuBLAS)."""
    from quantized_training.decomposed import expand
    torch.manual_seed(9)
    M, K, N, bs = 64, 256, 128, 32
    xq = torch.randint(-31, 32, (M, K), device=DEV).bfloat16()
    wq = torch.randint(-31, 32, (N, K), device=DEV).bfloat16()
    xs = (torch.rand(M, K // bs, device=DEV) * 0.1 + 0.01).bfloat16()
    ws = (torch.rand(N, K // bs, device=DEV) * 0.1 + 0.01).bfloat16()
    bias = torch.randn(N, device=DEV).bfloat16()
    got = torch.ops.quantized_ops.linear_mx(xq, wq, bias, input_scale=xs, weight_scale=ws, block_size=bs)
    want = torch.nn.functional.linear((xq * expand(xs, xq.shape, bs)).float(), (wq * expand(ws, wq.shape, bs)).float(),
                                      bias.float())
    assert got.dtype == torch.bfloat16 and got.shape == (M, N)
    assert float((got.float() - want).norm() / want.norm()) <= 2 ** -7
    # codebook operands: indices into a 16-entry code
    code = torch.linspace(-1, 1, 16, device=DEV).bfloat16()
    wi = torch.randint(0, 16, (N, K), device=DEV).bfloat16()
    got = torch.ops.quantized_ops.linear_mx(xq, wi, None, input_scale=xs, weight_scale=ws, block_size=bs, weight_code=code)
    want = torch.nn.functional.linear((xq * expand(xs, xq.shape, bs)).float(),
                                      (code[wi.long()] * expand(ws, wq.shape, bs)).float())
    assert float((got.float() - want).norm() / want.norm()) <= 2 ** -7
    a = torch.randint(-7, 8, (2, 4, 64, 128), device=DEV).bfloat16()
    b = torch.randint(-7, 8, (2, 4, 128, 64), device=DEV).bfloat16()
    sa = (torch.rand(2, 4, 64, 4, device=DEV) + 0.5).bfloat16()
    sb = (torch.rand(2, 4, 4, 64, device=DEV) + 0.5).bfloat16()
    got = torch.ops.quantized_ops.matmul_mx(a, b, input_scale=sa, weight_scale=sb, block_size=32)
    want = torch.matmul((a * expand(sa, a.shape, 32)).float(), (b * expand(sb, b.shape, 32)).float())
    assert float((got.float() - want).norm() / want.norm()) <= 2 ** -7


@pytest.mark.parametrize("xe,we", [("fp8_e4m3", "fp8_e4m3"), ("fp4_e2m1", "fp4_e2m1"), ("fp8_e5m2", "fp6_e3m2"),
                                   ("fp6_e2m3", "fp8_e5m2")])
@pytest.mark.parametrize("shape,N", [((256, 512), 384), ((4, 100, 4096), 1024), ((130, 160), 200)])
def test_linear_mx_on_the_block_scaled_tensor_cores(xe, we, shape, N, monkeypatch):
    """Microscaling operands made by the reference's own recipe (quantize_mx: block 32 along K, power-of-two scales)
    go to tcgen05.mma kind::mxf8f6f4.block_scale as one-byte codes + UE8M0 scale bytes -- nothing is dequantized in
    HBM.  Checked against the reference formula (decomposed.py:311-331: dequantize, F.linear) in fp64 at the GEMM
    tolerance of test_gemm_gpu.py, and against the dequantizing route of this library; operands that do not qualify
    (scales that are not powers of two, integer elements outside every fp8 grid) silently take the old route."""
    from quantized_training import _C, decomposed
    from quantized_training.decomposed import expand
    from quantized_training.quantizer import get_quant_min_max
    torch.manual_seed(21)
    K = shape[-1]

    def mx(t, element, pow2=True):
        qmax = float(get_quant_min_max(element)[1])
        return torch.ops.quantized_ops.quantize_mx(t, qt.get_quantization_map(element, DEV), [-1], 32, qmax, pow2, None)

    x = (torch.randn(shape, device=DEV) * torch.exp2(torch.randint(-4, 5, (*shape[:-1], 1), device=DEV).float())).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV).bfloat16()
    (xs, xq), (ws, wq) = mx(x, xe), mx(w, we)
    taken = []
    real = _C.gemm_nt
    monkeypatch.setattr(_C, "gemm_nt", lambda *a, **k: (taken.append(k.get("sf_a") is not None), real(*a, **k))[1])
    got = torch.ops.quantized_ops.linear_mx(xq, wq, bias, input_scale=xs, weight_scale=ws, block_size=32)
    assert taken == [True] and got.dtype == torch.bfloat16 and got.shape == (*shape[:-1], N)
    ref = (xq.double() * expand(xs, xq.shape, 32).double()).reshape(-1, K) @ (wq.double() * expand(ws, wq.shape, 32).double()).t()
    ref = (ref + bias.double()).reshape(got.shape)
    err = (got.double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -8 * ref.pow(2).mean().sqrt()
    assert not (err > tol).any(), f"{int((err > tol).sum())} of {err.numel()} outside tolerance"
    monkeypatch.setattr(decomposed, "MX_TENSOR_CORES", "0")
    old = torch.ops.quantized_ops.linear_mx(xq, wq, bias, input_scale=xs, weight_scale=ws, block_size=32)
    assert taken == [True, False]
    assert float((got.float() - old.float()).norm() / old.float().norm()) < 2 ** -7
    # not microscaling data: amax / qmax scales (not powers of two) -> the dequantizing route, same numbers as before
    monkeypatch.setattr(decomposed, "MX_TENSOR_CORES", "1")
    (xs2, xq2) = mx(x, xe, pow2=False)
    got2 = torch.ops.quantized_ops.linear_mx(xq2, wq, bias, input_scale=xs2, weight_scale=ws, block_size=32)
    assert taken[-1] is False
    ref2 = (xq2.double() * expand(xs2, xq2.shape, 32).double()).reshape(-1, K) @ (wq.double() * expand(ws, wq.shape, 32).double()).t()
    assert float((got2.double().reshape(-1, N) - ref2 - bias.double()).norm() / ref2.norm()) < 2 ** -7


def test_conv2d_mx_op():
    """conv2d_mx = dequantize both operands (block scales along the channel axis, expand semantics) + F.conv2d
    (decomposed.py:273-300): identical to the reference formula evaluated with torch ops on the same device."""
    from quantized_training.decomposed import expand
    torch.manual_seed(3)
    x = torch.randint(-7, 8, (2, 64, 12, 12), device=DEV).bfloat16()
    w = torch.randint(-7, 8, (16, 64, 3, 3), device=DEV).bfloat16()
    xs = torch.exp2(torch.randint(-3, 3, (2, 2, 12, 12), device=DEV).float()).bfloat16()     # 32 channels per block
    ws = torch.exp2(torch.randint(-3, 3, (16, 2, 3, 3), device=DEV).float()).bfloat16()
    bias = torch.randn(16, device=DEV).bfloat16()
    got = torch.ops.quantized_ops.conv2d_mx(x, w, bias, [1, 1], [1, 1], [1, 1], 1, input_scale=xs, weight_scale=ws,
                                            block_size=32)
    want = torch.nn.functional.conv2d(x * expand(xs, x.shape, 32), w * expand(ws, w.shape, 32), bias, 1, 1, 1, 1)
    assert got.shape == (2, 16, 12, 12) and torch.equal(got, want)


@pytest.mark.parametrize("shape_a,shape_b", [((2, 4, 256, 128), (2, 4, 128, 256)), ((3, 130, 160), (3, 160, 208)),
                                             ((192, 256), (256, 384)), ((1, 8, 384, 64), (1, 8, 64, 384))])
@pytest.mark.parametrize("ea,eb", [("fp8_e4m3", "fp8_e4m3"), ("fp4_e2m1", "fp8_e5m2")])
def test_matmul_mx_on_the_block_scaled_tensor_cores(shape_a, shape_b, ea, eb, monkeypatch):
    """matmul_mx (decomposed.py:341-363) with microscaling operands -- `self` scaled along its last axis, `other` along
    its second to last -- as ONE batched block-scaled product: other is read as it is stored (MN-major fp8 tiles), its
    scales are packed from the [K / 32, N] matrices.  Against the reference formula in fp64 at the GEMM tolerance."""
    from quantized_training import _C
    from quantized_training.decomposed import expand
    from quantized_training.quantizer import get_quant_min_max
    torch.manual_seed(33)

    def mx(t, element, axis):
        qmax = float(get_quant_min_max(element)[1])
        return torch.ops.quantized_ops.quantize_mx(t, qt.get_quantization_map(element, DEV), [axis], 32, qmax, True, None)

    a = (torch.randn(shape_a, device=DEV) * 2).bfloat16()
    b = (torch.randn(shape_b, device=DEV) * 0.1).bfloat16()
    (sa, aq), (sb, bq) = mx(a, ea, -1), mx(b, eb, -2)
    taken = []
    real = _C.gemm_nt
    monkeypatch.setattr(_C, "gemm_nt", lambda *x, **k: (taken.append(k.get("sf_a") is not None), real(*x, **k))[1])
    got = torch.ops.quantized_ops.matmul_mx(aq, bq, input_scale=sa, weight_scale=sb, block_size=32)
    assert taken == [True] and got.dtype == torch.bfloat16 and got.shape == (*shape_a[:-1], shape_b[-1])
    ref = torch.matmul(aq.double() * expand(sa, aq.shape, 32).double(), bq.double() * expand(sb, bq.shape, 32).double())
    err = (got.double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -8 * ref.pow(2).mean().sqrt()
    assert not (err > tol).any(), f"{int((err > tol).sum())} of {err.numel()} outside tolerance"
