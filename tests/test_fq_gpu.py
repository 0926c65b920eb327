"""GPU parity tests of the fake-quant hot path (run on the B200 box: pytest -m gpu).

Every comparison is bit-exact (NaN == NaN): integer/bit work, no tolerance.
Checkers: the reference's own outputs (tests/golden) and the CPU oracle on seeded inputs.
All calls go through the C ABI (quantized_training._C -> libqt_b200.so)."""
import zlib

import numpy as np
import pytest
import torch

from conftest import nan_eq, nan_eq16, nan_eq32

import quantized_training as qt
from quantized_training import _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

NORTH_STAR = ["int4", "int8", "e4m3", "e5m2", "fp8_e4m3", "fp8_e5m2", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1",
              "posit8_0", "posit8_1", "posit8_2", "posit16_1"]


def bits_of(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).reshape(-1)
    return t.view(torch.int32).numpy().view(np.uint32).reshape(-1)


def bf16_from_bits(a, device=DEV):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16).to(device)


def f32_from_bits(a, device=DEV):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).view(torch.float32).to(device)


def module_for(spec, **kw):
    qs = qt.QuantizationSpec.from_str(spec)
    return qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), device=DEV, **kw), qs


def oracle_for(oracle, qs, pow2=False):
    return oracle.FakeQuant(qs.dtype, qscheme=None if qs.qscheme is None else qs.qscheme.value,
                            quant_max=qs.quant_max, amax_history_len=qs.amax_history_len, ch_axis=qs.ch_axis,
                            force_scale_power_of_two=pow2)


def test_native_library_is_what_runs():
    assert torch.cuda.is_available()
    assert "sm_100a" in _C.version()
    maps = open("/proc/self/maps").read()
    assert "libqt_b200.so" in maps


def test_exhaustive_bf16_all_tables(golden):
    """All 65 536 bf16 inputs x every golden dtype, through the vector kernel, the scalar tail
    and a misaligned view."""
    all_bits = np.arange(65536, dtype=np.uint16)
    x = bf16_from_bits(all_bits)
    pad = torch.zeros(65536 + 8, dtype=torch.bfloat16, device=DEV)
    pad[3:3 + 65536] = x
    for d in golden.qmaps.files:
        mod = qt.FusedAmaxObsFakeQuantize(d, device=DEV)
        want = golden.qmaps[d]
        assert nan_eq16(bits_of(mod(x)), want).all(), d                       # 16-byte vectors
        assert nan_eq16(bits_of(mod(x[:65531])), want[:65531]).all(), d       # vectors + scalar tail
        assert nan_eq16(bits_of(mod(pad[3:3 + 65536])), want).all(), d        # misaligned base pointer


def test_exhaustive_fp32_round_to_odd(golden):
    xb = golden.vmap32["x_bits"]
    x = f32_from_bits(xb)
    for d in golden.vmap32.files:
        if d == "x_bits":
            continue
        mod = qt.FusedAmaxObsFakeQuantize(d, device=DEV)
        yb = bits_of(mod(x))
        not_nan = (yb & 0x7FFFFFFF) <= 0x7F800000
        assert np.all((yb[not_nan] & 0xFFFF) == 0), d  # outputs are bf16-exact (NaN payloads are free)
        assert nan_eq16((yb >> 16).astype(np.uint16), golden.vmap32[d]).all(), d
        yb = bits_of(mod(x[1:-2]))  # misaligned + tail
        assert nan_eq16((yb >> 16).astype(np.uint16), golden.vmap32[d][1:-2]).all(), d


def test_reference_call_sequences(golden):
    """Delayed scaling against the reference's own outputs: y, scale and amax_history after every call."""
    for case in golden.manifest["fq_cases"]:
        name = case["name"]
        mod, qs = module_for(case["spec"], force_scale_power_of_two=case["force_scale_power_of_two"])
        f32 = case["dtype"] == "fp32"
        mk = f32_from_bits if f32 else bf16_from_bits
        for k in range(case["calls"]):
            x = mk(golden.fq[f"{name}/x{k}"]).reshape(case["shape"])
            y = mod(x)
            assert y.shape == x.shape and y.dtype == x.dtype and y.is_contiguous()
            assert nan_eq(bits_of(y), golden.fq[f"{name}/y{k}"].reshape(-1)).all(), (name, k)
            assert list(mod.scale.shape) == case["scale_shape"][k], (name, k)
            assert list(mod.amax_history.shape) == case["hist_shape"][k], (name, k)
            assert nan_eq32(bits_of(mod.scale), golden.fq[f"{name}/scale{k}"]).all(), (name, k, "scale")
            if case["qscheme"] is not None:
                assert nan_eq32(bits_of(mod.amax_history), golden.fq[f"{name}/hist{k}"]).all(), (name, k, "hist")
        mod.disable_observer()
        x = mk(golden.fq[f"{name}/x_obsoff"]).reshape(case["shape"])
        assert nan_eq(bits_of(mod(x)), golden.fq[f"{name}/y_obsoff"].reshape(-1)).all(), (name, "observer off")
        assert nan_eq32(bits_of(mod.scale), golden.fq[f"{name}/scale_obsoff"]).all()


SPECS = ["{d}", "{d},qs=per_tensor_symmetric,ahl=3", "{d},qs=per_channel_symmetric,ax=0,ahl=2",
         "{d},qs=per_channel_symmetric,ax=-1,ahl=2", "{d},qs=per_channel_symmetric,ax=1,ahl=2"]


@pytest.mark.parametrize("dtype", ["int8", "int4", "fp8_e4m3", "fp8_e5m2", "fp6_e3m2", "fp4_e2m1", "posit8_1", "posit8_2"])
@pytest.mark.parametrize("elem", ["bf16", "fp32"])
def test_against_oracle_seeded(oracle, dtype, elem):
    """Seeded tensors in shapes that hit every kernel (flat / rows / cols / scalar) and ragged edges."""
    shapes = [(4, 256, 512), (3, 40, 264), (2, 33, 77), (1, 1, 8), (7,)]
    g = torch.Generator().manual_seed(zlib.crc32(f"{dtype}/{elem}".encode()))
    for spec_t in SPECS:
        for shape in shapes:
            if "per_channel" in spec_t and len(shape) < 3:
                continue
            mod, qs = module_for(spec_t.format(d=dtype))
            ref = oracle_for(oracle, qs)
            for call, mag in enumerate([1.0, 300.0, 1e-3]):
                x = torch.randn(shape, generator=g) * mag
                x = x.to(torch.bfloat16) if elem == "bf16" else x
                y = mod(x.to(DEV))
                xin = bits_of(x) if elem == "bf16" else x.numpy().reshape(-1)
                want = ref(xin, shape)
                want = want.view(np.uint32) if elem == "fp32" else want
                assert nan_eq(bits_of(y), want.reshape(-1)).all(), (spec_t, shape, call)
                assert nan_eq32(bits_of(mod.scale), ref.scale.view(np.uint32)).all(), (spec_t, shape, call)
                if qs.qscheme is not None:
                    assert nan_eq32(bits_of(mod.amax_history), ref.history.view(np.uint32)).all()


def test_edge_cases(oracle):
    mod = qt.FusedAmaxObsFakeQuantize("posit8_1", device=DEV)
    e = torch.empty(0, 5, dtype=torch.bfloat16, device=DEV)
    assert mod(e).shape == (0, 5)                                     # empty input, observer off
    obs, _ = module_for("int8,qs=per_tensor_symmetric")
    with pytest.raises(RuntimeError):
        obs(e)                                                        # amax of an empty tensor raises, as torch.amax does
    x = torch.randn(64, 48, device=DEV).to(torch.bfloat16)
    xt = x.t()                                                        # non-contiguous in -> contiguous out
    y = mod(xt)
    assert y.is_contiguous() and y.shape == xt.shape
    want = oracle.vmap(bits_of(xt.contiguous()), oracle.qmap("posit8_1"))
    assert nan_eq16(bits_of(y), want).all()
    with pytest.raises(TypeError):
        mod(x.to(torch.float64))                                      # float16 has its own path (test below)
    # observer on, fake quant off: statistics move, tensor passes through
    obs.disable_fake_quant()
    x32 = torch.randn(1000, device=DEV) * 7
    out = obs(x32)
    assert out.data_ptr() == x32.data_ptr()
    assert float(obs.amax_history[0]) == float(x32.abs().max())
    # straight-through gradient
    m2 = qt.FusedAmaxObsFakeQuantize("e4m3", device=DEV)
    xg = torch.randn(100, device=DEV, requires_grad=True)
    m2(xg).sum().backward()
    assert torch.equal(xg.grad, torch.ones_like(xg))


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "int8,qs=per_tensor_symmetric,ahl=1"])
def test_full_size_properties(spec):
    """2^28 elements (0.5 GiB bf16): size-independent properties instead of an oracle run.
    (i) idempotence fq(fq(x)) == fq(x) at fixed scale, (ii) odd symmetry away from zero,
    (iii) block k of the output equals the small-case result on block k (position independence),
    (iv) the observed amax equals torch's."""
    n = 1 << 28
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(n, device=DEV, generator=g, dtype=torch.float32).mul_(4.0).to(torch.bfloat16)
    mod, qs = module_for(spec)
    y = mod(x)
    if qs.qscheme is not None:
        assert float(mod.amax_history[0]) == float(x.abs().max())
        y = mod(x)          # second call: scale = amax/qmax of the first
        mod.disable_observer()
    y2 = mod(y)
    assert torch.equal(y2.view(torch.int16), y.view(torch.int16))
    if not spec.startswith("int"):  # intN clamps at -2^(N-1) / 2^(N-1)-1: not odd at saturation, like the reference
        yn = mod(-x)
        nz = y != 0
        assert torch.equal((-yn)[nz].view(torch.int16), y[nz].view(torch.int16))
    for start in (0, 12345 * 8, n - 4096):
        blk = mod(x[start:start + 4096].clone())
        assert torch.equal(blk.view(torch.int16), y[start:start + 4096].view(torch.int16))


def test_host_streaming_matches_device_call():
    """Chunked host pipeline == one device call: values, scale and amax history (delayed scaling across chunks)."""
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(3_000_001, generator=g) * 5).to(torch.bfloat16).pin_memory()
    for spec in ("posit8_1", "fp8_e4m3,qs=per_tensor_symmetric,ahl=3"):
        a, _ = module_for(spec)
        b, _ = module_for(spec)
        pipe = qt.HostPipeline(DEV, torch.bfloat16, chunk_elems=1 << 18, depth=3)
        for call in range(3):
            xi = (x.float() * (call + 1)).to(torch.bfloat16).pin_memory()
            want = a(xi.to(DEV))
            got = pipe.run(b, xi)
            torch.cuda.synchronize()
            assert nan_eq16(bits_of(got), bits_of(want)).all(), (spec, call)
            assert nan_eq32(bits_of(a.scale), bits_of(b.scale)).all()
            if a.qscheme is not None:
                assert nan_eq32(bits_of(a.amax_history), bits_of(b.amax_history)).all()
    # several modules over one host tensor: one upload per chunk, one result per module
    specs = ("int8", "e4m3", "posit8_2", "fp8_e4m3,qs=per_tensor_symmetric,ahl=3")
    mods = [module_for(sp)[0] for sp in specs]
    refs = [module_for(sp)[0] for sp in specs]
    pipe = qt.HostPipeline(DEV, torch.bfloat16, chunk_elems=1 << 18, depth=3)
    for call in range(2):
        xi = (x.float() * (call + 2)).to(torch.bfloat16).pin_memory()
        outs = pipe.run_many(mods, xi, [torch.empty_like(xi).pin_memory() for _ in specs])
        torch.cuda.synchronize()
        for sp, got, r, m in zip(specs, outs, refs, mods):
            assert nan_eq16(bits_of(got), bits_of(r(xi.to(DEV)))).all(), (sp, call)
            assert nan_eq32(bits_of(r.scale), bits_of(m.scale)).all()


@pytest.mark.parametrize("spec", ["e4m3", "e5m2", "fp8_e4m3", "fp8_e5m2", "fp8_e4m3,qs=per_tensor_symmetric,ahl=2"])
@pytest.mark.parametrize("elem", ["bf16", "fp32"])
def test_codes_decode_to_the_fake_quant_values(spec, elem):
    """decode(quantize_to_codes(x)) * scale == forward(x), bit for bit, incl. every bf16 pattern."""
    all_bits = np.arange(65536, dtype=np.uint16)
    g = torch.Generator().manual_seed(17)
    xs = [bf16_from_bits(all_bits), (torch.randn(100003, generator=g) * 30).to(torch.bfloat16).to(DEV)]
    if elem == "fp32":
        xs = [x.float() * 1.0001 for x in xs]
    for x in xs:
        a, qs = module_for(spec)
        b, _ = module_for(spec)
        for call in range(2):
            y = a(x)
            codes = b.quantize_to_codes(x)
            assert codes.dtype == torch.uint8 and codes.shape == x.shape
            f8 = torch.float8_e4m3fn if b.fp8_kind == "e4m3" else torch.float8_e5m2
            q = codes.view(f8).to(torch.float32)
            s = b.scale.to(x.dtype).float()
            dec = (q * s).to(x.dtype) if elem == "bf16" else q * b.scale
            got, want = bits_of(dec), bits_of(y)
            finite_codes = ~torch.isinf(y.float().cpu()).numpy().reshape(-1) if b.fp8_kind == "e4m3" else np.ones_like(got, bool)
            # e4m3 has no Inf: fpN_eXmY's Inf pass-through becomes the NaN code (documented), everything else is exact
            assert nan_eq(got[finite_codes], want[finite_codes]).all(), (spec, elem, call)
            assert nan_eq32(bits_of(a.scale), bits_of(b.scale)).all()


def test_more_than_four_giga_elements():
    """BASELINE's sweep goes up to 4 G elements: one call on 2^32 + 2^20 bf16 elements (8 GiB in, 8 GiB out; element
    and byte offsets beyond 32 bits).  The input is a 2^20-element block repeated, so every block of the output must
    equal the output of that block alone (position independence), the amax must be the block's, and the block-scaled
    kernel must write the same scales for every repetition."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~40 GiB of free device memory")
    g = torch.Generator(device=DEV).manual_seed(11)
    blk = (torch.randn(1 << 20, generator=g, device=DEV) * 7).to(torch.bfloat16)
    reps = (1 << 12) + 1
    x = blk.repeat(reps)
    assert x.numel() == (1 << 32) + (1 << 20)
    for spec in ("posit8_1", "fp8_e4m3,qs=per_tensor_symmetric,ahl=2", "int8"):
        mod, _ = module_for(spec)
        small, _ = module_for(spec)
        for _ in range(2):                      # second call: the delayed scale is live
            y = mod(x)
            want = small(blk)
            assert torch.equal(y.view(reps, -1).view(torch.int16), want.view(torch.int16).expand(reps, -1)), spec
            del y
        if mod.qscheme is not None:
            assert torch.equal(mod.amax_history, small.amax_history) and torch.equal(mod.scale, small.scale)
    mx = qt.FusedAmaxObsFakeQuantize("fp4_e2m1", qscheme="microscaling", quant_min=-6.0, quant_max=6.0, ch_axis=-1,
                                     block_size=32, force_scale_power_of_two=True, device=DEV)
    y = mx(x)
    s_all = mx.scale.clone()
    want = mx(blk)
    assert torch.equal(y.view(reps, -1).view(torch.int16), want.view(torch.int16).expand(reps, -1))
    assert torch.equal(s_all.view(reps, -1), mx.scale.view(1, -1).expand(reps, -1))


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "int8,qs=per_tensor_symmetric,ahl=4",
                                  "fp8_e4m3,qs=per_channel_symmetric,ax=0,ahl=2", "fp6_e3m2,qs=per_tensor_symmetric,ahl=1"])
def test_float16_tensors_follow_the_reference_arithmetic(oracle, spec):
    """float16 inputs (compatibility path): the reference's dtype-generic chain restated with torch ops on the device --
    amax of the fp16 tensor into the fp32 history, delayed scale, `scale.to(fp16)`, fp16 division, lookup of the
    round-to-odd-truncated fp32 widening in the oracle's table, narrowing to fp16, fp16 multiply
    (fake_quantize.py:217-246, decomposed.py:147-163) -- bit for bit over a sequence of calls."""
    mod, qs = module_for(spec)
    table = torch.from_numpy(oracle.qmap(qs.dtype).view(np.int16)).view(torch.bfloat16).to(DEV)
    g = torch.Generator().manual_seed(17)
    hist = scale = None
    for call in range(4):
        x = (torch.randn(48, 520, generator=g) * (4.0 ** call)).to(torch.float16).to(DEV)
        if call == 2:
            x[0, :4] = torch.tensor([float("inf"), float("nan"), 6e-8, -0.0], dtype=torch.float16)
        want = x
        if qs.qscheme is not None:
            cur = x.abs().amax(dim=1, keepdim=True) if qs.ch_axis is not None else x.abs().amax()
            if hist is None:
                hist = torch.zeros((qs.amax_history_len,) + tuple(cur.shape), device=DEV)
                scale = torch.ones(tuple(cur.shape), device=DEV)
            amax = hist.amax(dim=0)
            if hist.shape[0] > 1:
                hist = torch.roll(hist, -1, 0)
            hist[0] = cur
            sf = amax / torch.full_like(amax, qs.quant_max)   # a tensor divisor: true division (a python scalar is
            scale = torch.where((amax > 0) & torch.isfinite(amax), sf, scale)   # multiplied by its reciprocal on CUDA)
        s16 = (scale if scale is not None else torch.ones((), device=DEV)).to(torch.float16)
        bits = (x / s16).float().view(torch.int32)
        idx = ((bits >> 16) & 0xFFFF) | ((bits & 0xFFFF) != 0).to(torch.int32)
        want = table[idx].to(torch.float16) * s16
        got = mod(x)
        assert got.dtype == torch.float16
        a, b = got.view(torch.int16).cpu().numpy().view(np.uint16), want.view(torch.int16).cpu().numpy().view(np.uint16)
        same = (a == b) | (((a & 0x7FFF) > 0x7C00) & ((b & 0x7FFF) > 0x7C00))
        assert same.all(), f"call {call}: {int((~same).sum())} mismatches"
        if hist is not None:
            assert nan_eq32(bits_of(mod.amax_history.reshape(hist.shape)), bits_of(hist)).all()
            assert nan_eq32(bits_of(mod.scale.reshape(scale.shape)), bits_of(scale)).all()
