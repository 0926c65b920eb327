"""CPU-only checks of the product's host side: the C-ABI library loads and exports what
include/qt_b200.h declares, dtype parsing / limits follow the reference, the device rounding
logic (evaluated on the host by qt_table_host) reproduces the reference's tables, and the
product never touches the oracle or falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, nan_eq16

import quantized_training as qt
from quantized_training import _C


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qt_b200.h")).read()
    declared = set(re.findall(r"\b(qt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_C.EXPORTS), declared ^ set(_C.EXPORTS)
    L = ctypes.CDLL(_C.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert "sm_100a" in _C.version()


def test_rounding_logic_matches_reference_tables(golden):
    """Same code the kernels run (csrc/qt_round.h), evaluated on the host over all 2^16 bf16 inputs."""
    for d in golden.qmaps.files:
        t = qt.get_quantization_map(d).view(torch.int16).numpy().view(np.uint16)
        bad = np.nonzero(~nan_eq16(t, golden.qmaps[d]))[0]
        assert bad.size == 0, (d, [(hex(i), hex(t[i]), hex(golden.qmaps[d][i])) for i in bad[:8]])


def test_rounding_logic_matches_oracle_on_parameter_sweep(oracle):
    names = [f"posit{n}_{es}" for n in range(3, 25) for es in range(5) if (n - 2) * 2 ** es <= 126]
    names += [f"{p}{n}" for n in range(1, 25) for p in ("int", "uint")]
    names += [f"fp{e + m + 1 - u}_e{e}m{m}" for e in range(2, 6) for m in range(1, 6) for u in (0, 1) if e + m <= 8]
    for d in names:
        t = qt.get_quantization_map(d).view(torch.int16).numpy().view(np.uint16)
        assert nan_eq16(t, oracle.qmap(d)).all(), d


@pytest.mark.parametrize("bad", ["nf4", "float8", "int", "posit8", "fp8_e4m2", "e3m4", "", "posit8_1 ",
                                 "fp9_e5m4", "fp3_e1m1", "FP8_E4M3", "Posit8_1"])
def test_unsupported_dtype(bad):
    with pytest.raises(ValueError, match="Unsupported dtype"):
        qt.FusedAmaxObsFakeQuantize(bad)


def test_quant_min_max():
    from quantized_training.quantizer import get_quant_min_max as mm
    assert mm("int8") == (-128, 127) and mm("uint4") == (0, 15) and mm("INT4") == (-8, 7)
    assert mm("fp8_e4m3") == (-448.0, 448.0) and mm("fp8_e5m2") == (-57344.0, 57344.0)
    assert mm("fp6_e3m2") == (-28.0, 28.0) and mm("fp6_e2m3") == (-7.5, 7.5) and mm("fp4_e2m1") == (-6.0, 6.0)
    assert mm("posit8_1") == (-4096, 4096) and mm("posit8_2") == (-2 ** 24, 2 ** 24) and mm("posit8_0") == (-64, 64)
    assert mm("nf4") == (-1, 1) and mm("nf4_6") == (-31, 31)
    for bad in ("e4m3", "float32", "fp8"):
        with pytest.raises(ValueError):
            mm(bad)


def test_qspec_grammar():
    S = qt.QuantizationSpec.from_str
    s = S("fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10")
    assert (s.dtype, s.qscheme, s.quant_max, s.quant_min, s.amax_history_len) == \
        ("fp8_e5m2", qt.per_tensor_symmetric, 57344.0, -57344.0, 10)
    s = S("int8,qscheme=per_channel_symmetric,ch_axis=0")
    assert (s.quant_max, s.amax_history_len, s.ch_axis) == (127.0, 16, 0)
    s = S("posit8_1,qs=per_tensor_symmetric,qmax=64")
    assert s.quant_max == 64.0 and s.quant_min == -4096.0
    assert S("int4,qs=microscaling,bs=(1,32),ax=(0,1)").block_size == (1, 32)
    assert S("e4m3").qscheme is None and S("e4m3").quant_max is None
    assert S(s) is s  # --error arrives already parsed from argparse
    for bad, msg in [("", "None or empty"), ("int8,foo", "key=value"), ("int8,foo=1", "Unknown argument"),
                     ("e4m3,qs=per_tensor_symmetric,qmax=448", "Unsupported dtype"),
                     ("int8,qs=microscaling", "block_size is required")]:
        with pytest.raises(ValueError, match=msg):
            S(bad)
    with pytest.raises(ValueError):
        qt.QuantizationSpec("int8", qscheme=qt.per_tensor_symmetric)  # quant_max is required


def test_module_surface_and_state_dict():
    from torch.ao.quantization import FakeQuantizeBase
    m = qt.FusedAmaxObsFakeQuantize("e4m3")
    assert isinstance(m, FakeQuantizeBase)
    assert list(m.state_dict()) == ["fake_quant_enabled", "observer_enabled", "amax_history", "scale", "zero_point"]
    assert m.scale.shape == (1,) and m.amax_history.shape == (0,) and m.zero_point.shape == (1,)
    assert m._flags() == (False, True)
    m2 = qt.FusedAmaxObsFakeQuantize("int8", qt.per_tensor_symmetric, -128.0, 127.0, 16)
    assert m2._flags() == (True, True)
    m2.disable_observer()
    assert m2._flags() == (False, True)
    m2.observer_enabled[0] = 1  # direct buffer writes are picked up through the version counter
    assert m2._flags() == (True, True)
    # lazily shaped buffers load from a checkpoint taken after the first observed call
    sd = m2.state_dict()
    sd["scale"] = torch.tensor(0.5)
    sd["amax_history"] = torch.arange(16.0)
    sd["fake_quant_enabled"] = torch.tensor([0], dtype=torch.uint8)
    m3 = qt.FusedAmaxObsFakeQuantize("int8", qt.per_tensor_symmetric, -128.0, 127.0, 16)
    m3.load_state_dict(sd)
    assert m3.scale.shape == () and float(m3.scale) == 0.5 and m3.amax_history.shape == (16,)
    assert m3._flags() == (True, False)
    ctr = qt.get_qconfig("posit8_1", "posit8_1", None)
    assert ctr.error is torch.nn.Identity and isinstance(ctr.weight(), qt.FusedAmaxObsFakeQuantize)
    assert qt.get_qconfig(None, None, qt.QuantizationSpec.from_str("int8,qs=per_tensor_symmetric")).error().quant_max == 127.0


def test_no_cpu_fallback():
    m = qt.FusedAmaxObsFakeQuantize("posit8_1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _C.fq_forward(torch.zeros(8), torch.zeros(8), 1, 1, 8, m._fmt)
    off = qt.FusedAmaxObsFakeQuantize("posit8_1")
    off.disable_fake_quant()
    x = torch.randn(8)
    assert off(x).data_ptr() == x.data_ptr()  # both switches off: pass-through, as in the reference


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "quantized-training_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def _pow2_scale_from_table(tab, amax_bits, qmax, f32):
    """numpy restatement of mx_scale_of() (csrc/qt_block.cu) for the force_scale_power_of_two branch."""
    import math
    out = np.empty(amax_bits.size, dtype=np.uint32)
    qe = math.floor(math.log2(qmax))
    lo = -149 if f32 else -133
    for i, a in enumerate(amax_bits.tolist()):
        if a > 0x7F800000:
            out[i] = 0x3F800000
            continue
        if a == 0x7F800000:
            out[i] = a
            continue
        if a == 0:
            E = -126
        elif a >> 23:
            e = a >> 23
            E = e - 127 + (1 if (a & 0x7FFFFF) >= int(tab[e]) else 0)
        else:
            k = a.bit_length() - 1
            E = k - 149 + (1 if a >= int(tab[256 + k]) else 0)
        E -= qe
        if E < lo:
            out[i] = 0x3F800000
        elif E > 127:
            out[i] = 0x7F800000
        else:
            out[i] = (E + 127) << 23 if E >= -126 else 1 << (E + 149)
    return out


def test_block_pow2_thresholds_match_reference(golden):
    """force_scale_power_of_two: the tabulated floor(log2(amax)) reproduces calculate_mx_qparam of the reference
    on every positive bf16 amax and on fp32 values around every power of two."""
    allb = np.arange(0x8000, dtype=np.uint32) << 16
    fb = golden.mx_scale["f32_amax_bits"]
    t16 = _C.pow2_table_host(_C.QT_BF16).numpy().view(np.uint32)
    t32 = _C.pow2_table_host(_C.QT_F32).numpy().view(np.uint32)
    for qmax in golden.mx_manifest["scale_fn_quant_max"]:
        got = _pow2_scale_from_table(t16, allb, qmax, False)
        want = golden.mx_scale[f"bf16/pow2/{qmax}"].astype(np.uint32) << 16
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, ("bf16", qmax, [(hex(allb[i]), hex(got[i]), hex(want[i])) for i in bad[:6]])
        got = _pow2_scale_from_table(t32, fb, qmax, True)
        want = golden.mx_scale[f"f32/pow2/{qmax}"]
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, ("f32", qmax, [(hex(fb[i]), hex(got[i]), hex(want[i])) for i in bad[:6]])


def test_block_view():
    from quantized_training.fake_quantize import _block_view
    assert _block_view((4, 6, 128), -1, 64) == ((24, 128, 1, 1, 1), False, [4, 6, 2])
    assert _block_view((3, 160, 64), -2, 64) == ((3, 160, 64, 1, 1), False, [3, 3, 64])
    assert _block_view((2, 3, 40, 50), (-2, -1), 16) == ((6, 40, 1, 50, 1), True, [2, 3, 3, 4])
    assert _block_view((5, 7, 9, 11), (0, 2), 4) == ((1, 5, 7, 9, 11), True, [2, 7, 3, 11])
    assert _block_view((7,), 0, 32) == ((1, 7, 1, 1, 1), False, [1])
    with pytest.raises(NotImplementedError):
        _block_view((2, 2, 2), (0, 1, 2), 2)


def test_block_scaled_module_surface():
    spec = qt.QuantizationSpec.from_str("int6,qs=microscaling,bs=64,ax=-1,scale=fp8_e5m3")
    m = spec.observer_or_fake_quant_ctr(**spec.fake_quant_kwargs())
    assert m.is_block_scaled and m.quant_max == 31.0 and m.block_size == 64 and m.scale_dtype == "fp8_e5m3"
    assert m.scale_qmap.shape == (65536,) and m.calculate_qparams() is m.scale
    g = qt.QuantizationSpec.from_str("uint2,bs=64,qs=group_wise_affine,ax=-2,scale=fp8_e5m3")
    gm = g.observer_or_fake_quant_ctr(**g.fake_quant_kwargs())
    assert (gm.quant_min, gm.quant_max) == (0.0, 3.0) and len(gm.calculate_qparams()) == 2
    with pytest.raises(ValueError, match="block_size is required"):
        qt.QuantizationSpec.from_str("int6,qs=microscaling,ax=-1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(4, 64, dtype=torch.bfloat16))


def test_operator_surface_is_registered():
    """torch.ops.quantized_ops.{vmap, quantize, dequantize}: same schemas as decomposed.py:143,166-169,213-216; CUDA
    implementations only (a CPU call fails in the dispatcher, no silent fallback)."""
    for name in ("vmap", "quantize", "dequantize"):
        assert hasattr(torch.ops.quantized_ops, name)
    s = str(torch.ops.quantized_ops.quantize.default._schema)
    assert "Tensor? zero_point=None" in s and "int? block_size=None" in s and "Tensor? qmap=None" in s
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.quantized_ops.vmap(torch.zeros(8, dtype=torch.bfloat16), torch.zeros(65536, dtype=torch.bfloat16))


def test_per_tensor_pow2_scale_matches_libm():
    """force_scale_power_of_two of the per-tensor scheme: the log2-free rule the device uses equals
    2 ** ceil(log2(sf)) in fp32 (numpy / libm, what the reference's CPU run computes) on the first 64 and last 64
    mantissas of every exponent and on random scales."""
    L = _C.lib()
    tails = np.concatenate([np.arange(0, 64), np.arange(0x7FFFC0, 0x800000), [0x400000, 0x123456]]).astype(np.uint32)
    bits = (np.arange(1, 255, dtype=np.uint32)[:, None] << 23 | tails[None, :]).reshape(-1)
    rng = np.random.default_rng(0)
    bits = np.concatenate([bits, rng.integers(0x00800000, 0x7F800000, 20000, dtype=np.uint32)])
    sf = bits.view(np.float32)
    with np.errstate(over="ignore"):
        want = np.power(np.float32(2.0), np.ceil(np.log2(sf)), dtype=np.float32)
    got = np.array([L.qt_scale_pow2_host(float(v)) for v in sf], dtype=np.float32)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
    assert bad.size == 0, [(hex(bits[i]), got[i], want[i]) for i in bad[:8]]


def test_expand_helper():
    from quantized_training.decomposed import expand
    s = torch.arange(6.0).reshape(2, 3)
    e = expand(s, (2, 70), 32)
    assert e.shape == (2, 70) and torch.equal(e[:, 0], s[:, 0]) and torch.equal(e[:, 31], s[:, 0])
    assert torch.equal(e[:, 32], s[:, 1]) and torch.equal(e[:, 69], s[:, 2])
    e3 = expand(s, (5, 2, 96), 32)           # a leading axis is added and repeated like any other
    assert e3.shape == (5, 2, 96) and torch.equal(e3[4], e3[0]) and torch.equal(e3[0, :, 64], s[:, 2])
    g = torch.arange(4.0).reshape(2, 2)
    e2 = expand(g, (40, 50), 32)
    assert e2.shape == (40, 50) and e2[31, 31] == 0 and e2[32, 31] == 2 and e2[31, 32] == 1 and e2[39, 49] == 3
