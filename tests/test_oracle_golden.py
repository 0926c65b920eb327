"""Pin the CPU oracle (oracle/qt_oracle.c) to outputs of the unmodified reference
(tests/golden/*.npz, made by tests/golden/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import nan_eq, nan_eq16, nan_eq32


def test_tables_bit_exact(golden, oracle):
    assert len(golden.qmaps.files) >= 30
    for d in golden.qmaps.files:
        t = oracle.qmap(d)
        bad = np.nonzero(~nan_eq16(t, golden.qmaps[d]))[0]
        assert bad.size == 0, (d, [(hex(i), hex(t[i]), hex(golden.qmaps[d][i])) for i in bad[:8]])


def test_posit_codes_bit_exact(golden, oracle):
    for k in golden.pbits.files:
        n, es = k[len("posit"):].split("_")
        vals, pb = oracle.posit(int(n), int(es))
        assert np.array_equal(pb, golden.pbits[k]), k
        assert nan_eq16(vals, golden.qmaps[k]).all() if k in golden.qmaps.files else True


@pytest.mark.parametrize("bad", ["nf4", "float8", "int", "posit8", "fp8_e4m2", "e3m4", "", "posit8_1 "])
def test_unsupported_dtype_raises(oracle, bad):
    with pytest.raises(ValueError):
        oracle.qmap(bad)


def test_vmap_fp32_round_to_odd(golden, oracle):
    xb = golden.vmap32["x_bits"]
    x = xb.view(np.float32)
    for d in golden.vmap32.files:
        if d == "x_bits":
            continue
        y = oracle.vmap(x, oracle.qmap(d))
        yb = y.view(np.uint32)
        assert np.all((yb & 0xFFFF) == 0)
        assert nan_eq16((yb >> 16).astype(np.uint16), golden.vmap32[d]).all(), d


def test_fake_quant_sequences(golden, oracle):
    """Delayed scaling: outputs, scale and amax history after every call."""
    for case in golden.manifest["fq_cases"]:
        name = case["name"]
        fq = oracle.FakeQuant(case["spec"].split(",")[0], qscheme=case["qscheme"], quant_max=case["quant_max"],
                              amax_history_len=case["amax_history_len"], ch_axis=case["ch_axis"],
                              force_scale_power_of_two=case["force_scale_power_of_two"])
        f32 = case["dtype"] == "fp32"
        for k in range(case["calls"]):
            x = golden.fq[f"{name}/x{k}"]
            xin = x.view(np.float32) if f32 else x
            y = fq(xin, case["shape"])
            yb = y.view(np.uint32) if f32 else y
            assert nan_eq(yb, golden.fq[f"{name}/y{k}"]).all(), (name, k)
            assert nan_eq32(fq.scale.view(np.uint32), golden.fq[f"{name}/scale{k}"]).all(), (name, k, "scale")
            if case["qscheme"] is not None:
                assert nan_eq32(fq.history.view(np.uint32), golden.fq[f"{name}/hist{k}"]).all(), (name, k, "hist")
        fq.observer_enabled = False
        x = golden.fq[f"{name}/x_obsoff"]
        y = fq(x.view(np.float32) if f32 else x, case["shape"])
        yb = y.view(np.uint32) if f32 else y
        assert nan_eq(yb, golden.fq[f"{name}/y_obsoff"]).all(), (name, "obsoff")
        assert nan_eq32(fq.scale.view(np.uint32), golden.fq[f"{name}/scale_obsoff"]).all()


def _as_input(bits, dtype):
    return bits.view(np.float32) if dtype == "f32" else bits


def test_block_scaled_cases(golden, oracle):
    """microscaling / group_wise_affine forward: outputs, scale and zero_point of the reference run."""
    assert len(golden.mx_manifest["cases"]) >= 20
    for case in golden.mx_manifest["cases"]:
        name = case["name"]
        x = _as_input(golden.mx[f"{name}/x"], case["dtype"])
        stab = oracle.qmap(case["scale_dtype"]) if case["scale_dtype"] else None
        if case["qscheme"] == "microscaling":
            y, s = oracle.mx_fake_quant(x, case["shape"], case["ch_axis"], case["block_size"], case["quant_max"],
                                        oracle.qmap(case["element"]), case["force_scale_power_of_two"], stab)
        else:
            y, s, zp = oracle.gwa_fake_quant(x, case["shape"], case["ch_axis"], case["block_size"],
                                             case["quant_min"], case["quant_max"], stab)
            assert nan_eq32(zp.reshape(-1).view(np.uint32), golden.mx[f"{name}/zero_point"]).all(), (name, "zp")
        assert list(s.shape) == case["scale_shape"], name
        assert nan_eq32(s.reshape(-1).view(np.uint32), golden.mx[f"{name}/scale"]).all(), (name, "scale")
        yb = y.view(np.uint32) if case["dtype"] == "f32" else y
        assert nan_eq(yb.reshape(-1), golden.mx[f"{name}/y"].reshape(-1)).all(), name


def test_mx_scale_function(golden, oracle):
    """calculate_mx_qparam on one-element blocks: every positive bf16 amax, fp32 values around every power of
    two -- pins the dtype-dependent floor(log2()) of force_scale_power_of_two."""
    e5m3 = oracle.qmap("fp8_e5m3")
    ident = np.arange(65536, dtype=np.uint16)  # get_quantization_map(None): identity table
    allb = np.arange(0x8000, dtype=np.uint16)
    fb = golden.mx_scale["f32_amax_bits"]
    for qmax in golden.mx_manifest["scale_fn_quant_max"]:
        for mode, pow2, stab in (("pow2", True, None), ("amax", False, None), ("e5m3", False, e5m3)):
            _, s = oracle.mx_fake_quant(allb, (allb.size, 1), -1, 1, qmax, ident, pow2, stab)
            # the bf16 reference returns the scale as bf16: compare through its fp32 widening
            want = golden.mx_scale[f"bf16/{mode}/{qmax}"].astype(np.uint32) << 16
            assert nan_eq32(s.reshape(-1).view(np.uint32), want).all(), ("bf16", mode, qmax)
            _, s = oracle.mx_fake_quant(fb.view(np.float32), (fb.size, 1), -1, 1, qmax, ident, pow2, stab)
            bad = np.nonzero(~nan_eq32(s.reshape(-1).view(np.uint32), golden.mx_scale[f"f32/{mode}/{qmax}"]))[0]
            assert bad.size == 0, ("f32", mode, qmax, [hex(fb[i]) for i in bad[:8]])
