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
