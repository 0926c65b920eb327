import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
# the drop-in package keeps the reference's import name: `import quantized_training`
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def nan_eq16(a, b):
    """bit equality of bf16 patterns with NaN == NaN (any payload / sign)."""
    a = np.asarray(a).astype(np.uint16)
    b = np.asarray(b).astype(np.uint16)
    return (a == b) | (((a & 0x7FFF) > 0x7F80) & ((b & 0x7FFF) > 0x7F80))


def nan_eq32(a, b):
    a = np.asarray(a).astype(np.uint32)
    b = np.asarray(b).astype(np.uint32)
    return (a == b) | (((a & 0x7FFFFFFF) > 0x7F800000) & ((b & 0x7FFFFFFF) > 0x7F800000))


def nan_eq(a, b):
    a = np.asarray(a)
    return nan_eq16(a, b) if a.dtype == np.uint16 else nan_eq32(a, b)


@pytest.fixture(scope="session")
def golden():
    class G:
        qmaps = np.load(os.path.join(GOLDEN, "qmaps.npz"))
        pbits = np.load(os.path.join(GOLDEN, "pbits.npz"))
        vmap32 = np.load(os.path.join(GOLDEN, "vmap32.npz"))
        fq = np.load(os.path.join(GOLDEN, "fq_cases.npz"))
        mx = np.load(os.path.join(GOLDEN, "mx_cases.npz"))
        mx_scale = np.load(os.path.join(GOLDEN, "mx_scale.npz"))
        ops = np.load(os.path.join(GOLDEN, "ops_cases.npz"))
        with open(os.path.join(GOLDEN, "manifest.json")) as f:
            manifest = json.load(f)
        with open(os.path.join(GOLDEN, "mx_manifest.json")) as f:
            mx_manifest = json.load(f)
    return G


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
