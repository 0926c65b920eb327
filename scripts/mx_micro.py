"""Micro-benchmark of the block-scaled fake-quant kernels (one B200): GB/s of algorithmic traffic
(read + write + 4 B of scale per block) per spec / layout.  python scripts/mx_micro.py [--log2 28] [--only SUBSTR]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quantized-training_b200"))
import quantized_training as qt  # noqa: E402

CASES = [
    ("fp4_e2m1,qs=microscaling,bs=32,ax=-1", True),
    ("fp8_e4m3,qs=microscaling,bs=32,ax=-1", True),
    ("fp6_e3m2,qs=microscaling,bs=32,ax=-1", True),
    ("int6,qs=microscaling,bs=64,ax=-1,scale=fp8_e5m3", False),
    ("int8,qs=microscaling,bs=32,ax=-1", False),
    ("posit8_1,qs=microscaling,bs=32,ax=-1", False),
    ("int6,qs=microscaling,bs=64,ax=0,scale=fp8_e5m3", False),
    ("fp4_e2m1,qs=microscaling,bs=32,ax=0", True),
    ("int6,qs=microscaling,bs=16,ax=(0,1),scale=fp8_e5m3", False),
    ("int6,qs=microscaling,bs=64,ax=(0,1),scale=fp8_e5m3", False),
    ("fp4_e2m1,qs=microscaling,bs=32,ax=(0,1)", True),
    ("uint4,qs=group_wise_affine,bs=64,ax=-1", False),
    ("uint2,qs=group_wise_affine,bs=64,ax=0,scale=fp8_e5m3", False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=28)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = 1 << a.log2
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    x = (torch.randn(n // 4096, 4096, device=dev) * 1.7).to(dt)
    esz = x.element_size()
    for spec, pow2 in CASES:
        if a.only and a.only not in spec:
            continue
        qs = qt.QuantizationSpec.from_str(spec)
        m = qt.FusedAmaxObsFakeQuantize(**qs.fake_quant_kwargs(), force_scale_power_of_two=pow2, device=dev)
        m(x)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(a.reps):
            m(x)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / a.reps
        nblk = m.scale.numel() * (2 if qs.qscheme.value == "group_wise_affine" else 1)
        gbs = (2.0 * esz * n + 4.0 * nblk) / (ms * 1e-3) / 1e9
        print(f"{gbs:8.1f} GB/s  {ms:7.3f} ms  {spec}{' pow2' if pow2 else ''} {a.dtype}", flush=True)


if __name__ == "__main__":
    main()
