"""GPU parity tests of the tcgen05 quantized GEMM (qt_gemm_nt) through the C ABI.

Floating-point kernel: compared with an fp64 matmul of the SAME quantized operands plus the same epilogue.
Tolerance (written here, as the task demands): the kernel accumulates in fp32 and rounds once to bf16, so
|got - ref| <= 2^-8 * |ref| + 2^-8 * rms(ref)   (half a bf16 ulp is 2^-9 relative; the rms term covers
cancellation in the fp32 accumulation of K products)."""
import pytest
import torch

import quantized_training as qt
from quantized_training import _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def check(got, ref):
    ref = ref.double()
    err = (got.double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -8 * ref.pow(2).mean().sqrt()
    bad = err > tol
    assert not bad.any(), f"{int(bad.sum())} of {bad.numel()} outside tolerance, max err {float(err.max())}"


def quantized_operand(shape, spec, gen, scale=1.0):
    x = (torch.randn(shape, generator=gen) * scale).to(torch.bfloat16).to(DEV)
    return qt.FusedAmaxObsFakeQuantize(spec, device=DEV)(x)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (256, 512, 192), (1024, 4096, 4096),
                                   (6144, 768, 768), (100, 264, 72), (1, 8, 8), (333, 1000, 1111 // 8 * 8),
                                   (2048, 3072, 768)])
def test_linear_shapes(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = quantized_operand((M, K), "posit8_1", g)
    w = quantized_operand((N, K), "posit8_1", g, 0.05)
    out = _C.gemm_nt(a, w)
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    check(out, a.double() @ w.double().t())


def test_epilogue_bias_act_residual_alpha():
    g = torch.Generator().manual_seed(5)
    M, N, K = 384, 768, 512
    a = quantized_operand((M, K), "e4m3", g)
    w = quantized_operand((N, K), "e4m3", g, 0.05)
    bias = torch.randn(N, generator=g).to(torch.bfloat16).to(DEV)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(DEV)
    base = a.double() @ w.double().t()
    check(_C.gemm_nt(a, w, bias=bias), base + bias.double())
    check(_C.gemm_nt(a, w, alpha=0.125), base * 0.125)
    check(_C.gemm_nt(a, w, bias=bias, activation="relu"), torch.relu(base + bias.double()))
    check(_C.gemm_nt(a, w, bias=bias, activation="gelu"), torch.nn.functional.gelu(base + bias.double()))
    check(_C.gemm_nt(a, w, activation="silu"), torch.nn.functional.silu(base))
    check(_C.gemm_nt(a, w, bias=bias, residual=res), base + bias.double() + res.double())
    check(_C.gemm_nt(a, w, alpha=0.5, bias=bias, activation="gelu", residual=res),
          torch.nn.functional.gelu(base * 0.5 + bias.double()) + res.double())


def test_batched_attention_scores():
    """q k^T for [B, H, S, D] operands that are transposed views of [B, S, H*D] projections (no copies)."""
    g = torch.Generator().manual_seed(9)
    B, H, S, D = 2, 12, 384, 64
    q = quantized_operand((B, S, H * D), "posit8_1", g).view(B, S, H, D).transpose(1, 2)
    k = quantized_operand((B, S, H * D), "posit8_1", g).view(B, S, H, D).transpose(1, 2)
    out = _C.gemm_nt(q, k, alpha=D ** -0.5)
    assert out.shape == (B, H, S, S)
    check(out, (q.double() @ k.double().transpose(-1, -2)) * D ** -0.5)
    # probabilities x values needs V^T as the K-major operand
    p = torch.softmax(out.float(), -1).to(torch.bfloat16)
    v = quantized_operand((B, H, S, D), "posit8_1", g)
    ctx = _C.gemm_nt(p, v.transpose(-1, -2).contiguous())
    check(ctx, p.double() @ v.double())


@pytest.mark.parametrize("kind", ["e4m3", "e5m2", "e4m3_e5m2", "e5m2_e4m3"])
def test_fp8_codes(kind):
    g = torch.Generator().manual_seed(11)
    M, N, K = 512, 768, 1024
    ta = torch.float8_e5m2 if kind.startswith("e5m2") else torch.float8_e4m3fn
    tb = torch.float8_e5m2 if kind.endswith("e5m2") else torch.float8_e4m3fn
    a8 = torch.randn(M, K, generator=g).to(DEV).to(ta)
    b8 = (torch.randn(N, K, generator=g) * 0.1).to(DEV).to(tb)
    op = {"e4m3": _C.GEMM_E4M3, "e5m2": _C.GEMM_E5M2, "e4m3_e5m2": _C.GEMM_E4M3_E5M2, "e5m2_e4m3": _C.GEMM_E5M2_E4M3}[kind]
    out = _C.gemm_nt(a8.view(torch.uint8), b8.view(torch.uint8), operand_type=op)
    check(out, a8.double() @ b8.double().t())


def test_argument_errors():
    a = torch.zeros(16, 20, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ValueError):
        _C.gemm_nt(a, torch.zeros(8, 24, dtype=torch.bfloat16, device=DEV))      # K mismatch
    with pytest.raises(ValueError):
        _C.gemm_nt(a, torch.zeros(8, 20, dtype=torch.bfloat16, device=DEV))      # lda = 20 is not 16-byte aligned
    with pytest.raises(TypeError):
        _C.gemm_nt(a.float(), a.float())
