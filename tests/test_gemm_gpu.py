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


@pytest.fixture(autouse=True, params=["auto", "pair"])
def cta_pairs(request, monkeypatch):
    """Every test of this file runs twice: with the launcher's own choice between single CTAs and CTA pairs
    (cta_group::2, 256-row tiles), and with pairs forced wherever the kernel has them (everything but the causal
    schedules and the one-byte-code operands) -- including shapes with an odd number of row tiles or a single one,
    where the second CTA of a pair works on rows outside the problem."""
    if request.param == "pair":
        monkeypatch.setenv("QT_GEMM_PAIR", "1")
    yield request.param


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


@pytest.mark.parametrize("B,M,N,K", [(32, 1024, 1024, 128), (192, 384, 384, 64), (192, 384, 64, 384), (3, 130, 72, 40),
                                     (5, 128, 64, 64), (2, 257, 136, 264)])
def test_batched_tile_widths(B, M, N, K):
    """Short-K batched products: the launcher narrows the tile (64 / 128 / 256 columns) to fill the SMs; every
    width, with ragged M and N (rows / columns clipped by the TMA store), must give the same numbers."""
    g = torch.Generator().manual_seed(B * 11 + M + N + K)
    a = quantized_operand((B, M, K), "e4m3", g)
    w = quantized_operand((B, N, K), "e4m3", g, 0.25)
    out = _C.gemm_nt(a, w)
    assert out.shape == (B, M, N)
    check(out, a.double() @ w.double().transpose(-1, -2))


def test_two_level_batch_strided_attention_layouts():
    """[B, S, H*D] projections viewed as [B, H, S, D] (two batch strides, no copies) and a context written
    straight into the [B, S, H*D] layout the output projection reads."""
    g = torch.Generator().manual_seed(21)
    B, H, S, D = 3, 4, 200, 64
    q = quantized_operand((B, S, H * D), "posit8_1", g).view(B, S, H, D).transpose(1, 2)
    k = quantized_operand((B, S, H * D), "posit8_1", g).view(B, S, H, D).transpose(1, 2)
    assert not q.is_contiguous()
    scores = _C.gemm_nt(q, k, alpha=0.125)
    check(scores, (q.double() @ k.double().transpose(-1, -2)) * 0.125)
    p = quantized_operand((B, H, S, S), "posit8_1", g, 0.1)
    vt = quantized_operand((B, H, D, S), "posit8_1", g)
    ctx = torch.zeros(B, S, H * D, dtype=torch.bfloat16, device=DEV)
    out_view = ctx.view(B, S, H, D).transpose(1, 2)              # [B, H, S, D] strided destination
    _C.gemm_nt(p, vt, out=out_view)
    check(out_view, p.double() @ vt.double().transpose(-1, -2))


def test_output_is_fully_written_and_nothing_else():
    """The epilogue stores 32 x 64 boxes through TMA: a strided C (ldc > N) must keep its padding untouched."""
    g = torch.Generator().manual_seed(3)
    M, N, K = 200, 72, 128
    a = quantized_operand((M, K), "posit8_1", g)
    w = quantized_operand((N, K), "posit8_1", g, 0.1)
    big = torch.full((M + 5, 128), 7.0, dtype=torch.bfloat16, device=DEV)
    view = big[:M, :N]
    _C.gemm_nt(a, w, out=view)
    check(view, a.double() @ w.double().t())
    assert bool((big[M:] == 7.0).all()) and bool((big[:, N:] == 7.0).all())


def test_epilogue_bias_act_residual_alpha():
    g = torch.Generator().manual_seed(5)
    M, N, K = 384, 768, 512
    a = quantized_operand((M, K), "e4m3", g)
    w = quantized_operand((N, K), "e4m3", g, 0.05)
    bias = torch.randn(N, generator=g).to(torch.bfloat16).to(DEV)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(DEV)
    base = a.double() @ w.double().t()
    r16 = lambda x: x.to(torch.bfloat16).double()      # the reference materialises bf16 tensors between its ops
    F = torch.nn.functional
    check(_C.gemm_nt(a, w, bias=bias), base + bias.double())
    check(_C.gemm_nt(a, w, alpha=0.125), base * 0.125)
    check(_C.gemm_nt(a, w, bias=bias, activation="relu"), torch.relu(r16(base + bias.double())))
    check(_C.gemm_nt(a, w, bias=bias, activation="gelu"), F.gelu(r16(base + bias.double())))
    check(_C.gemm_nt(a, w, activation="silu"), F.silu(r16(base)))
    check(_C.gemm_nt(a, w, bias=bias, residual=res), r16(base + bias.double()) + res.double())
    check(_C.gemm_nt(a, w, alpha=0.5, bias=bias, activation="gelu", residual=res),
          r16(F.gelu(r16(base * 0.5 + bias.double()))) + res.double())
    # the epilogue rounds where the reference's separate bf16 ops do: bit-identical to that chain when the
    # accumulation is exact (small K, e4m3 operands: every partial sum is representable in fp32)
    a2, w2 = a[:, :64].contiguous(), w[:, :64].contiguous()
    lin = torch.nn.functional.linear(a2, w2, bias)
    got = _C.gemm_nt(a2, w2, bias=bias, activation="gelu", residual=res)
    assert torch.equal(got, torch.nn.functional.gelu(lin) + res)


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


def _fq_table(oracle, dtype):
    import numpy as np
    table = torch.from_numpy(oracle.qmap(dtype).view(np.int16)).view(torch.bfloat16).to(DEV)
    return lambda x: table[(x.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF)].view(x.shape)


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "int8", "fp8_e5m2"])
def test_epilogue_requantization_is_the_consumers_fake_quant(oracle, spec):
    """C = fq(bf16(A B^T + bias)) in the epilogue == the separate fake-quant pass on the plain GEMM's output, bit for
    bit (same accumulator, same rounding to bf16, same rounding engine); fp8 codes decode to the same values."""
    g = torch.Generator().manual_seed(31)
    M, N, K = 300, 328, 256
    a = quantized_operand((M, K), "e4m3", g)
    w = quantized_operand((N, K), "e4m3", g, 0.3)
    bias = torch.randn(N, generator=g).to(torch.bfloat16).to(DEV)
    m = qt.FusedAmaxObsFakeQuantize(spec, device=DEV)
    plain = _C.gemm_nt(a, w, bias=bias)
    want = _fq_table(oracle, spec)(plain)
    got = _C.gemm_nt(a, w, bias=bias, fq=(m._fmt, m.lut))
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    if m.fp8_kind is not None:
        codes = _C.gemm_nt(a[:, :], w[:320], bias=bias[:320].contiguous(), fq=(m._fmt, m.lut), out_codes=True)
        tdt = torch.float8_e4m3fn if m.fp8_kind == "e4m3" else torch.float8_e5m2
        assert codes.dtype == torch.uint8 and torch.equal(codes.view(tdt).to(torch.bfloat16).view(torch.int16),
                                                          want[:, :320].contiguous().view(torch.int16))
    # batched, strided destination (the PV product writing a fake-quantized context into [B, S, H*D])
    B, H, S, D = 2, 4, 136, 64
    p_ = quantized_operand((B, H, S, S), "e4m3", g, 0.1)
    vt = quantized_operand((B, H, D, S), "e4m3", g)
    ctx = torch.zeros(B, S, H * D, dtype=torch.bfloat16, device=DEV)
    _C.gemm_nt(p_, vt, out=ctx.view(B, S, H, D).transpose(1, 2), fq=(m._fmt, m.lut))
    ref = _fq_table(oracle, spec)(_C.gemm_nt(p_, vt))
    assert torch.equal(ctx.view(B, S, H, D).transpose(1, 2).contiguous().view(torch.int16), ref.view(torch.int16))


@pytest.mark.parametrize("spec,codes", [("posit8_1", False), ("e4m3", False), ("e4m3", True)])
@pytest.mark.parametrize("with_bias", [False, True])
def test_gated_epilogue_matches_the_mlp_op_chain(oracle, spec, codes, with_bias):
    """silu(gate) * up + fake quant inside the gate|up GEMM (weights interleaved in blocks of 64 rows) vs the HF
    LlamaMLP op chain on the two separate projections: act_fn(gate_proj(x)) * up_proj(x), then down_proj's input hook.
    Tolerance: SiLU goes through exp, so as for qt_act_mul_fq >= 99.5 % of the values are bit-identical and none is
    further than one grid step."""
    g = torch.Generator().manual_seed(41)
    M, K, I = 200, 256, 704                       # I % 64 == 0
    x = quantized_operand((M, K), "e4m3", g)
    wg = quantized_operand((I, K), "e4m3", g, 0.2)
    wu = quantized_operand((I, K), "e4m3", g, 0.2)
    bg = torch.randn(I, generator=g).to(torch.bfloat16).to(DEV) if with_bias else None
    bu = torch.randn(I, generator=g).to(torch.bfloat16).to(DEV) if with_bias else None
    inter = lambda t, u: torch.stack((t.view(I // 64, 64, *t.shape[1:]), u.view(I // 64, 64, *u.shape[1:])), 1).reshape(2 * I, *t.shape[1:])
    w = inter(wg, wu).contiguous()
    b = inter(bg, bu).contiguous() if with_bias else None
    m = qt.FusedAmaxObsFakeQuantize(spec, device=DEV)
    got = _C.gemm_nt(x, w, bias=b, activation="silu", glu=True, fq=(m._fmt, m.lut), out_codes=codes)
    gate, up = _C.gemm_nt(x, wg, bias=bg), _C.gemm_nt(x, wu, bias=bu)
    want = _fq_table(oracle, spec)(torch.nn.functional.silu(gate) * up)
    if codes:
        got = got.view(torch.float8_e4m3fn).to(torch.bfloat16)
    assert got.shape == (M, I)
    same = got.view(torch.int16) == want.view(torch.int16)
    assert float(same.float().mean()) >= 0.995
    gf, wf = got.float(), want.float()
    assert not bool((~same & ((gf - wf).abs() > 0.26 * wf.abs().clamp_min(1e-30)) & ((gf - wf).abs() > 2.0 ** -14)).any())


# ---- MN-major operands: the backward products of a Linear and x @ y, read as the tensors lie (no transpose copies)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 192), (2048, 768, 768), (2048, 3072, 768),
                                   (100, 264, 72), (8, 8, 8), (333, 1000, 1112), (1024, 4096, 4096)])
def test_dgrad_wgrad_mn_major(M, N, K):
    """y = x W^T with x [M, K], W [N, K], g = dL/dy [M, N].
    dgrad gx = g W       : A = g K-major (contraction over N), B = W stored [N, K] = [k', n'] -> MN-major
    wgrad gW = g^T x     : A = g stored [M, N] = [k', m'] -> MN-major, B = x stored [M, K] = [k', n'] -> MN-major"""
    gen = torch.Generator().manual_seed(M + 3 * N + 5 * K)
    x = quantized_operand((M, K), "e4m3", gen)
    w = quantized_operand((N, K), "e4m3", gen, 0.05)
    g = quantized_operand((M, N), "e5m2", gen, 0.01)
    gx = _C.gemm_nt(g, w, b_mn=True)
    assert gx.shape == (M, K)
    check(gx, g.double() @ w.double())
    gw = _C.gemm_nt(g, x, a_mn=True, b_mn=True)
    assert gw.shape == (N, K)
    check(gw, g.double().t() @ x.double())
    # A MN-major alone (x^T @ B^T with B K-major)
    if M % 8 == 0:   # row stride of a K-major B = M elements: the raw entry point wants 16-byte multiples
        b = quantized_operand((N, M), "e4m3", gen, 0.1)      # [n', k'] K-major, contraction over M
        out = _C.gemm_nt(x, b, a_mn=True)                    # x stored [k'=M, m'=K]
        assert out.shape == (K, N)
        check(out, x.double().t() @ b.double().t())


@pytest.mark.parametrize("op,ta,tb", [(_C.GEMM_E4M3, "e4m3", "e4m3"), (_C.GEMM_E5M2_E4M3, "e5m2", "e4m3"),
                                      (_C.GEMM_E4M3_E5M2, "e4m3", "e5m2")])
@pytest.mark.parametrize("M,N,K", [(256, 512, 384), (208, 272, 144), (1024, 4096, 1024)])
def test_fp8_codes_mn_major(op, ta, tb, M, N, K):
    """One-byte operands read MN-major (fp8 dgrad / wgrad): same products as the bf16 path."""
    f8 = {"e4m3": torch.float8_e4m3fn, "e5m2": torch.float8_e5m2}
    gen = torch.Generator().manual_seed(M + N + K)
    a = quantized_operand((M, K), ta, gen)               # [m, k]
    bt = quantized_operand((K, N), tb, gen, 0.1)         # B stored [k, n]: MN-major
    at = quantized_operand((K, M), ta, gen)              # A stored [k, m]: MN-major
    codes = lambda t, kind: t.to(f8[kind]).view(torch.uint8)
    if N % 16 == 0:
        check(_C.gemm_nt(codes(a, ta), codes(bt, tb), operand_type=op, b_mn=True), a.double() @ bt.double())
    if M % 16 == 0 and N % 16 == 0:
        check(_C.gemm_nt(codes(at, ta), codes(bt, tb), operand_type=op, a_mn=True, b_mn=True),
              at.double().t() @ bt.double())


def test_batched_mn_major_matmul_backward():
    """torch.matmul(x, y) with x [B, H, S, S'], y [B, H, S', D] (attention P V): forward reads y MN-major;
    gx = g y^T is K-major on both; gy = x^T g reads both MN-major.  Strided [B, S, H, D] views included."""
    gen = torch.Generator().manual_seed(9)
    B, H, S, D = 2, 3, 136, 64
    p = quantized_operand((B, H, S, S), "e4m3", gen, 0.1)
    v = quantized_operand((B, S, H * D), "e4m3", gen).view(B, S, H, D).transpose(1, 2)   # [B, H, S, D] strided
    g = quantized_operand((B, H, S, D), "e5m2", gen, 0.01)
    check(_C.gemm_nt(p, v, b_mn=True), p.double() @ v.double())
    check(_C.gemm_nt(g, v), g.double() @ v.double().transpose(-1, -2))
    gy = _C.gemm_nt(p, g, a_mn=True, b_mn=True)
    assert gy.shape == (B, H, S, D)
    check(gy, p.double().transpose(-1, -2) @ g.double())


def _grad_check(got, ref):
    check(got, ref)


@pytest.mark.parametrize("shape,N,bias", [((4, 128, 768), 768, True), ((2048, 768), 3072, True), ((16, 768), 3, True),
                                           ((3, 50, 72), 40, False), ((2, 7, 20), 12, True)])
def test_ops_linear_autograd_on_kernel(shape, N, bias):
    """ops.linear forward + dgrad + wgrad on the tcgen05 kernel (MN-major operands, zero-padded odd shapes such as a
    3-class classifier head) against fp64 autograd of the same operands."""
    from quantized_training import ops
    gen = torch.Generator().manual_seed(sum(shape) + N)
    K = shape[-1]
    x = quantized_operand(shape, "e4m3", gen).requires_grad_(True)
    w = quantized_operand((N, K), "e4m3", gen, 0.05).requires_grad_(True)
    b = torch.randn(N, generator=gen).to(torch.bfloat16).to(DEV).requires_grad_(True) if bias else None
    g = quantized_operand((*shape[:-1], N), "e5m2", gen, 0.01)
    y = ops.linear(x, w, b)
    y.backward(g)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(g.double())
    check(y.detach(), yd.detach())
    check(x.grad, xd.grad)
    check(w.grad, wd.grad)
    if bias:
        check(b.grad, bd.grad)


@pytest.mark.parametrize("transposed_y", [False, True])
def test_ops_matmul_autograd_on_kernel(transposed_y):
    from quantized_training import ops
    gen = torch.Generator().manual_seed(77 + transposed_y)
    B, H, S, D = 2, 4, 136, 64
    if transposed_y:   # scores = q @ k^T
        x = quantized_operand((B, H, S, D), "e4m3", gen).requires_grad_(True)
        yb = quantized_operand((B, H, S, D), "e4m3", gen).requires_grad_(True)
        y = yb.transpose(-1, -2)
    else:              # context = p @ v
        x = quantized_operand((B, H, S, S), "e4m3", gen, 0.1).requires_grad_(True)
        yb = quantized_operand((B, H, S, D), "e4m3", gen).requires_grad_(True)
        y = yb
    out = ops.matmul(x, y)
    g = quantized_operand(tuple(out.shape), "e5m2", gen, 0.01)
    out.backward(g)
    xd, ybd = x.detach().double().requires_grad_(True), yb.detach().double().requires_grad_(True)
    outd = xd @ (ybd.transpose(-1, -2) if transposed_y else ybd)
    outd.backward(g.double())
    check(out.detach(), outd.detach())
    check(x.grad, xd.grad)
    check(yb.grad, ybd.grad)


def test_fp8_linear_backward_on_fp8_tensor_cores():
    """Bare e4m3 forward + bare e5m2 gradient: dgrad / wgrad as QT_GEMM_E5M2_E4M3 products of one-byte codes."""
    from quantized_training import ops
    gen = torch.Generator().manual_seed(123)
    M, N, K = 256, 384, 512
    x = quantized_operand((M, K), "e4m3", gen).requires_grad_(True)
    w = (torch.randn(N, K, generator=gen) * 0.05).to(torch.bfloat16).to(DEV).requires_grad_(True)
    wq = qt.FusedAmaxObsFakeQuantize("e4m3", device=DEV)
    g = quantized_operand((M, N), "e5m2", gen, 0.01)
    y = ops.linear_fp8(x, w, None, wq, "e4m3", g_kind="e5m2")
    y.backward(g)
    wv = wq(w.detach()).double()
    check(y.detach(), x.detach().double() @ wv.t())
    check(x.grad, g.double() @ wv)
    check(w.grad, g.double().t() @ x.detach().double())


def test_ops_raise_instead_of_falling_back():
    from quantized_training import ops
    x = torch.randn(8, 16, device=DEV)
    w = torch.randn(8, 16, device=DEV)
    with pytest.raises(TypeError):
        ops.linear(x, w)
    with pytest.raises(TypeError):
        ops.matmul(x, w.t())
    with pytest.raises(RuntimeError):
        ops.linear(x.cpu().bfloat16(), w.cpu().bfloat16())


# ---- one-byte code operands decoded to bf16 inside the kernel (QT_GEMM_CODE8_B / _AB)
@pytest.mark.parametrize("spec", ["posit8_1", "posit8_0", "int8", "fp6_e3m2", "posit6_1", "uint4"])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1024, 4096, 1024), (100, 264, 80), (8, 4096, 4096), (333, 1000, 1104),
                                   (256, 64, 512), (130, 136, 4096)])
def test_code8_operands_are_bit_identical_to_bf16_operands(spec, M, N, K):
    """Weights (and activations) stored as one-byte codes, decoded exactly to bf16 in shared memory by the kernel's
    decode warps: the tensor cores see the same bf16 tiles as on the bf16-operand path, so the results are
    BIT-IDENTICAL, not merely close -- and codes decode to the fake-quantized values exactly."""
    gen = torch.Generator().manual_seed(M + N + K)
    fqm = qt.FusedAmaxObsFakeQuantize(spec, device=DEV)
    scale_x = 4.0 if spec in ("int8", "uint4") else 1.0
    x = ((torch.randn(M, K, generator=gen) * scale_x).to(torch.bfloat16)).to(DEV)
    w = ((torch.randn(N, K, generator=gen) * scale_x * 0.5).to(torch.bfloat16)).to(DEV)
    xq, wq = fqm(x), fqm(w)
    lut = _C.code_table_host(fqm._fmt).to(DEV)
    wc = _C.quantize_codes8(w, torch.empty(N, K, dtype=torch.uint8, device=DEV), fqm._fmt)
    xc = _C.quantize_codes8(x, torch.empty(M, K, dtype=torch.uint8, device=DEV), fqm._fmt)
    # decode(code) == the fake-quantized value (sign of zero aside)
    dec = lut[wc.long()]
    assert bool(((dec == wq) | ((dec == 0) & (wq == 0))).all())
    bias = torch.randn(N, generator=gen).to(torch.bfloat16).to(DEV)
    want = _C.gemm_nt(xq, wq, bias=bias, alpha=0.5)
    got_b = _C.gemm_nt(xq, wc, bias=bias, alpha=0.5, operand_type=_C.GEMM_CODE8_B, code_lut=lut)
    assert torch.equal(got_b, want), float((got_b.double() - want.double()).abs().max())
    got_ab = _C.gemm_nt(xc, wc, bias=bias, alpha=0.5, operand_type=_C.GEMM_CODE8_AB, code_lut=lut)
    assert torch.equal(got_ab, want)
    check(got_b, (xq.double() @ wq.double().t()) * 0.5 + bias.double())


def test_code8_batched_and_residual():
    gen = torch.Generator().manual_seed(5)
    fqm = qt.FusedAmaxObsFakeQuantize("posit8_1", device=DEV)
    B, M, N, K = 3, 200, 136, 144
    x = fqm(torch.randn(B, M, K, generator=gen).to(torch.bfloat16).to(DEV))
    w = torch.randn(B, N, K, generator=gen).to(torch.bfloat16).to(DEV)
    res = torch.randn(B, M, N, generator=gen).to(torch.bfloat16).to(DEV)
    lut = _C.code_table_host(fqm._fmt).to(DEV)
    wc = _C.quantize_codes8(w, torch.empty(B, N, K, dtype=torch.uint8, device=DEV), fqm._fmt)
    want = _C.gemm_nt(x, fqm(w), residual=res)
    got = _C.gemm_nt(x, wc, residual=res, operand_type=_C.GEMM_CODE8_B, code_lut=lut)
    assert torch.equal(got, want)


@pytest.mark.parametrize("bn", [16, 48, 80, 112, 144, 176, 208, 224, 240])
def test_tile_widths_that_are_not_multiples_of_64(bn, monkeypatch):
    """The tile width is any multiple of 16 (chosen so that the tiles fill the 148 SMs in the fewest waves): the partial
    last 64-column chunk of such a tile is written with plain 16-byte stores instead of a TMA box.  Forced here through
    QT_GEMM_BN on ragged shapes, every epilogue variant of the plain bf16 output, bf16 and fp8 operands."""
    monkeypatch.setenv("QT_GEMM_BN", str(bn))
    gen = torch.Generator().manual_seed(bn)
    for (M, N, K) in [(300, 1000, 136), (128, 8 * bn + 8, 64), (257, 2 * bn, 200)]:
        a = quantized_operand((M, K), "e4m3", gen)
        w = quantized_operand((N, K), "e4m3", gen, 0.05)
        bias = torch.randn(N, generator=gen).to(torch.bfloat16).to(DEV)
        res = torch.randn(M, N, generator=gen).to(torch.bfloat16).to(DEV)
        base = a.double() @ w.double().t()
        r16 = lambda x: x.to(torch.bfloat16).double()
        big = torch.full((M + 3, N + 16), 7.0, dtype=torch.bfloat16, device=DEV)
        _C.gemm_nt(a, w, out=big[:M, :N])
        check(big[:M, :N], base)
        assert bool((big[M:] == 7.0).all()) and bool((big[:, N:] == 7.0).all())      # nothing outside C is touched
        check(_C.gemm_nt(a, w, bias=bias, activation="gelu", residual=res),
              r16(torch.nn.functional.gelu(r16(base + bias.double()))) + res.double())
        if K % 16 == 0:
            c8 = lambda t: t.to(torch.float8_e4m3fn).view(torch.uint8)
            check(_C.gemm_nt(c8(a), c8(w), operand_type=_C.GEMM_E4M3, alpha=0.5), base * 0.5)
    B, Mb, Nb, Kb = 3, 130, 3 * bn + 24, 72
    ab = quantized_operand((B, Mb, Kb), "e4m3", gen)
    wb = quantized_operand((B, Nb, Kb), "e4m3", gen, 0.1)
    check(_C.gemm_nt(ab, wb), ab.double() @ wb.double().transpose(-1, -2))
