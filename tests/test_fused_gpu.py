"""GPU parity tests of the fused producer + fake-quant kernels (qt_fused.cu) through the C ABI.

Checker: the reference's own op chain restated with torch bf16 ops on the same device (HF modeling code: mul, add,
softmax, RMSNorm / LayerNorm, SiLU, rotary embedding) with the fake-quant steps done as lookups in the ORACLE's
65 536-entry tables (bit-exact fake quant).  Tolerances, per kernel:
  * elementwise chains without transcendental functions (rope, transpose, pre-softmax steps): bit-exact;
  * chains through exp / rsqrt / row sums (softmax, norms, SiLU): the fused kernel rounds to bf16 at the same points
    as the chain but sums rows in a different order and its exp may differ in the last ulp, so a value that sits on a
    bf16 (or format) rounding boundary can land on the neighbouring grid point: >= 99 % of the elements must be
    bit-identical and every element must be within ONE step of the coarser grid (|a - b| <= 2^-2 |b| for 8-bit formats
    is a generous statement of "adjacent code"; checked as relative error <= 0.26 or absolute <= tiny)."""
import numpy as np
import pytest
import torch

import quantized_training as qt
from quantized_training import _C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def table_fq(oracle, dtype):
    table = torch.from_numpy(oracle.qmap(dtype).view(np.int16)).view(torch.bfloat16).to(DEV)
    return lambda t: table[(t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF)].view(t.shape)


def fmt_lut(dtype):
    m = qt.FusedAmaxObsFakeQuantize(dtype, device=DEV)
    return m._fmt, m.lut


def bits(t):
    return t.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF


def assert_close_codes(got, want, min_equal=0.99):
    same = bits(got) == bits(want)
    both_nan = torch.isnan(got.float()) & torch.isnan(want.float())
    frac = float((same | both_nan).float().mean())
    assert frac >= min_equal, f"only {frac:.4f} of the elements are bit-identical"
    g, w = got.float(), want.float()
    bad = ~(same | both_nan) & ((g - w).abs() > 0.26 * w.abs().clamp_min(1e-30)) & ((g - w).abs() > 2.0 ** -14)
    assert not bool(bad.any()), f"{int(bad.sum())} elements further than one grid step apart"


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "int8"])
@pytest.mark.parametrize("cols,points", [(1024, _C.FQ_POST), (384, _C.FQ_POST), (128, 7), (2048, _C.FQ_POST | _C.FQ_MID),
                                         (4096, _C.FQ_POST), (72, _C.FQ_PRE)])
def test_softmax_chain(oracle, spec, cols, points):
    torch.manual_seed(cols + points)
    B, H, Sq = 2, 3, 40
    scores = (torch.randn(B, H, Sq, cols, device=DEV) * 6).bfloat16()
    mask = torch.zeros(B, 1, Sq, cols, device=DEV, dtype=torch.bfloat16)
    mask[..., cols // 2:] = torch.finfo(torch.bfloat16).min
    mask[0, 0, 0] = 0
    alpha = 0.125
    fq = table_fq(oracle, spec)
    fmt, lut = fmt_lut(spec)
    probs = torch.empty_like(scores)
    _C.softmax_fq(scores, probs, alpha, mask.reshape(B, Sq, cols), H * Sq, Sq, B, points, fmt, lut=lut)
    s = fq(scores) if points & _C.FQ_PRE else scores
    s = s * alpha
    s = s + mask
    if points & _C.FQ_MID:
        s = fq(s)
    p = torch.softmax(s, dim=-1)
    if points & _C.FQ_POST:
        p = fq(p)
    assert_close_codes(probs, p)
    # batch-broadcast mask and no mask / no scaling
    _C.softmax_fq(scores, probs, 1.0, None, H * Sq, Sq, 1, _C.FQ_POST, fmt, lut=lut)
    assert_close_codes(probs, fq(torch.softmax(scores, dim=-1)))
    _C.softmax_fq(scores, probs, alpha, mask[:1].reshape(1, Sq, cols), H * Sq, Sq, 1, _C.FQ_POST, fmt, lut=lut)
    assert_close_codes(probs, fq(torch.softmax(scores * alpha + mask[:1], dim=-1)))


def test_softmax_pre_steps_are_exact(oracle):
    """Without the exp: with one column per row softmax is 1, so use the PRE / MID path through a trick -- a huge
    negative mask on all but one column makes every probability 0 or 1 exactly; the test is on those exact values."""
    fmt, lut = fmt_lut("posit8_1")
    scores = (torch.randn(64, 256, device=DEV) * 3).bfloat16()
    mask = torch.full((1, 64, 256), torch.finfo(torch.bfloat16).min, device=DEV, dtype=torch.bfloat16)
    mask[0, torch.arange(64), torch.arange(64)] = 0
    probs = torch.empty_like(scores)
    _C.softmax_fq(scores, probs, 0.5, mask, 64, 64, 1, 7, fmt, lut=lut)
    want = torch.zeros_like(scores)
    want[torch.arange(64), torch.arange(64)] = 1
    assert torch.equal(probs, want)


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "bfloat16"])
@pytest.mark.parametrize("cols", [4096, 768, 128, 512, 2048, 8192, 200])
def test_rmsnorm_and_layernorm(oracle, spec, cols):
    torch.manual_seed(cols)
    rows = 37
    x = (torch.randn(rows, cols, device=DEV) * 2).bfloat16()
    w = (1 + 0.1 * torch.randn(cols, device=DEV)).bfloat16()
    b = (0.1 * torch.randn(cols, device=DEV)).bfloat16()
    fq = table_fq(oracle, spec) if spec != "bfloat16" else (lambda t: t)
    fmt, lut = fmt_lut(spec)
    y = torch.empty_like(x)
    # LlamaRMSNorm, HF modeling_llama.py: fp32 inside, cast, then weight *
    _C.norm_fq(x, y, _C.NORM_RMS, w, None, 1e-5, _C.FQ_POST, fmt, lut=lut)
    h = x.float()
    h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + 1e-5)
    assert_close_codes(y, fq(w * h.to(torch.bfloat16)))
    # nn.LayerNorm with an input fake-quant as well (the "layernorm" op group)
    _C.norm_fq(x, y, _C.NORM_LAYER, w, b, 1e-12, _C.FQ_PRE | _C.FQ_POST, fmt, lut=lut)
    assert_close_codes(y, fq(torch.nn.functional.layer_norm(fq(x), (cols,), w, b, 1e-12)))
    _C.norm_fq(x, y, _C.NORM_LAYER, w, None, 1e-5, 0, fmt, lut=lut)
    assert_close_codes(y, torch.nn.functional.layer_norm(x, (cols,), w, None, 1e-5))


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3", "int4"])
def test_silu_mul_and_gelu(oracle, spec):
    torch.manual_seed(3)
    rows, inter = 50, 1376
    gu = (torch.randn(rows, 2 * inter, device=DEV) * 2).bfloat16()
    gate, up = gu[:, :inter], gu[:, inter:]                    # two halves of a fused projection (strided rows)
    fq = table_fq(oracle, spec)
    fmt, lut = fmt_lut(spec)
    out = torch.empty(rows, inter, device=DEV, dtype=torch.bfloat16)
    _C.act_mul_fq(gate, up, out, "silu", _C.FQ_POST, fmt, lut=lut)
    assert_close_codes(out, fq(torch.nn.functional.silu(gate) * up), min_equal=0.995)
    _C.act_mul_fq(gate, None, out, "gelu", _C.FQ_POST, fmt, lut=lut)
    assert_close_codes(out, fq(torch.nn.functional.gelu(gate)), min_equal=0.995)
    _C.act_mul_fq(gate, None, out, "relu", _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(out), bits(fq(torch.relu(gate))))
    _C.act_mul_fq(gate, up, out, None, 0, fmt, lut=lut)
    assert torch.equal(bits(out), bits(gate * up))


@pytest.mark.parametrize("spec", ["posit8_1", "e4m3"])
@pytest.mark.parametrize("heads,kv_heads,d", [(32, 32, 128), (12, 4, 64)])
def test_rope_is_bit_exact(oracle, spec, heads, kv_heads, d):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    torch.manual_seed(heads)
    B, S = 2, 33
    qkv = (torch.randn(B * S, (heads + 2 * kv_heads) * d, device=DEV) * 2).bfloat16()   # fused QKV buffer
    q = qkv[:, :heads * d].view(B * S, heads, d)
    k = qkv[:, heads * d:(heads + kv_heads) * d].view(B * S, kv_heads, d)
    pos = torch.arange(S, device=DEV, dtype=torch.float32)
    inv = 1.0 / (10000 ** (torch.arange(0, d, 2, device=DEV, dtype=torch.float32) / d))
    fr = torch.outer(pos, inv)
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos().bfloat16(), emb.sin().bfloat16()                                # [S, d], shared by the batch
    fq = table_fq(oracle, spec)
    fmt, lut = fmt_lut(spec)
    qo = torch.empty(B * S, heads, d, device=DEV, dtype=torch.bfloat16)
    ko = torch.empty(B * S, kv_heads, d, device=DEV, dtype=torch.bfloat16)
    _C.rope_fq(q, qo, k, ko, cos, sin, _C.FQ_POST, fmt, lut=lut)
    q4 = q.reshape(B, S, heads, d).transpose(1, 2)
    k4 = k.reshape(B, S, kv_heads, d).transpose(1, 2)
    qr, kr = apply_rotary_pos_emb(q4, k4, cos[None], sin[None])
    assert torch.equal(bits(qo.view(B, S, heads, d).transpose(1, 2)), bits(fq(qr)))
    assert torch.equal(bits(ko.view(B, S, kv_heads, d).transpose(1, 2)), bits(fq(kr)))


@pytest.mark.parametrize("shape", [(1, 1024, 32, 128), (3, 100, 4, 64), (2, 77, 2, 32)])
def test_fq_transpose_is_bit_exact(oracle, shape):
    B, S, H, D = shape
    torch.manual_seed(S)
    qkv = (torch.randn(B, S, 3 * H * D, device=DEV) * 2).bfloat16()
    v = qkv[..., 2 * H * D:].view(B, S, H, D)
    fq = table_fq(oracle, "posit8_1")
    fmt, lut = fmt_lut("posit8_1")
    out = torch.empty(B, H, D, S, device=DEV, dtype=torch.bfloat16)
    _C.fq_transpose(v, out, _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(out), bits(fq(v).permute(0, 2, 3, 1)))


def test_fused_ops_reject_what_they_cannot_do():
    fmt, lut = fmt_lut("posit8_1")
    x = torch.zeros(4, 8200, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        _C.softmax_fq(x, torch.empty_like(x), 1.0, None, 4, 4, 1, 4, fmt, lut=lut)     # row too long
    with pytest.raises(ValueError):
        _C.norm_fq(x[:, :64], x[:, :64], 0, x[0, :64], None, 1e-5, 4, fmt, lut=None) if False else \
            _C.norm_fq(x[:, :64].contiguous(), torch.empty(4, 64, device=DEV, dtype=torch.bfloat16), 0,
                       torch.ones(64, device=DEV, dtype=torch.bfloat16), None, 1e-5, 4, fmt, lut=None)  # no table
    with pytest.raises(TypeError):
        _C.norm_fq(x.float(), x.float(), 0, x[0], None, 1e-5, 4, fmt, lut=lut)


@pytest.mark.parametrize("spec,tdt", [("e4m3", torch.float8_e4m3fn), ("e5m2", torch.float8_e5m2), ("fp8_e4m3", torch.float8_e4m3fn)])
def test_fp8_code_outputs_decode_to_the_bf16_outputs(spec, tdt):
    """uint8 destinations receive the fp8 codes of exactly the values the bf16 destinations receive."""
    torch.manual_seed(8)
    fmt, lut = fmt_lut(spec)
    dec = lambda c: c.view(tdt).to(torch.bfloat16)
    x = (torch.randn(40, 1024, device=DEV) * 3).bfloat16()
    x[0, :4] = torch.tensor([float("inf"), -float("inf"), 1e30, -0.0], device=DEV)
    w = (1 + 0.1 * torch.randn(1024, device=DEV)).bfloat16()
    yb, yc = torch.empty_like(x), torch.empty(x.shape, dtype=torch.uint8, device=DEV)
    _C.norm_fq(x[1:], yb[1:], _C.NORM_RMS, w, None, 1e-5, _C.FQ_POST, fmt, lut=lut)
    _C.norm_fq(x[1:], yc[1:], _C.NORM_RMS, w, None, 1e-5, _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(dec(yc[1:])), bits(yb[1:]))
    _C.act_mul_fq(x, None, yb, None, _C.FQ_POST, fmt, lut=lut)          # plain fake quant incl. Inf / huge / -0
    _C.act_mul_fq(x, None, yc, None, _C.FQ_POST, fmt, lut=lut)
    got, want = dec(yc).float(), yb.float()
    same = (got == want) | (torch.isnan(got) & torch.isnan(want)) | (torch.isnan(got) & torch.isinf(want))  # e4m3 has no Inf
    assert bool(same.all())
    _C.act_mul_fq(x, x, yb, "silu", _C.FQ_POST, fmt, lut=lut)
    _C.act_mul_fq(x, x, yc, "silu", _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(dec(yc[1:])), bits(yb[1:]))
    s = (torch.randn(2, 2, 16, 1024, device=DEV) * 4).bfloat16()
    pb, pc = torch.empty_like(s), torch.empty(s.shape, dtype=torch.uint8, device=DEV)
    _C.softmax_fq(s, pb, 0.5, None, 32, 16, 1, _C.FQ_POST, fmt, lut=lut)
    _C.softmax_fq(s, pc, 0.5, None, 32, 16, 1, _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(dec(pc)), bits(pb))
    v = (torch.randn(2, 100, 4, 64, device=DEV) * 2).bfloat16()
    tb, tc = torch.empty(2, 4, 64, 100, device=DEV, dtype=torch.bfloat16), torch.empty(2, 4, 64, 100, device=DEV, dtype=torch.uint8)
    _C.fq_transpose(v, tb, _C.FQ_POST, fmt, lut=lut)
    _C.fq_transpose(v, tc, _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(dec(tc)), bits(tb))
    q = (torch.randn(50, 4, 64, device=DEV)).bfloat16()
    cos = torch.rand(50, 64, device=DEV).bfloat16(); sin = torch.rand(50, 64, device=DEV).bfloat16()
    rb, rc = torch.empty_like(q), torch.empty(q.shape, dtype=torch.uint8, device=DEV)
    _C.rope_fq(q, rb, None, None, cos, sin, _C.FQ_POST, fmt, lut=lut)
    _C.rope_fq(q, rc, None, None, cos, sin, _C.FQ_POST, fmt, lut=lut)
    assert torch.equal(bits(dec(rc)), bits(rb))
    with pytest.raises(ValueError):       # codes need the output fake-quant step ...
        _C.norm_fq(x, yc, _C.NORM_RMS, w, None, 1e-5, 0, fmt, lut=lut)
    pfmt, plut = fmt_lut("posit8_1")
    with pytest.raises(ValueError):       # ... of an fp8 format
        _C.norm_fq(x, yc, _C.NORM_RMS, w, None, 1e-5, _C.FQ_POST, pfmt, lut=plut)


@pytest.mark.parametrize("spec,codes", [("posit8_1", False), ("e4m3", False), ("e4m3", True)])
@pytest.mark.parametrize("B,H,S,D", [(1, 4, 1024, 128), (2, 3, 384, 64), (1, 2, 128, 128), (1, 1, 2048, 64)])
def test_causal_schedule_is_bit_identical(spec, codes, B, H, S, D):
    """The causal schedule of the three-kernel chain (QT_CAUSAL_OUT_LOWER scores, QT_SOFTMAX_CAUSAL,
    QT_CAUSAL_A_LOWER context) skips work whose contribution is exactly zero: the context equals the full
    computation bit for bit, computed score tiles are identical, and so are the probabilities that get written."""
    if codes and D != 128:
        pytest.skip("fp8 q / k rows are kept a multiple of 128 bytes")
    torch.manual_seed(S * 7 + D)
    m = qt.FusedAmaxObsFakeQuantize(spec, device=DEV)
    fmt, lut = m._fmt, m.lut
    qkv = m((torch.randn(B, S, 3 * H * D, device=DEV) * 1.5).bfloat16())
    vt = torch.empty(B, H, D, S, device=DEV, dtype=torch.bfloat16)
    _C.fq_transpose(qkv[..., 2 * H * D:].view(B, S, H, D), vt, 0, fmt, lut=lut)
    t_op = _C.GEMM_BF16
    if codes:
        tdt = torch.float8_e4m3fn
        enc = lambda t: t.contiguous().to(tdt).view(torch.uint8)
        qc = enc(qkv[..., :2 * H * D])
        q = qc[..., :H * D].view(B, S, H, D).transpose(1, 2)
        k = qc[..., H * D:].view(B, S, H, D).transpose(1, 2)
        vt = enc(vt)
        t_op = _C.GEMM_E4M3
    else:
        q = qkv[..., :H * D].view(B, S, H, D).transpose(1, 2)
        k = qkv[..., H * D:2 * H * D].view(B, S, H, D).transpose(1, 2)
    mask = torch.full((S, S), torch.finfo(torch.bfloat16).min, device=DEV, dtype=torch.bfloat16).triu(1)[None].contiguous()
    alpha = D ** -0.5

    def chain(causal, mask=mask, flag=None):
        scores = torch.full((B, H, S, S), float("nan"), device=DEV, dtype=torch.bfloat16)  # poison what is skipped
        _C.gemm_nt(q, k, out=scores, operand_type=t_op, causal=_C.CAUSAL_OUT_LOWER if causal else 0, causal_flag=flag)
        probs = torch.full((B, H, S, S), 77 if codes else float("nan"), device=DEV,
                           dtype=torch.uint8 if codes else torch.bfloat16)
        _C.softmax_fq(scores, probs, alpha, mask, H * S, S, 1, _C.FQ_POST | (_C.SOFTMAX_CAUSAL if causal else 0),
                      fmt, lut=lut, causal_flag=flag)
        ctx = torch.empty(B, S, H * D, device=DEV, dtype=torch.bfloat16)
        _C.gemm_nt(probs, vt, out=ctx.view(B, S, H, D).transpose(1, 2), operand_type=t_op,
                   causal=_C.CAUSAL_A_LOWER if causal else 0, causal_flag=flag)
        return scores, probs, ctx

    s0, p0, c0 = chain(False)
    flag = _C.causal_mask_check(mask)
    assert int(flag.item()) == 1
    s1, p1, c1 = chain(True, flag=flag)
    # a mask that is NOT the causal one switches the schedule off on the device: same results as the full computation
    other = mask.clone()
    other[0, S // 2, 3] = torch.finfo(torch.bfloat16).min
    flag0 = _C.causal_mask_check(other)
    assert int(flag0.item()) == 0
    sa, pa, ca = chain(False, mask=other)
    sb, pb_, cb = chain(True, mask=other, flag=flag0)
    assert torch.equal(bits(cb), bits(ca)) and torch.equal(bits(sb), bits(sa))
    assert torch.equal(pb_ if codes else bits(pb_), pa if codes else bits(pa))
    assert torch.equal(bits(c1), bits(c0))
    assert not bool(torch.isnan(c1.float()).any())
    rows = torch.arange(S, device=DEV)[:, None]
    cols = torch.arange(S, device=DEV)[None, :]
    lower = cols <= rows
    assert torch.equal(bits(s1)[..., lower], bits(s0)[..., lower])
    written = cols < ((rows // 128) + 1) * 128
    pb = (lambda t: t) if codes else bits
    assert torch.equal(pb(p1)[..., written], pb(p0)[..., written])
    if S > 128:  # something was actually skipped
        assert bool((pb(p1)[..., ~written] != pb(p0)[..., ~written]).any())


def test_causal_flags_are_validated():
    a = torch.zeros(2, 128, 256, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(2, 256, 256, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="causal"):
        _C.gemm_nt(a, b, causal=_C.CAUSAL_OUT_LOWER)          # M = 128, N = 256: not a square score matrix
    with pytest.raises(ValueError, match="causal"):
        _C.gemm_nt(a, b, causal=_C.CAUSAL_A_LOWER)            # A is 128 x 256: not a square probability matrix
    m = qt.FusedAmaxObsFakeQuantize("e4m3", device=DEV)
    sc = torch.zeros(1, 1, 128, 128, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="CAUSAL"):
        _C.softmax_fq(sc, torch.empty_like(sc), 1.0, None, 128, 128, 1, _C.FQ_POST | _C.SOFTMAX_CAUSAL, m._fmt, lut=m.lut)
