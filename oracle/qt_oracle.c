/*
 * qt_oracle.c -- CPU ORACLE (test infrastructure only; see qt_oracle.h).
 *
 * Restates, function by function, the reference's fake-quant path:
 *   src/quantized_training/fake_quantize.py   get_quantization_map :31-95,
 *                                             FusedAmaxObsFakeQuantFunction.forward :202-248
 *   src/quantized_training/fp8.py             quantize_to_fp8_e4m3/_e5m2 :10-67,
 *                                             _round_mantissa :104-135, _quantize_elemwise_core :147-203
 *   src/quantized_training/posit.py           quantize_to_posit :6-67
 *   src/quantized_training/decomposed.py      vmap :146-163
 * The reference evaluates these with torch CPU tensor ops; every place where
 * torch's dtype rules matter (bf16 rounding after each op of the "fpN_eXmY"
 * path, out-of-range tensor shifts yielding 0, clamp bounds cast to bf16) is
 * reproduced and commented.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).  No dependency on the
 * product sources.
 */
#include "qt_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ helpers */

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float bf2f(uint16_t h) { return u2f((uint32_t)h << 16); }

/* c10::BFloat16 round_to_nearest_even: NaN -> 0x7FC0 */
static inline uint16_t f2bf(float f)
{
    uint32_t u = f2u(f);
    if (isnan(f)) return 0x7FC0;
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
/* round a float through bf16 (what every bf16 tensor op does to its fp32 result) */
static inline float bfr(float f) { return bf2f(f2bf(f)); }

/* torch tensor shifts on int32: ATen lshift returns 0 for counts <0 or >=32,
 * rshift sign-fills for counts <0 or >=31 (BinaryOpsKernel.cpp). */
static inline int32_t shl32(int32_t a, int32_t b)
{
    if (b < 0 || b >= 32) return 0;
    return (int32_t)((uint32_t)a << b);
}
static inline int32_t shr32(int32_t a, int32_t b)
{
    if (b < 0 || b >= 31) return a >> 31;
    return a >> b;
}
static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* torch.clamp(t, lo, hi) on floats: NaN propagates; std::min(std::max(a, lo), hi) */
static inline float clampf_t(float a, float lo, float hi)
{
    if (isnan(a)) return a;
    float m = (a < lo) ? lo : a;
    return (hi < m) ? hi : m;
}
/* torch.sign */
static inline float signf_t(float a) { return (float)((0.0f < a) - (a < 0.0f)); }

/* ------------------------------------------------- int / uint (fake_quantize.py:43-52) */
/* torch.clamp(torch.round(values), quant_min, quant_max) evaluated on a bf16 tensor:
 * round half-even, bounds cast to bf16. */
static uint16_t q_int(uint16_t idx, double qmin, double qmax)
{
    float v = bf2f(idx);
    float r = bfr(nearbyintf(v)); /* torch.round = rint, half to even */
    float lo = bfr((float)qmin), hi = bfr((float)qmax);
    return f2bf(clampf_t(r, lo, hi));
}

/* ------------------------------------------------- e4m3 / e5m2 (fp8.py:10-67) */
static uint16_t q_fp8_custom(uint16_t idx, int mbits, float fp8_max, double fp8_min)
{
    float input = bf2f(idx);
    int32_t raw_bits = (int32_t)f2u(input);
    int32_t exp = ((raw_bits & 0x7f800000) >> 23) - 127;
    int32_t fraction = (raw_bits & 0x7fffff) | 0x800000;

    int32_t min_exp = (int32_t)floor(log2(fp8_min));
    int32_t nf_mask = 23 - mbits + (min_exp - exp > 0 ? min_exp - exp : 0);
    int lb = (fraction & shl32(1, nf_mask)) != 0;
    int gb = (fraction & shl32(1, nf_mask - 1)) != 0;
    int sb = (fraction & (shl32(1, nf_mask - 1) - 1)) != 0;
    int rb = (lb & gb) | (gb & sb);

    int32_t nf_mask_clamped = nf_mask > 23 ? 23 : nf_mask;
    raw_bits &= shl32(-1, nf_mask_clamped);
    if (rb) raw_bits = (int32_t)((uint32_t)raw_bits + (uint32_t)shl32(1, nf_mask_clamped));

    float output = u2f((uint32_t)raw_bits);
    output = clampf_t(output, -fp8_max, fp8_max);
    /* threshold is compared against a bf16 tensor: it is a power of two, exact in bf16 */
    float thr = bfr((float)(fp8_min * pow(2.0, -(mbits + 1))));
    if (fabsf(input) <= thr) output = 0.0f;
    if (input == 0.0f) output = 0.0f;
    if (!isfinite(input)) output = NAN;
    return f2bf(output);
}

/* ------------------------------------------------- fpN_eXmY (fake_quantize.py:63-80 -> fp8.py:147-203) */
/* _round_mantissa(A, bits, "even") on a bf16 tensor (fp8.py:123-127) */
static float mx_round_even_bf16(float A)
{
    float absA = fabsf(A);
    /* (absA - 0.5) % 2 == 0 : torch.remainder in fp32 opmath, result cast to bf16 */
    float d = bfr(absA - 0.5f);
    float mod = fmodf(d, 2.0f);
    if (mod != 0.0f && (mod < 0.0f)) mod += 2.0f;
    mod = bfr(mod);
    float maskA = (mod == 0.0f) ? 1.0f : 0.0f;
    float fl = floorf(bfr(absA + 0.5f)); /* absA + 0.5 is rounded to bf16 BEFORE the floor */
    float v = bfr(fl - maskA);
    return bfr(signf_t(A) * v);
}

static uint16_t q_mx(uint16_t idx, int is_unsigned, int bits /* mbits+2 */, int exp_bits, float max_norm)
{
    float A = bf2f(idx);
    if (is_unsigned) A = fabsf(A);
    float out = A;
    float private_exp = 0.0f;
    if (exp_bits != 0) {
        /* floor(log2(abs(A) + (A == 0))) in bf16 arithmetic */
        float t = bfr(fabsf(A) + ((A == 0.0f) ? 1.0f : 0.0f));
        private_exp = floorf(bfr(log2f(t)));
        float min_exp = (float)(-(1 << (exp_bits - 1)) + 2);
        if (private_exp < min_exp) private_exp = min_exp; /* clip(min=); NaN stays NaN */
    }
    float p2e = bfr(exp2f(private_exp));      /* 2 ** private_exp, bf16; 2**128 -> inf */
    float p2b = (float)(1 << (bits - 2));     /* python int */
    /* _safe_lshift: x / 2**exp * 2**bits */
    out = bfr(out / p2e);
    out = bfr(out * p2b);
    out = mx_round_even_bf16(out);
    /* _safe_rshift: x / 2**bits * 2**exp */
    out = bfr(out / p2b);
    out = bfr(out * p2e);
    /* saturate_normals=True: clamp, bounds cast to bf16 */
    float mn = bfr(max_norm);
    out = clampf_t(out, -mn, mn);
    if (A == INFINITY) out = INFINITY;
    if (A == -INFINITY) out = -INFINITY;
    return f2bf(out);
}

/* ------------------------------------------------- posit (posit.py:6-67) */
static void q_posit(uint16_t idx, int nbits, int es, uint16_t *val, int32_t *pbits_out)
{
    float input = bf2f(idx);
    int32_t raw_bits = (int32_t)f2u(input);
    int32_t scale = ((raw_bits & 0x7f800000) >> 23) - 127;
    int32_t fraction = raw_bits & 0x7fffff;
    int r = scale >= 0;

    int32_t max_scale = (nbits - 2) * (1 << es);
    int regime_dominated = r ? (scale > max_scale) : (scale < -max_scale);

    int32_t run = r ? 1 + (scale >> es) : -(scale >> es);
    int32_t regime = (r ? (shl32(1, run + 1) - 1) : 0) ^ 1;
    int32_t exponent = ((scale % (1 << es)) + (1 << es)) % (1 << es); /* python-style % */
    int32_t pt_bits = shl32(regime, 23 + es) | shl32(exponent, 23) | fraction;

    int32_t len = 2 + run + es + 23;
    int32_t lb_mask = shl32(1, len - nbits);
    int32_t gb_mask = shr32(lb_mask, 1);
    int32_t sb_mask = gb_mask - 1;

    int lb = (pt_bits & lb_mask) != 0;
    int gb = (pt_bits & gb_mask) != 0;
    int sb = (pt_bits & sb_mask) != 0;
    int rb = ((lb & gb) | (gb & sb)) & !regime_dominated;

    /* truncate exponent bits */
    int32_t ne_mask = clampi(2 + run + es - nbits, 0, es);
    scale &= shl32(-1, ne_mask);
    scale = clampi(scale, -max_scale, max_scale);

    /* truncate fraction bits */
    int32_t nf_mask = clampi(len - nbits, 0, 23);
    fraction &= shl32(-1, nf_mask);

    int32_t output = shl32(scale + 127, 23) | fraction;
    if (rb) output = (int32_t)((uint32_t)output + (uint32_t)shl32(1, nf_mask + ne_mask));
    float outf = u2f((uint32_t)output) * signf_t(input);

    /* round_to_even: flush below 2^floor(-(nbits-1)*2^es + 2^(es-1)); threshold cast to bf16 */
    double thr_d = pow(2.0, floor(-(double)(nbits - 1) * (double)(1 << es) + pow(2.0, es - 1)));
    float thr = bfr((float)thr_d);
    if (fabsf(input) < thr) outf = 0.0f;

    if (input == 0.0f) outf = 0.0f;
    if (!isfinite(input)) outf = NAN;
    *val = f2bf(outf);

    if (pbits_out) {
        int32_t pb = shr32(pt_bits, len - nbits);
        pb &= (1 << (nbits - 1)) - 1;
        if (rb) pb += 1;
        pb *= (int32_t)signf_t(input);
        *pbits_out = pb;
    }
}

/* ------------------------------------------------- dtype-string dispatch (fake_quantize.py:31-95) */
static int parse_uint(const char **p, int *out)
{
    if (!isdigit((unsigned char)**p)) return 0;
    long v = 0;
    while (isdigit((unsigned char)**p)) { v = v * 10 + (**p - '0'); (*p)++; if (v > 100000) return 0; }
    *out = (int)v;
    return 1;
}
static int ci_prefix(const char *s, const char *pre) /* case-insensitive prefix */
{
    while (*pre) { if (tolower((unsigned char)*s) != *pre) return 0; s++; pre++; }
    return 1;
}

int qto_posit(int nbits, int es, uint16_t *qmap, int32_t *pbits)
{
    if (nbits < 2 || nbits > 24 || es < 0 || es > 4) return 1;
    for (uint32_t i = 0; i < 65536; i++) q_posit((uint16_t)i, nbits, es, &qmap[i], pbits ? &pbits[i] : NULL);
    return 0;
}

int qto_qmap(const char *dtype, uint16_t *qmap)
{
    const char *p;
    int n, e, m;
    if (dtype == NULL) return 1;

    if (!strcmp(dtype, "float32") || !strcmp(dtype, "bfloat16")) {
        for (uint32_t i = 0; i < 65536; i++) qmap[i] = (uint16_t)i;
        return 0;
    }
    /* int(\d+), IGNORECASE */
    p = dtype;
    if (ci_prefix(p, "int")) {
        p += 3;
        if (parse_uint(&p, &n) && *p == 0 && n >= 1 && n <= 24) {
            double qmin = -ldexp(1.0, n - 1), qmax = ldexp(1.0, n - 1) - 1;
            for (uint32_t i = 0; i < 65536; i++) qmap[i] = q_int((uint16_t)i, qmin, qmax);
            return 0;
        }
    }
    /* uint(\d+), IGNORECASE */
    p = dtype;
    if (ci_prefix(p, "uint")) {
        p += 4;
        if (parse_uint(&p, &n) && *p == 0 && n >= 1 && n <= 24) {
            double qmax = ldexp(1.0, n) - 1;
            for (uint32_t i = 0; i < 65536; i++) qmap[i] = q_int((uint16_t)i, 0.0, qmax);
            return 0;
        }
    }
    /* (?:fp8\.)?(e4m3|e5m2), IGNORECASE */
    p = dtype;
    if (ci_prefix(p, "fp8.")) p += 4;
    if (ci_prefix(p, "e4m3") && p[4] == 0) {
        for (uint32_t i = 0; i < 65536; i++) qmap[i] = q_fp8_custom((uint16_t)i, 3, 448.0f, ldexp(1.0, -6));
        return 0;
    }
    if (ci_prefix(p, "e5m2") && p[4] == 0) {
        for (uint32_t i = 0; i < 65536; i++) qmap[i] = q_fp8_custom((uint16_t)i, 2, 57344.0f, ldexp(1.0, -14));
        return 0;
    }
    /* fp(\d+)_e(\d+)m(\d+), case-sensitive */
    p = dtype;
    if (!strncmp(p, "fp", 2)) {
        p += 2;
        if (parse_uint(&p, &n) && *p == '_' && p[1] == 'e') {
            p += 2;
            if (parse_uint(&p, &e) && *p == 'm') {
                p += 1;
                if (parse_uint(&p, &m) && *p == 0) {
                    if (!(n == e + m + 1 || n == e + m)) return 1; /* reference: assert */
                    if (e < 1 || e > 7 || m < 0 || m > 8) return 1;
                    int is_unsigned = (n == e + m);
                    int bits = m + 2;
                    int emax = e > 4 ? (1 << (e - 1)) - 1 : (1 << (e - 1));
                    double max_norm;
                    if (strcmp(dtype, "fp8_e4m3") != 0)
                        max_norm = ldexp(1.0, emax) * (double)((1 << (bits - 1)) - 1) / ldexp(1.0, bits - 2);
                    else
                        max_norm = ldexp(1.0, emax) * 1.75;
                    for (uint32_t i = 0; i < 65536; i++)
                        qmap[i] = q_mx((uint16_t)i, is_unsigned, bits, e, (float)max_norm);
                    return 0;
                }
            }
        }
    }
    /* posit(\d+)_(\d+), case-sensitive */
    p = dtype;
    if (!strncmp(p, "posit", 5)) {
        p += 5;
        if (parse_uint(&p, &n) && *p == '_') {
            p += 1;
            if (parse_uint(&p, &e) && *p == 0) return qto_posit(n, e, qmap, NULL);
        }
    }
    return 1; /* ValueError("Unsupported dtype") -- nf* is outside the hot-path scope */
}

/* ------------------------------------------------- vmap (decomposed.py:146-163) */
void qto_vmap_bf16(const uint16_t *x, uint16_t *y, size_t n, const uint16_t *qmap)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) y[i] = qmap[x[i]];
}

static inline uint32_t rto_index(float v) /* fp32 -> bf16 index, round-to-odd (:151-153) */
{
    uint32_t b = f2u(v);
    return ((b >> 16) & 0xffffu) | ((b & 0xffffu) != 0);
}

void qto_vmap_f32(const float *x, float *y, size_t n, const uint16_t *qmap)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) y[i] = bf2f(qmap[rto_index(x[i])]);
}

/* ------------------------------------------------- observer (fake_quantize.py:217-242) */
/* |v| as ordered bits: NaN > Inf > finite, so an unsigned max propagates NaN like torch.amax */
static inline uint32_t absbits_bf16(uint16_t h) { return ((uint32_t)h & 0x7fffu) << 16; }
static inline uint32_t absbits_f32(float v) { return f2u(v) & 0x7fffffffu; }

void qto_amax(const void *x, int is_f32, size_t outer, size_t C, size_t inner, float *amax_cur)
{
    const uint16_t *xb = (const uint16_t *)x;
    const float *xf = (const float *)x;
    if (C == 1) {
        size_t n = outer * inner;
        uint32_t m = 0;
#pragma omp parallel for reduction(max : m) schedule(static)
        for (size_t i = 0; i < n; i++) {
            uint32_t b = is_f32 ? absbits_f32(xf[i]) : absbits_bf16(xb[i]);
            if (b > m) m = b;
        }
        amax_cur[0] = u2f(m);
        return;
    }
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < C; c++) {
        uint32_t m = 0;
        for (size_t o = 0; o < outer; o++) {
            size_t base = (o * C + c) * inner;
            for (size_t i = 0; i < inner; i++) {
                uint32_t b = is_f32 ? absbits_f32(xf[base + i]) : absbits_bf16(xb[base + i]);
                if (b > m) m = b;
            }
        }
        amax_cur[c] = u2f(m);
    }
}

void qto_scale_update(float *history, int ahl, size_t C, const float *amax_cur,
                      float *scale, float quant_max, int force_pow2)
{
    for (size_t c = 0; c < C; c++) {
        /* amax = torch.amax(amax_history, dim=0), read BEFORE the insert (:230); NaN propagates */
        float amax = history[c];
        for (int i = 1; i < ahl; i++) {
            float h = history[(size_t)i * C + c];
            if (isnan(amax)) break;
            if (isnan(h) || h > amax) amax = h;
        }
        /* roll(-1, 0) then slot 0 <- current (:232-235) */
        if (ahl > 1) {
            float first = history[c];
            for (int i = 0; i < ahl - 1; i++) history[(size_t)i * C + c] = history[(size_t)(i + 1) * C + c];
            history[(size_t)(ahl - 1) * C + c] = first;
        }
        history[c] = amax_cur[c];
        /* sf = amax / quant_max, kept only if amax > 0 and finite (:237-239) */
        float sf = amax / quant_max;
        if (!(amax > 0.0f)) sf = scale[c];
        if (!isfinite(amax)) sf = scale[c];
        if (force_pow2) sf = powf(2.0f, ceilf(log2f(sf))); /* :240-241, fp32 */
        scale[c] = sf;
    }
}

/* ------------------------------------------------- fake quant (fake_quantize.py:244-246) */
void qto_fake_quant_bf16(const uint16_t *x, uint16_t *y, size_t outer, size_t C, size_t inner,
                         const float *scale, const uint16_t *qmap)
{
    if (C == 1) { /* per tensor / bare spec: one scale, parallel over elements */
        const float s = bfr(scale[0]); /* scale.to(bf16) */
        const size_t n = outer * inner;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            uint16_t u = f2bf(bf2f(x[i]) / s); /* input / scale, bf16 */
            float q = bf2f(qmap[u]);           /* vmap */
            y[i] = f2bf(q * s);                /* * scale, bf16 */
        }
        return;
    }
    const size_t n = outer * C * inner;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const float s = bfr(scale[(i / inner) % C]);
        uint16_t u = f2bf(bf2f(x[i]) / s);
        float q = bf2f(qmap[u]);
        y[i] = f2bf(q * s);
    }
}

void qto_fake_quant_f32(const float *x, float *y, size_t outer, size_t C, size_t inner,
                        const float *scale, const uint16_t *qmap)
{
    const size_t n = outer * C * inner;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const float s = scale[C == 1 ? 0 : (i / inner) % C];
        float u = x[i] / s;
        float q = bf2f(qmap[rto_index(u)]);
        y[i] = q * s;
    }
}

/* ------------------------------------------------- block-scaled qschemes
 * microscaling:      MXFakeQuantFunction.forward fake_quantize.py:105-129,
 *                    calculate_mx_qparam decomposed.py:372-419, quantize :171-210, expand :127-140,
 *                    _reshape_to_blocks mx_utils.py:62-121, _shared_exponents :16-59
 * group_wise_affine: GroupWiseAffineFakeQuantFunction.forward fake_quantize.py:138-190
 *
 * The tensor is contiguous with `ndim` dims; axes flagged in is_block[] are tiled with
 * `bs` (the reference pads with zeros to a multiple of bs, which changes nothing for
 * amax but DOES enter the min / max of the affine scheme).  Parameters come out with
 * shape[d] -> ceil(shape[d] / bs) on the block axes, as the reference's `scale` buffer.
 * Every op of the reference runs in the tensor's dtype: for bf16 tensors each
 * intermediate is rounded to bf16 (bfr), python scalars enter as fp32. */

#define QTO_MAXD 8
typedef struct {
    int ndim;
    size_t shape[QTO_MAXD], bshape[QTO_MAXD], bstride[QTO_MAXD];
    int blocked[QTO_MAXD];
    size_t n, nblocks;
    int bs, padded;
} blk_t;

static void blk_init(blk_t *B, int ndim, const size_t *shape, const int *is_block, int bs)
{
    B->ndim = ndim; B->bs = bs; B->n = 1; B->nblocks = 1; B->padded = 0;
    for (int d = 0; d < ndim; d++) {
        B->shape[d] = shape[d];
        B->blocked[d] = is_block[d] != 0;
        B->bshape[d] = is_block[d] ? (shape[d] + (size_t)bs - 1) / (size_t)bs : shape[d];
        if (is_block[d] && shape[d] % (size_t)bs) B->padded = 1;
        B->n *= shape[d];
    }
    for (int d = ndim - 1; d >= 0; d--) { B->bstride[d] = B->nblocks; B->nblocks *= B->bshape[d]; }
}
static inline size_t blk_of(const blk_t *B, size_t i)
{
    size_t b = 0;
    for (int d = B->ndim - 1; d >= 0; d--) {
        size_t c = i % B->shape[d];
        i /= B->shape[d];
        b += (B->blocked[d] ? c / (size_t)B->bs : c) * B->bstride[d];
    }
    return b;
}
static inline float ld_elem(const void *x, int is_f32, size_t i)
{
    return is_f32 ? ((const float *)x)[i] : bf2f(((const uint16_t *)x)[i]);
}
/* an op result in the tensor's dtype */
static inline float rnd(int is_f32, float v) { return is_f32 ? v : bfr(v); }
/* vmap of one value of the tensor's dtype */
static inline float vmap1(int is_f32, float v, const uint16_t *qmap)
{
    return bf2f(qmap[is_f32 ? rto_index(v) : f2bf(v)]);
}

/* scale of one block from its amax (decomposed.py:391-419) */
static float mx_block_scale(int is_f32, float amax, float quant_max, int force_pow2, const uint16_t *scale_qmap)
{
    float s;
    if (force_pow2) {
        /* shared_exp = floor(log2(amax + FP32_MIN_NORMAL * (amax == 0))) - floor(log2(quant_max)); 2 ** shared_exp */
        float a = rnd(is_f32, amax + (amax == 0.0f ? 0x1p-126f : 0.0f));
        float e = floorf(rnd(is_f32, log2f(a)));
        e = rnd(is_f32, e - (float)floor(log2((double)quant_max)));
        s = rnd(is_f32, powf(2.0f, e));
    } else {
        s = rnd(is_f32, amax / quant_max);
        if (scale_qmap) s = vmap1(is_f32, s, scale_qmap);
    }
    return s > 0.0f ? s : 1.0f; /* torch.where(scale > 0.0, scale, 1.0): NaN -> 1 */
}

void qto_mx_fake_quant(const void *x, void *y, int is_f32, int ndim, const size_t *shape, const int *is_block,
                       int bs, float quant_max, int force_pow2, const uint16_t *scale_qmap,
                       const uint16_t *qmap, float *scale_out)
{
    blk_t B;
    blk_init(&B, ndim, shape, is_block, bs);
    uint32_t *am = (uint32_t *)calloc(B.nblocks ? B.nblocks : 1, sizeof(uint32_t));
    for (size_t i = 0; i < B.n; i++) { /* amax(|x|) per block; NaN patterns order above Inf */
        uint32_t a = f2u(ld_elem(x, is_f32, i)) & 0x7fffffffu;
        size_t b = blk_of(&B, i);
        if (a > am[b]) am[b] = a;
    }
    for (size_t b = 0; b < B.nblocks; b++)
        scale_out[b] = mx_block_scale(is_f32, u2f(am[b]), quant_max, force_pow2, scale_qmap);
    free(am);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < B.n; i++) {
        const float s = scale_out[blk_of(&B, i)];
        const float q = vmap1(is_f32, rnd(is_f32, ld_elem(x, is_f32, i) / s), qmap); /* quantize(): input / scale, vmap */
        const float r = q * s;                                                     /* input * expand(sf) */
        if (is_f32) ((float *)y)[i] = r; else ((uint16_t *)y)[i] = f2bf(r);
    }
}

/* torch.amin / amax: NaN propagates */
static inline float nan_min(float a, float b) { return (isnan(a) || isnan(b)) ? NAN : (b < a ? b : a); }
static inline float nan_max(float a, float b) { return (isnan(a) || isnan(b)) ? NAN : (b > a ? b : a); }

void qto_gwa_fake_quant(const void *x, void *y, int is_f32, int ndim, const size_t *shape, const int *is_block,
                        int bs, float quant_min, float quant_max, const uint16_t *scale_qmap,
                        float *scale_out, float *zp_out)
{
    blk_t B;
    blk_init(&B, ndim, shape, is_block, bs);
    const size_t nb = B.nblocks ? B.nblocks : 1;
    float *mn = (float *)malloc(nb * sizeof(float)), *mx = (float *)malloc(nb * sizeof(float));
    /* a block cut by the tensor edge holds padding zeros (F.pad, mx_utils.py:94-96) */
    for (size_t b = 0; b < nb; b++) mn[b] = mx[b] = NAN;
    unsigned char *seen = (unsigned char *)calloc(nb, 1);
    for (size_t i = 0; i < B.n; i++) {
        const float v = ld_elem(x, is_f32, i);
        const size_t b = blk_of(&B, i);
        if (!seen[b]) { mn[b] = mx[b] = v; seen[b] = 1; }
        else { mn[b] = nan_min(mn[b], v); mx[b] = nan_max(mx[b], v); }
    }
    if (B.padded) {
        for (size_t b = 0; b < B.nblocks; b++) { /* is block b cut on any block axis? */
            size_t r = b; int cut = 0;
            for (int d = B.ndim - 1; d >= 0; d--) {
                size_t c = r % B.bshape[d]; r /= B.bshape[d];
                if (B.blocked[d] && (c + 1) * (size_t)B.bs > B.shape[d]) cut = 1;
            }
            if (cut) { mn[b] = nan_min(mn[b], 0.0f); mx[b] = nan_max(mx[b], 0.0f); }
        }
    }
    const float range = quant_max - quant_min;
    for (size_t b = 0; b < B.nblocks; b++) {
        float sf = rnd(is_f32, rnd(is_f32, mx[b] - mn[b]) / range);
        sf = sf > 0.0f ? sf : 1.0f;
        float zp = rnd(is_f32, rnd(is_f32, -mn[b] / sf) + quant_min);
        if (scale_qmap) { sf = vmap1(is_f32, sf, scale_qmap); zp = vmap1(is_f32, zp, scale_qmap); }
        scale_out[b] = sf; zp_out[b] = zp;
    }
    free(mn); free(mx); free(seen);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < B.n; i++) {
        const size_t b = blk_of(&B, i);
        const float sf = scale_out[b], zp = zp_out[b];
        float q = rnd(is_f32, rnd(is_f32, ld_elem(x, is_f32, i) / sf) + zp);
        q = nearbyintf(q);                        /* torch.round: half to even */
        q = clampf_t(q, quant_min, quant_max);    /* torch.clamp */
        const float r = rnd(is_f32, rnd(is_f32, q - zp) * sf);
        if (is_f32) ((float *)y)[i] = r; else ((uint16_t *)y)[i] = f2bf(r);
    }
}

int qto_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void qto_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
