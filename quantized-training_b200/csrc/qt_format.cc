// qt_format.cc -- host side of the format layer: dtype-string parsing, derived kernel
// parameters, host evaluation of the rounding logic.  Pure C++ (no CUDA calls), so it
// works on a box without a GPU.
//
// Mirrors the dispatch of get_quantization_map (fake_quantize.py:31-95) and
// get_quant_min_max (quantizer/quantizer.py:53-94) of the reference.
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "qt_internal.h"

static thread_local char g_err[512] = "";

void qt_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *qt_last_error(void) { return g_err; }
extern "C" const char *qt_version(void) { return "qt_b200 0.1 (sm_100a; fake-quant hot path of quantized-training)"; }

static float bf16_round(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if (f != f) return f;
    u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
    memcpy(&f, &u, 4);
    return f;
}
static uint32_t fbits(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

static bool parse_uint(const char *&p, int &out)
{
    if (!isdigit((unsigned char)*p)) return false;
    long v = 0;
    while (isdigit((unsigned char)*p)) {
        v = v * 10 + (*p - '0');
        if (v > 1000000) return false;
        ++p;
    }
    out = (int)v;
    return true;
}
static bool ieq_prefix(const char *s, const char *pre)
{
    for (; *pre; ++s, ++pre)
        if (tolower((unsigned char)*s) != *pre) return false;
    return true;
}

struct Parsed {
    enum Family { NONE, NATIVE, INT, UINT, FP8C, FPMX, POSIT, NF } fam = NONE;
    int n = 0, e = 0, m = 0;
};

// `loose`: get_quant_min_max matches every family case-insensitively and has no bare e4m3/e5m2;
// get_quantization_map matches fpN_eXmY / positN_ES case-sensitively.
static Parsed parse_dtype(const char *s, bool loose)
{
    Parsed r;
    const char *p;
    if (!s) return r;
    if (!loose && (!strcmp(s, "float32") || !strcmp(s, "bfloat16"))) {
        r.fam = Parsed::NATIVE;
        return r;
    }
    p = s;
    if (ieq_prefix(p, "int") && (p += 3, parse_uint(p, r.n)) && *p == 0) {
        r.fam = Parsed::INT;
        return r;
    }
    p = s;
    if (ieq_prefix(p, "uint") && (p += 4, parse_uint(p, r.n)) && *p == 0) {
        r.fam = Parsed::UINT;
        return r;
    }
    if (!loose) {
        p = s;
        if (ieq_prefix(p, "fp8.")) p += 4;
        if ((ieq_prefix(p, "e4m3") || ieq_prefix(p, "e5m2")) && p[4] == 0) {
            r.fam = Parsed::FP8C;
            r.n = 8;
            r.e = p[1] - '0';
            r.m = p[3] - '0';
            return r;
        }
    }
    p = s;
    if ((loose ? ieq_prefix(p, "fp") : !strncmp(p, "fp", 2)) && (p += 2, parse_uint(p, r.n)) && *p == '_' &&
        (loose ? tolower((unsigned char)p[1]) == 'e' : p[1] == 'e') && (p += 2, parse_uint(p, r.e)) &&
        (loose ? tolower((unsigned char)*p) == 'm' : *p == 'm') && (p += 1, parse_uint(p, r.m)) && *p == 0) {
        r.fam = Parsed::FPMX;
        return r;
    }
    p = s;
    if ((loose ? ieq_prefix(p, "posit") : !strncmp(p, "posit", 5)) && (p += 5, parse_uint(p, r.n)) && *p == '_' &&
        (p += 1, parse_uint(p, r.e)) && *p == 0) {
        r.fam = Parsed::POSIT;
        return r;
    }
    p = s;
    if ((loose ? ieq_prefix(p, "nf") : !strncmp(p, "nf", 2)) && (p += 2, parse_uint(p, r.n))) {
        if (*p == 0 || (*p == '_' && (p += 1, parse_uint(p, r.e)) && *p == 0)) {
            r.fam = Parsed::NF;
            return r;
        }
    }
    r = Parsed();
    return r;
}

static bool qt_mx_supported(int e, int m) { return e >= 2 && e <= 5 && m >= 1 && m <= 5 && e + m <= 8; }

static double mx_max_norm(const char *dtype, int ebits, int mbits_explicit)
{
    int bits = mbits_explicit + 2;
    int emax = ebits > 4 ? (1 << (ebits - 1)) - 1 : (1 << (ebits - 1));
    bool is_e4m3 = (strlen(dtype) == 8) && ieq_prefix(dtype, "fp8_e4m3");
    if (is_e4m3) return ldexp(1.75, emax);
    return ldexp((double)((1 << (bits - 1)) - 1), emax - (bits - 2));
}

extern "C" int qt_format_from_string(const char *dtype, qt_format_t *fmt)
{
    if (!fmt) {
        qt_set_error("qt_format_from_string: fmt is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    memset(fmt, 0, sizeof(*fmt));
    Parsed r = parse_dtype(dtype, false);
    switch (r.fam) {
    case Parsed::NATIVE:
        fmt->kind = QT_KIND_IDENTITY;
        fmt->nbits = 16;
        fmt->max_value = INFINITY;
        fmt->min_value = -INFINITY;
        return QT_OK;
    case Parsed::INT:
    case Parsed::UINT:
        if (r.n < 1 || r.n > 24) break;
        fmt->kind = QT_KIND_INT;
        fmt->nbits = r.n;
        fmt->is_unsigned = r.fam == Parsed::UINT;
        fmt->max_value = (float)(r.fam == Parsed::UINT ? ldexp(1.0, r.n) - 1 : ldexp(1.0, r.n - 1) - 1);
        fmt->min_value = (float)(r.fam == Parsed::UINT ? 0.0 : -ldexp(1.0, r.n - 1));
        return QT_OK;
    case Parsed::FP8C:
        fmt->kind = QT_KIND_FP;
        fmt->flavour = QT_FP_CUSTOM;
        fmt->nbits = 8;
        fmt->ebits = r.e;
        fmt->mbits = r.m;
        fmt->max_value = r.e == 4 ? 448.0f : 57344.0f;
        fmt->min_value = -fmt->max_value;
        return QT_OK;
    case Parsed::FPMX:
        // The reference evaluates this family in bf16 arithmetic.  Its result equals exact saturating
        // RNE plus three documented quirks for 2 <= ebits <= 5, 1 <= mbits <= 5, ebits + mbits <= 8
        // (checked against the reference's tables and, parameter by parameter, in the test-suite);
        // outside that box bf16 rounding inside the reference algorithm (log2, +0.5) changes values in
        // format-specific ways this library does not claim to reproduce, so it refuses them.
        if (!(r.n == r.e + r.m + 1 || r.n == r.e + r.m)) break;
        if (!qt_mx_supported(r.e, r.m)) break;
        fmt->kind = QT_KIND_FP;
        fmt->flavour = QT_FP_MX;
        fmt->nbits = r.n;
        fmt->ebits = r.e;
        fmt->mbits = r.m;
        fmt->is_unsigned = (r.n == r.e + r.m);
        fmt->max_value = (float)mx_max_norm(dtype, r.e, r.m);
        fmt->min_value = fmt->is_unsigned ? 0.0f : -fmt->max_value;
        return QT_OK;
    case Parsed::POSIT:
        if (r.n < 3 || r.n > 24 || r.e < 0 || r.e > 4) break;
        if ((r.n - 2) * (1 << r.e) > 126) break;  // maxpos must be a finite normal fp32/bf16
        fmt->kind = QT_KIND_POSIT;
        fmt->nbits = r.n;
        fmt->ebits = r.e;
        fmt->max_value = (float)ldexp(1.0, (r.n - 2) * (1 << r.e));
        fmt->min_value = -fmt->max_value;
        return QT_OK;
    default:
        break;
    }
    memset(fmt, 0, sizeof(*fmt));
    qt_set_error("Unsupported dtype: %s", dtype ? dtype : "(null)");
    return QT_ERR_UNSUPPORTED_DTYPE;
}

extern "C" int qt_format_min_max(const char *dtype, double *qmin, double *qmax)
{
    Parsed r = parse_dtype(dtype, true);
    double hi = 0, lo = 0;
    switch (r.fam) {
    case Parsed::INT:
        hi = ldexp(1.0, r.n - 1) - 1;
        lo = -ldexp(1.0, r.n - 1);
        break;
    case Parsed::UINT:
        hi = ldexp(1.0, r.n) - 1;
        lo = 0;
        break;
    case Parsed::FPMX:
        hi = mx_max_norm(dtype, r.e, r.m);
        lo = -hi;
        break;
    case Parsed::POSIT:
        hi = pow(pow(2.0, (double)(1 << r.e)), (double)(r.n - 2));
        lo = -hi;
        break;
    case Parsed::NF:
        hi = r.e > 0 ? ldexp(1.0, r.e - 1) - 1 : 1.0;
        lo = -hi;
        break;
    default:
        qt_set_error("Unsupported dtype: %s", dtype ? dtype : "(null)");
        return QT_ERR_UNSUPPORTED_DTYPE;
    }
    if (qmin) *qmin = lo;
    if (qmax) *qmax = hi;
    return QT_OK;
}

int qt_make_round(const qt_format_t *fmt, QtRound *P)
{
    memset(P, 0, sizeof(*P));
    switch (fmt->kind) {
    case QT_KIND_IDENTITY:
        P->kind = QTR_IDENTITY;
        return QT_OK;
    case QT_KIND_INT:
        P->kind = QTR_INT;
        P->qmin = bf16_round(fmt->min_value);
        P->qmax = bf16_round(fmt->max_value);
        P->tiny_safe = (P->qmin <= 0.0f && P->qmax >= 0.0f) ? 1 : 0;
        return QT_OK;
    case QT_KIND_FP: {
        if (fmt->flavour == QT_FP_MX ? !qt_mx_supported(fmt->ebits, fmt->mbits)
                                     : !((fmt->ebits == 4 && fmt->mbits == 3) || (fmt->ebits == 5 && fmt->mbits == 2)))
            break;
        P->kind = fmt->flavour == QT_FP_MX ? QTR_FP_MX : QTR_FP_CUSTOM;
        P->mshift = 23 - fmt->mbits;
        int min_exp = (fmt->flavour == QT_FP_MX) ? -(1 << (fmt->ebits - 1)) + 2 : (fmt->ebits == 4 ? -6 : -14);
        P->min_exp_biased = (uint32_t)(min_exp + 127);
        P->max_bits = fbits(fmt->max_value);
        P->sign_mask = fmt->is_unsigned ? 0u : 0x80000000u;
        float min_sub = (float)ldexp(1.0, min_exp - fmt->mbits);
        P->min_sub_bits = fbits(min_sub);
        // the bf16 value immediately below half the smallest subnormal rounds up (fp8.py:123-127 in bf16)
        P->quirk_bits = fmt->flavour == QT_FP_MX ? fbits(min_sub * 0.5f) - 0x10000u : 0u;
        P->tiny_safe = 1;  // half the smallest subnormal is >= 2^-21 for every supported (ebits, mbits)
        return QT_OK;
    }
    case QT_KIND_POSIT: {
        int n = fmt->nbits, es = fmt->ebits;
        if (n < 3 || n > 24 || es < 0 || es > 4) break;
        int max_scale = (n - 2) * (1 << es);
        P->kind = QTR_POSIT;
        P->es_shift = 23 + es;
        P->c0 = 25 + es - n;
        P->maxpos_bits = fbits((float)ldexp(1.0, max_scale));
        P->minpos_bits = fbits((float)ldexp(1.0, -max_scale));
        double thr = floor(-(double)(n - 1) * (double)(1 << es) + pow(2.0, es - 1));
        P->flush_bits = fbits(bf16_round((float)ldexp(1.0, (int)thr)));
        P->tiny_safe = P->flush_bits >= 0x03800000u ? 1 : 0;  // everything below 2^-120 is flushed to zero
        return QT_OK;
    }
    default:
        break;
    }
    qt_set_error("qt_format_t is not a valid format (kind=%d nbits=%d ebits=%d mbits=%d)", fmt->kind, fmt->nbits,
                 fmt->ebits, fmt->mbits);
    return QT_ERR_INVALID_ARGUMENT;
}

extern "C" int qt_table_host(const qt_format_t *fmt, uint16_t *table_host)
{
    {
        QtRound chk;
        if (fmt && qt_make_round(fmt, &chk) == QT_OK && chk.tiny_safe && !qt_tiny_safe(chk)) {
            qt_set_error("internal: tiny_safe flag inconsistent for this format");
            return QT_ERR_INVALID_ARGUMENT;
        }
    }
    if (!fmt || !table_host) {
        qt_set_error("qt_table_host: NULL argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    QtRound P;
    int rc = qt_make_round(fmt, &P);
    if (rc != QT_OK) return rc;
    for (uint32_t i = 0; i < 65536u; ++i) table_host[i] = (uint16_t)(qt_round_dyn(P, i << 16) >> 16);
    return QT_OK;
}

// ---- force_scale_power_of_two of the microscaling qscheme --------------------------------------------------
// shared_exp = floor(log2(amax)) evaluated in the tensor's dtype (mx_utils.py:44-48).  For each exponent the
// result is e - 127 up to some mantissa and e - 126 from there on (log2 rounds up to the next integer); the
// thresholds are found by bisection over the same libm call the reference's CPU kernel makes.
static float floor_log2_in_dtype(uint32_t bits, bool f32)
{
    float a;
    memcpy(&a, &bits, 4);
    float l = log2f(a);
    if (!f32) l = bf16_round(l);
    return floorf(l);
}

extern "C" int qt_block_pow2_table_host(int elem_type, uint32_t *table_host)
{
    if (!table_host || (elem_type != QT_BF16 && elem_type != QT_F32)) {
        qt_set_error("qt_block_pow2_table_host: invalid argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const bool f32 = elem_type == QT_F32;
    const uint32_t step = f32 ? 1u : 0x10000u;  // bf16 values have the low 16 bits clear
    for (int i = 0; i < QT_POW2_TABLE_WORDS; ++i) table_host[i] = 0xFFFFFFFFu;
    for (uint32_t e = 1; e <= 254; ++e) {
        const float base = (float)((int)e - 127);
        // first mantissa (multiple of step) with floor(log2) > e - 127, or 2^23 if none
        uint32_t lo = 0, hi = 0x800000u / step;  // search over m / step in [lo, hi)
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (floor_log2_in_dtype((e << 23) | (mid * step), f32) > base)
                hi = mid;
            else
                lo = mid + 1;
        }
        table_host[e] = lo * step;  // 0x800000 = never
    }
    for (int k = 0; k <= 22; ++k) {
        // subnormal patterns [2^k, 2^(k+1)): value = pattern * 2^-149; bf16 subnormals use bits 16..22 only
        const uint32_t first = 1u << k, last = (2u << k);
        if (!f32 && k < 16) continue;
        const float base = (float)(k - 149);
        uint32_t lo = first / step, hi = last / step;
        if (lo == 0) lo = 1;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (floor_log2_in_dtype(mid * step, f32) > base)
                hi = mid;
            else
                lo = mid + 1;
        }
        table_host[256 + k] = lo * step;  // == last: never
    }
    return QT_OK;
}

// brute-force statement of QtRound::tiny_safe (the analytic flag set by qt_make_round is checked against it in
// qt_lut_build_host and qt_table_host, i.e. once per module, never per launch)
int qt_tiny_safe(const QtRound &P)
{
    for (uint32_t sign = 0; sign < 2; ++sign) {
        const uint32_t first = qt_round_dyn(P, (sign << 31) | (1u << 16));
        for (uint32_t pat = 1; pat < 0x0380u; ++pat)  // bf16 patterns below 2^-120 (exponent field < 7)
            if (qt_round_dyn(P, (sign << 31) | (pat << 16)) != first) return 0;
    }
    return 1;
}

extern "C" float qt_scale_pow2_host(float sf)
{
    const float p2 = qt_pow2_ceil(sf);
    return p2 >= 0.0f ? p2 : exp2f(ceilf(log2f(sf)));
}
