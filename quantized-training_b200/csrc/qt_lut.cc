// qt_lut.cc -- host-side derivation of the per-binade constants (see qt_lut.h) from the bitwise
// rounding functions of qt_round.h, followed by an exhaustive check of the result.
#include <math.h>
#include <string.h>

#include "qt_internal.h"
#include "qt_lut.h"

namespace {

inline bool same_bits(uint32_t a, uint32_t b)
{
    const bool na = (a & 0x7FFFFFFFu) > 0x7F800000u, nb = (b & 0x7FFFFFFFu) > 0x7F800000u;
    return (a == b) || (na && nb);
}
inline bool is_nan_bits(uint32_t a) { return (a & 0x7FFFFFFFu) > 0x7F800000u; }

struct Binade {
    uint32_t u[128];    // inputs (fp32 bits)
    uint32_t ac[128];   // |input| after the optional clamp: what the FMAs see
    uint32_t ref[128];  // what the bitwise spec returns
    bool care[128];     // false where a post-patch overrides the table (fpN_eXmY NaN band)
};

bool check(const QtLutEntry &e, const Binade &b)
{
    for (int m = 0; m < 128; ++m) {
        if (!b.care[m]) continue;
        const float t = qt_saturate(qt_fma(qt_bits2f(b.ac[m]), e.p1, e.p2));
        const uint32_t q = qt_f2bits(qt_fma(t, e.d, e.l));
        if (!same_bits(q, b.ref[m])) return false;
    }
    return true;
}

bool fit_constant(const Binade &b, QtLutEntry *e)
{
    int first = -1;
    for (int m = 0; m < 128; ++m) {
        if (!b.care[m]) continue;
        if (first < 0) first = m;
        if (!same_bits(b.ref[m], b.ref[first])) return false;
    }
    const uint32_t v = first < 0 ? 0u : b.ref[first];
    e->p1 = 0.0f;
    e->p2 = 0.0f;
    e->l = qt_bits2f(v);
    e->d = (v >> 31) ? -1.0f : 0.0f;  // 0 * d keeps the sign of a -0.0 constant
    return check(*e, b);
}

bool fit_step(const Binade &b, int s, QtLutEntry *e)
{
    // exactly one switch point between two values, both finite
    int j = -1;
    int prev = -1;
    for (int m = 0; m < 128; ++m) {
        if (!b.care[m]) continue;
        if (prev >= 0 && !same_bits(b.ref[m], b.ref[prev])) {
            if (j >= 0) return false;
            j = m;
        }
        prev = m;
    }
    if (j <= 0) return false;
    int last_lo = j - 1;
    while (last_lo >= 0 && !b.care[last_lo]) --last_lo;
    if (last_lo < 0) return false;
    const uint32_t v0 = b.ref[last_lo], v1 = b.ref[j];
    if (is_nan_bits(v0) || is_nan_bits(v1)) return false;
    if (9 - s > 127 || 9 - s < -126) return false;
    const float big = ldexpf(1.0f, 9 - s);
    const float lo_in = qt_bits2f(b.ac[last_lo]), hi_in = qt_bits2f(b.ac[j]);
    if (!(hi_in > lo_in)) return false;
    const float thr = 0.5f * lo_in + 0.5f * hi_in;  // exact: neighbours on the bf16 grid
    const double dd = (double)qt_bits2f(v1) - (double)qt_bits2f(v0);
    if ((double)(float)dd != dd) return false;
    e->p1 = big;
    e->p2 = -thr * big;
    e->d = (float)dd;
    e->l = qt_bits2f(v0);
    return check(*e, b);
}

bool fit_rne(const Binade &b, int s, QtLutEntry *e)
{
    // output polarity comes from the spec (unsigned scale formats map negative inputs to positive values)
    bool negative = false;
    for (int m = 0; m < 128; ++m)
        if (b.care[m]) {
            negative = (b.ref[m] >> 31) != 0;
            break;
        }
    for (int fb = 7; fb >= 0; --fb) {
        const int me = s - fb + 23;  // M = 2^me
        if (me + 1 > 126 || me + 1 < -125) continue;
        const float M = ldexpf(1.0f, me);
        e->p1 = ldexpf(1.0f, -(me + 1));
        e->p2 = 0.5f;
        e->d = negative ? -2.0f * M : 2.0f * M;
        e->l = negative ? M : -M;
        if (check(*e, b)) return true;
    }
    return false;
}

}  // namespace

int qt_lut_config(const QtRound &P, QtLutCfg *cfg)
{
    cfg->clamp_bits = 0x7FFFFFFFu;
    cfg->mx_band = 0;
    cfg->tiny_safe = (uint32_t)P.tiny_safe;
    switch (P.kind) {
    case QTR_FP_CUSTOM: cfg->clamp_bits = P.max_bits; return QT_OK;
    case QTR_FP_MX:
        cfg->clamp_bits = P.max_bits;
        cfg->mx_band = 1;
        return QT_OK;
    case QTR_POSIT: return QT_OK;
    default: return QT_NO_LUT;  // int / identity are a handful of native instructions already
    }
}

extern "C" int qt_lut_build_host(const qt_format_t *fmt, void *lut_host)
{
    if (!fmt || !lut_host) {
        qt_set_error("qt_lut_build_host: NULL argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    QtRound P;
    int rc = qt_make_round(fmt, &P);
    if (rc != QT_OK) return rc;
    QtLutCfg cfg;
    if (qt_lut_config(P, &cfg) != QT_OK) return QT_NO_LUT;
    QtLutEntry *tab = static_cast<QtLutEntry *>(lut_host);
    memset(tab, 0, QT_LUT_BYTES);

    for (uint32_t idx = 0; idx < QT_LUT_ENTRIES; ++idx) {
        const bool negative = (idx >> 8) != 0;
        const int E = (int)(idx & 0xFF);
        Binade b;
        for (int m = 0; m < 128; ++m) {
            b.u[m] = (idx << 23) | ((uint32_t)m << 16);
            const uint32_t a = b.u[m] & 0x7FFFFFFFu;
            b.ac[m] = qt_umin(a, cfg.clamp_bits);
            b.ref[m] = qt_round_dyn(P, b.u[m]);
            b.care[m] = !(cfg.mx_band && a >= 0x7F580000u && a != 0x7F800000u);
        }
        QtLutEntry e;
        bool ok = fit_constant(b, &e);
        if (!ok && E == 0) {
            // zero and bf16 subnormals: the only non-constant case is fpN_eXmY on the negative side
            // (-0.0 -> +0.0 but a negative subnormal -> -0.0).  t in (0, 2^-26] for non-zero inputs, and the
            // fused product t * -2^-149 underflows to -0.0; for a zero input fma(0, d, +0) is +0.0.
            e.p1 = ldexpf(1.0f, 100);
            e.p2 = 0.0f;
            e.d = -ldexpf(1.0f, -149);
            e.l = 0.0f;
            ok = check(e, b);
        }
        if (!ok && E >= 1 && E <= 254) ok = fit_step(b, E - 127, &e) || fit_rne(b, E - 127, &e);
        if (!ok) {
            qt_set_error("no binade-constant form for sign=%d exponent=%d of this format; use the direct path",
                         (int)negative, E);
            return QT_NO_LUT;
        }
        tab[idx] = e;
    }
    // the table must reproduce the bitwise spec on every bf16 input
    for (uint32_t i = 0; i < 65536u; ++i) {
        const uint32_t u = i << 16;
        if (!same_bits(qt_lut_round_dyn(tab, cfg, u), qt_round_dyn(P, u))) {
            qt_set_error("binade-constant table disagrees with the bitwise rounding at bf16 pattern 0x%04x", i);
            return QT_NO_LUT;
        }
    }
    return QT_OK;
}
