// qt_round.h -- bitwise round-to-format logic, one scalar function per format family.
//
// This is the product's replacement for the reference's 65 536-entry lookup table
// (get_quantization_map, fake_quantize.py:31-95): every function maps the fp32 bit
// pattern of a bf16-representable value (low 16 bits zero) to the fp32 bit pattern
// of the rounded value, using only integer/fp32 ALU operations -- no table.
// The functions are __host__ __device__ so the very same logic can be evaluated on
// the CPU over all 2^16 inputs (qt_table_host) and compared with the reference's
// tables in the CPU test-suite, before any GPU time is spent.
//
// What each family has to reproduce (derived from the reference code, verified
// against its tables):
//   INT    torch.clamp(torch.round(v), qmin, qmax) on a bf16 tensor: half-even,
//          -0.0 survives, NaN stays NaN, +-Inf clamp.
//   FP     saturating RNE with subnormals.  CUSTOM flavour (fp8.py:10-67): non-finite
//          -> NaN, every zero result is +0.  MX flavour (fp8.py:147-203 run in bf16):
//          +-Inf pass through, |x| >= 0x7F58 (bf16) -> NaN, the one bf16 value just
//          below half the smallest subnormal rounds UP, result keeps the input's sign
//          even when it is zero (except for +-0 inputs, which give +0).
//   POSIT  posit.py:6-67: round-half-even on the posit BIT STRING (regime|exponent|
//          fraction), which is an RNE of the integer  X = bits - 0x3F800000  at a
//          regime-dependent bit position; saturate at minpos/maxpos, flush below
//          2^floor(-(n-1)2^es + 2^(es-1)), non-finite -> NaN, zero is +0.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#else
#define QT_HD inline
#endif

#define QTR_IDENTITY 0
#define QTR_INT 1
#define QTR_FP_CUSTOM 2
#define QTR_FP_MX 3
#define QTR_POSIT 4

// Kernel-side parameters, derived on the host from qt_format_t (see qt_make_round()).
struct QtRound {
    int32_t kind;  // QTR_*
    // INT
    float qmin, qmax;  // already rounded to bf16, as torch.clamp does on a bf16 tensor
    // FP
    int32_t mshift;           // 23 - mbits
    uint32_t min_exp_biased;  // fp32 biased exponent of the smallest normal
    uint32_t max_bits;        // fp32 bits of max_norm
    uint32_t quirk_bits;      // MX: |x| bits that round up to min_sub (0 = none)
    uint32_t min_sub_bits;    // MX: fp32 bits of the smallest subnormal
    uint32_t sign_mask;       // 0x80000000, or 0 for unsigned formats
    // POSIT
    int32_t es_shift;      // 23 + es
    int32_t c0;            // 25 + es - nbits
    uint32_t minpos_bits;  // fp32 bits of 2^-max_scale
    uint32_t maxpos_bits;  // fp32 bits of 2^max_scale
    uint32_t flush_bits;   // |x| bits below which the result is 0
    // 1: round(u) of a bf16 u with 0 < |u| < 2^-120 depends on the sign of u only, so a quotient in that range may
    // be computed inexactly (x * rcp(s)) as long as its sign and zero-ness are right (checked by brute force in
    // qt_lut_build_host / tests)
    int32_t tiny_safe;
};

QT_HD float qt_bits2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
QT_HD uint32_t qt_f2bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
QT_HD float qt_rint(float v)
{
#if defined(__CUDA_ARCH__)
    return rintf(v);
#else
    return __builtin_nearbyintf(v);
#endif
}
QT_HD float qt_add_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;  // keep the two roundings separate
    return r;
#endif
}
QT_HD uint32_t qt_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
QT_HD uint32_t qt_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
QT_HD int32_t qt_imin(int32_t a, int32_t b) { return a < b ? a : b; }

#define QT_NAN_BITS 0x7FC00000u

QT_HD uint32_t qt_round_int(const QtRound &P, uint32_t u)
{
    float r = qt_rint(qt_bits2f(u));
    r = (r < P.qmin) ? P.qmin : r;  // compare-select keeps NaN and -0.0 exactly as torch.clamp does
    r = (r > P.qmax) ? P.qmax : r;
    return qt_f2bits(r);
}

template <bool MX>
QT_HD uint32_t qt_round_fp(const QtRound &P, uint32_t u)
{
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t sign = u & P.sign_mask;
    // saturate first: max_norm is a grid point and rounding is monotone
    const uint32_t ac = qt_umin(a, P.max_bits);
    // RNE to a multiple of 2^(e - mbits) by adding and subtracting 2^(e + 23 - mbits)
    const uint32_t e = qt_umax(ac >> 23, P.min_exp_biased);
    const float magic = qt_bits2f((e + (uint32_t)P.mshift) << 23);
    const float r = qt_add_rn(qt_add_rn(qt_bits2f(ac), magic), -magic);
    uint32_t q = qt_f2bits(r);
    if (MX) {
        if (a == P.quirk_bits) q = P.min_sub_bits;
        q = (a == 0u) ? 0u : (q | sign);
        if (a >= 0x7F580000u) q = (a == 0x7F800000u) ? (a | sign) : QT_NAN_BITS;
    } else {
        q = (q == 0u) ? 0u : (q | sign);
        if (a >= 0x7F800000u) q = QT_NAN_BITS;
    }
    return q;
}

QT_HD uint32_t qt_round_posit(const QtRound &P, uint32_t u)
{
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t sign = u & 0x80000000u;
    const uint32_t ac = qt_umin(qt_umax(a, P.minpos_bits), P.maxpos_bits);
    const int32_t X = (int32_t)(ac - 0x3F800000u);  // scale * 2^23 + fraction
    const int32_t k = X >> P.es_shift;              // regime value
    const int32_t run = (k ^ (k >> 31)) + 1;        // k >= 0 ? k + 1 : -k
    const int32_t sh = qt_imin(run + P.c0, P.es_shift);
    // last kept bit of the posit string: a fraction/exponent bit of X, or -- when every
    // exponent and fraction bit is dropped -- the last regime bit (0 after a run of ones)
    const int32_t lb = (sh == P.es_shift) ? (int32_t)((uint32_t)k >> 31) : ((X >> sh) & 1);
    const int32_t hm1 = (int32_t)((1u << (sh - 1)) - 1u);
    const int32_t X2 = (X + hm1 + lb) & (int32_t)(0xFFFFFFFFu << sh);
    uint32_t q = qt_umin((uint32_t)X2 + 0x3F800000u, P.maxpos_bits);
    if (a < P.flush_bits) q = 0u;
    q = (q == 0u) ? 0u : (q | sign);
    if (a >= 0x7F800000u) q = QT_NAN_BITS;
    return q;
}

template <int KIND>
QT_HD uint32_t qt_round(const QtRound &P, uint32_t u)
{
    if (KIND == QTR_INT) return qt_round_int(P, u);
    if (KIND == QTR_FP_CUSTOM) return qt_round_fp<false>(P, u);
    if (KIND == QTR_FP_MX) return qt_round_fp<true>(P, u);
    if (KIND == QTR_POSIT) return qt_round_posit(P, u);
    return u;
}

QT_HD uint32_t qt_round_dyn(const QtRound &P, uint32_t u)
{
    switch (P.kind) {
    case QTR_INT: return qt_round_int(P, u);
    case QTR_FP_CUSTOM: return qt_round_fp<false>(P, u);
    case QTR_FP_MX: return qt_round_fp<true>(P, u);
    case QTR_POSIT: return qt_round_posit(P, u);
    default: return u;
    }
}

// ---- force_scale_power_of_two of the per-tensor / per-channel scheme (fake_quantize.py:240-241) ---------------
// scale = 2 ** ceil(log2(sf)) evaluated in fp32 by the reference.  log2() of sf = 2^e (1 + m 2^-23) is e + t with
// t ~ 1.4427 m 2^-23, and fl32(e + t) == e whenever t is below half an ulp of e: for |e| >= 4 the first few
// mantissas above a power of two come out as ceil = e, not e + 1.  Restated exactly (no log2 on the device, whose
// last-ulp behaviour differs from the host's libm): with |e| in [2^k, 2^(k+1)), ceil = e iff m < ln2 * 2^(k-1)
// (ln2 * 2^(k-2) when e is a negative power of two, where the spacing below |e| is half as wide).
// Checked against libm on every exponent in tests (qt_scale_pow2_host).
QT_HD float qt_pow2_ceil(float sf)
{
    const uint32_t b = qt_f2bits(sf);
    const uint32_t ef = (b >> 23) & 0xFFu, m = b & 0x7FFFFFu;
    if ((b >> 31) || ef == 0u || ef == 255u) return -1.0f;  // caller falls back (sign, zero / subnormal, Inf / NaN)
    int e = (int)ef - 127;
    if (m != 0u) {
        const int ae = e < 0 ? -e : e;
        bool stays = false;
        if (ae >= 4) {
            int k = 2;
            while ((2 << k) <= ae) ++k;                       // ae in [2^k, 2^(k+1))
            float thr = 0.69314718f * (float)(1 << (k - 1));  // ln2 * 2^(k-1)
            if (e < 0 && ae == (1 << k)) thr *= 0.5f;
            stays = (float)m < thr;
        }
        if (!stays) ++e;
    }
    if (e > 127) return qt_bits2f(0x7F800000u);
    return qt_bits2f((uint32_t)(e + 127) << 23);  // e >= -126 here
}
