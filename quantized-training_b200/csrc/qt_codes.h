// qt_codes.h -- one-byte CODES of the <= 8-bit formats: the storage form of quantized GEMM operands.
//
// The reference emits codes for one family only -- posit, `return_pbits` (posit.py:60-65: the posit bit string,
// negated for negative inputs) -- and keeps everything else as bf16 values; SURVEY.md App. A leaves the other layouts
// to the build "subject to decode(code) == qmap[idx]".  The layouts defined here (QT_CODE_NATIVE):
//   positN_ES  N <= 8   the N-bit posit word (regime | exponent | fraction), two's complement for negative values,
//                       sign-extended to 8 bits == the reference's pbits;  NaR (0x80 for N = 8) for NaN
//   intN       N <= 8   the integer, two's complement, sign-extended;      uintN: the integer
//   fp formats N <= 8   sign | biased exponent | mantissa (bias 2^(e-1) - 1, subnormals; the OCP layouts of
//                       E4M3 / E5M2 / E3M2 / E2M3 / E2M1), in the low N bits
// Values with no code in their format -- NaN of an int / fp6 / fp4 tensor, +-Inf passed through by the fpN_eXmY
// flavour of a format without an Inf encoding -- get the code of +0 (NaN) / of +-max (Inf): a quantized GEMM operand
// holding them is meaningless either way; everything else round-trips exactly, which tests check on all 2^16 inputs.
// Container codes (QT_CODE_E4M3 / QT_CODE_E5M2): the OCP fp8 encoding of the SAME value, available when every value of
// the format is an e4m3 (e5m2) value -- fp6_e3m2, fp6_e2m3, fp4_e2m1, int2..int5, uint2..uint4 are subsets of e4m3 --
// so that those operands run on the FP8 tensor cores at twice the bf16 rate with no decode at all.
//
// encode: __host__ __device__, from the fp32 bits of a value that IS a member of the format (a rounder's output).
// decode: host only, used to build the 256-entry table the GEMM's decode warps read.
#pragma once
#include "qt_round.h"

struct QtCode {
    int32_t kind;   // QTR_*
    int32_t nbits;  // total bits of the code (<= 8 for storage; posit up to 16 is accepted by the functions)
    int32_t ebits;  // fp: exponent bits; posit: es
    int32_t mbits;  // fp: mantissa bits
    int32_t is_unsigned;
};

// ---------------------------------------------------------------- posit
QT_HD int32_t qt_encode_posit(int n, int es, uint32_t q)
{
    const uint32_t a = q & 0x7FFFFFFFu;
    if (a == 0u) return 0;
    if (a > 0x7F800000u) return -(1 << (n - 1));  // NaR
    const int32_t e = (int32_t)(a >> 23) - 127;    // a posit value is a normal fp32 number
    const uint32_t mant = a & 0x7FFFFFu;
    const int32_t k = e >> es;                     // regime value (floor)
    const uint32_t ex = (uint32_t)(e - (k << es));
    int32_t reglen;
    uint32_t regime;
    if (k >= 0) {
        reglen = k + 2;                            // k + 1 ones, one zero
        regime = ((1u << (k + 1)) - 1u) << 1;
    } else {
        reglen = -k + 1;                           // -k zeros, one one
        regime = 1u;
    }
    const int32_t rem = n - 1 - reglen;            // bits left for exponent and fraction
    uint32_t code;
    if (rem <= 0) {
        code = regime >> (-rem);                   // regime fills the word (the terminating bit may fall off)
    } else {
        // exponent (es bits) followed by the fraction (23 bits): keep the top `rem` bits -- the rest is zero for a
        // member of the format
        const uint64_t body = ((uint64_t)ex << 23) | mant;
        const int32_t drop = es + 23 - rem;
        const uint32_t tail = drop >= 0 ? (uint32_t)(body >> drop) : (uint32_t)(body << (-drop));
        code = (regime << rem) | tail;
    }
    return (q >> 31) ? -(int32_t)code : (int32_t)code;
}
inline double qt_decode_posit(int n, int es, int32_t code)
{
    const int32_t nar = -(1 << (n - 1));
    if (code == 0) return 0.0;
    if (code == nar) return __builtin_nan("");
    const bool neg = code < 0;
    uint32_t c = (uint32_t)(neg ? -code : code) & ((1u << (n - 1)) - 1u);
    // regime
    int pos = n - 2;
    const int first = (c >> pos) & 1;
    int run = 0;
    while (pos >= 0 && (int)((c >> pos) & 1) == first) {
        ++run;
        --pos;
    }
    --pos;  // the terminating bit
    const int k = first ? run - 1 : -run;
    int ex = 0, got = 0;
    while (got < es && pos >= 0) {
        ex = (ex << 1) | ((c >> pos) & 1);
        --pos;
        ++got;
    }
    ex <<= (es - got);
    double frac = 1.0, w = 0.5;
    while (pos >= 0) {
        if ((c >> pos) & 1) frac += w;
        w *= 0.5;
        --pos;
    }
    const double v = __builtin_ldexp(frac, k * (1 << es) + ex);
    return neg ? -v : v;
}

// ---------------------------------------------------------------- integers
QT_HD int32_t qt_encode_int(uint32_t q)
{
    const float f = qt_bits2f(q);
    if (!(f == f)) return 0;  // NaN has no integer code
    return (int32_t)f;        // exact: q is an integer within the format's range
}

// ---------------------------------------------------------------- fp (sign | exponent | mantissa)
QT_HD int32_t qt_encode_fp(int ebits, int mbits, bool is_unsigned, uint32_t q)
{
    const uint32_t a = q & 0x7FFFFFFFu;
    const int32_t bias = (1 << (ebits - 1)) - 1;
    const uint32_t sign = (!is_unsigned && (q >> 31)) ? (1u << (ebits + mbits)) : 0u;
    const bool ieee = ebits > 4;                  // top exponent field holds Inf / NaN (E5M2); otherwise OCP "fn" style
    const uint32_t top = (1u << ebits) - 1u;
    if (a > 0x7F800000u) {                        // NaN
        if (ieee) return (int32_t)(sign | (top << mbits) | ((1u << mbits) - 1u));
        if (ebits == 4 && mbits == 3) return (int32_t)(sign | 0x7Fu);
        return 0;
    }
    if (a == 0x7F800000u) {                       // Inf: IEEE code, else the largest magnitude
        if (ieee) return (int32_t)(sign | (top << mbits));
        if (ebits == 4 && mbits == 3) return (int32_t)(sign | 0x7Eu);
        return (int32_t)(sign | (top << mbits) | ((1u << mbits) - 1u));
    }
    if (a == 0u) return (int32_t)sign;            // -0 keeps its sign bit (the MX flavour produces it)
    const int32_t e = (int32_t)(a >> 23) - 127;
    const uint32_t mant = a & 0x7FFFFFu;
    const int32_t field = e + bias;
    if (field >= 1) return (int32_t)(sign | ((uint32_t)field << mbits) | (mant >> (23 - mbits)));
    // subnormal: value = m * 2^(1 - bias - mbits)
    const uint32_t full = (mant | 0x800000u) >> (23 - mbits);   // 1.mant scaled to mbits fraction bits
    return (int32_t)(sign | (full >> (1 - field)));
}
inline double qt_decode_fp(int ebits, int mbits, bool is_unsigned, int32_t code)
{
    const int32_t bias = (1 << (ebits - 1)) - 1;
    const uint32_t c = (uint32_t)code;
    const bool neg = !is_unsigned && ((c >> (ebits + mbits)) & 1u);
    const uint32_t field = (c >> mbits) & ((1u << ebits) - 1u), mant = c & ((1u << mbits) - 1u);
    const uint32_t top = (1u << ebits) - 1u;
    double v;
    if (ebits > 4 && field == top)
        v = mant ? __builtin_nan("") : __builtin_inf();
    else if (ebits == 4 && mbits == 3 && field == top && mant == 7u)
        v = __builtin_nan("");
    else if (field == 0)
        v = __builtin_ldexp((double)mant, 1 - bias - mbits);
    else
        v = __builtin_ldexp(1.0 + (double)mant / (double)(1u << mbits), (int)field - bias);
    return neg ? -v : v;
}

// ---------------------------------------------------------------- dispatch
QT_HD int32_t qt_encode_native(const QtCode &C, uint32_t q)
{
    switch (C.kind) {
    case QTR_INT: return qt_encode_int(q);
    case QTR_POSIT: return qt_encode_posit(C.nbits, C.ebits, q);
    default: return qt_encode_fp(C.ebits, C.mbits, C.is_unsigned != 0, q);
    }
}
