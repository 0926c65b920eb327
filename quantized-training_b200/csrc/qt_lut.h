// qt_lut.h -- "binade constants" form of the rounding logic (the fast path of the kernels).
//
// For a bf16 input every format of this library behaves, inside one binade (fixed sign and
// exponent, 128 mantissa values), in one of three ways:
//   A  round-to-nearest-even to a fixed quantum 2^(s - fb)            (fraction bits are dropped)
//   B  a two-valued step: L below a threshold, U at/above it          (exponent bits are dropped: posit
//      regime edges with their geometric tie points; sub-minimum values; the fpN_eXmY quirk)
//   C  a constant                                                     (saturation, flush to zero, NaN)
// All three are the same two fused multiply-adds with four per-binade constants:
//        t = saturate(fma(|x|, p1, p2));      q = fma(t, d, l)
//   A: p1 = 1/(2M), p2 = 1/2, d = +-2M, l = -+M with M = 2^23 * quantum: the first FMA lands in
//      [1/2, 1) where one fp32 ulp is exactly one quantum, so the hardware's RNE does the rounding
//      (ties to even on the same bit the bit-string algorithm looks at); the second FMA undoes the shift.
//   B: p1 = BIG, p2 = -T*BIG with T between the last "L" input and the first "U" input: the first FMA
//      saturates to exactly 0 or 1;  d = U - L, l = L.
//   C: p1 = p2 = 0 (t = 0 even for Inf/NaN inputs, saturate(NaN) = 0), l = the constant.
// The sign lives in the table (index = sign:exponent, 512 entries), so signed zeros come out right:
// fma(0, d, +0) is +0 (posit, e4m3), and d < 0 with l = -0 gives -0 (fpN_eXmY tiny negatives).
// 512 x 16 B = 8 KB, staged in shared memory by every CTA.  This moves the per-element work from the
// ALU pipe (about 30 integer ops for the posit bit-string algorithm) to 2 FFMA + 1 LDS.128.
//
// The table is DERIVED from the bitwise functions of qt_round.h on the host and then verified against
// them on all 65 536 inputs (qt_lut_build_host); a format whose binades do not fit is reported as such
// and runs on the direct bitwise path.
#pragma once
#include "../../include/qt_b200.h"
#include "qt_round.h"

#define QT_LUT_ENTRIES 512
static_assert(QT_LUT_BYTES == QT_LUT_ENTRIES * 16, "QT_LUT_BYTES (include/qt_b200.h) must hold 512 entries");

struct QtLutEntry {
    float p1, p2, d, l;
};

// per-format switches of the table path (derived from QtRound on the host)
struct QtLutCfg {
    uint32_t clamp_bits;  // |x| is clamped to this bit pattern first (max_norm); 0x7FFFFFFF = no clamp
    uint32_t mx_band;     // 1: |x| >= 0x7F58 (bf16) other than Inf gives NaN -- fpN_eXmY only
    uint32_t tiny_safe;   // QtRound::tiny_safe
};

QT_HD float qt_saturate(float x)
{
#if defined(__CUDA_ARCH__)
    return __saturatef(x);
#else
    if (!(x == x)) return 0.0f;
    return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
#endif
}
QT_HD float qt_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// u: fp32 bits of a bf16-representable value.  Same contract as qt_round<KIND>().
template <bool CLAMP, bool MXBAND>
QT_HD uint32_t qt_lut_round(const QtLutEntry *tab, const QtLutCfg &cfg, uint32_t u)
{
    const uint32_t a = u & 0x7FFFFFFFu;
    const uint32_t ac = CLAMP ? qt_umin(a, cfg.clamp_bits) : a;
    const QtLutEntry e = tab[u >> 23];
    const float t = qt_saturate(qt_fma(qt_bits2f(ac), e.p1, e.p2));
    uint32_t q = qt_f2bits(qt_fma(t, e.d, e.l));
    if (MXBAND) {
        if (a >= 0x7F580000u && a != 0x7F800000u) q = QT_NAN_BITS;  // finite band and NaN; Inf comes from the table
    }
    return q;
}

#if defined(__CUDACC__)
// Shared-memory form used by the kernels.  A warp-wide LDS.128 is served in four 128-byte wavefronts, one per
// quarter warp, and two lanes of a quarter warp that need DIFFERENT entries from the same 16-byte bank group
// serialise (measured: 7.8 wavefronts per load with a plain 8 KB table, LSU data pipe 96 % busy).  So the table
// is replicated eight times, interleaved: replica r of entry i lives at byte (i * 8 + r) * 16, and lane l always
// reads replica l & 7 -- every lane of a quarter warp owns its bank group, conflict-free for any data.
// `slot16` = (lane & 7) * 16 is folded into the masking instruction, so the index costs one LOP3 for the low
// bf16 of a packed word and SHF + LOP3 for the high one.
#define QT_LUT_REPLICAS 8
#define QT_LUT_SMEM_BYTES (QT_LUT_BYTES * QT_LUT_REPLICAS)

// hi16: the bf16 bit pattern in bits [15:0] (anything above is ignored); ac: clamped |x| as fp32 bits.
// REPL = 8: the replicated layout above (streaming kernels).  REPL = 1: a plain 8 KB table, entry i at byte i * 16
// (fused row / elementwise kernels: small launches where staging 64 KB per CTA would dominate, and whose inputs
// span few binades, so the residual bank conflicts are rare).
template <bool MXBAND, int REPL = QT_LUT_REPLICAS>
__device__ __forceinline__ uint32_t qt_lut_round_smem(const unsigned char *smem_table, uint32_t slot16,
                                                      uint32_t pattern16, uint32_t a, uint32_t ac)
{
    static_assert(REPL == 8 || REPL == 1, "table layouts: 8 interleaved replicas or one");
    // (sign:exponent) * 128 + replica * 16, or (sign:exponent) * 16
    const uint32_t off = REPL == 8 ? ((pattern16 & 0xFF80u) | slot16) : ((pattern16 >> 3) & 0x1FF0u);
    const float4 e = *reinterpret_cast<const float4 *>(smem_table + off);
    const float t = __saturatef(__fmaf_rn(__uint_as_float(ac), e.x, e.y));
    uint32_t q = __float_as_uint(__fmaf_rn(t, e.z, e.w));
    if (MXBAND) {
        if (a >= 0x7F580000u && a != 0x7F800000u) q = QT_NAN_BITS;
    }
    return q;
}
#endif

QT_HD uint32_t qt_lut_round_dyn(const QtLutEntry *tab, const QtLutCfg &cfg, uint32_t u)
{
    if (cfg.mx_band) return qt_lut_round<true, true>(tab, cfg, u);
    if (cfg.clamp_bits != 0x7FFFFFFFu) return qt_lut_round<true, false>(tab, cfg, u);
    return qt_lut_round<false, false>(tab, cfg, u);
}
