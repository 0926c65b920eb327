// qt_block_common.cuh -- types and device helpers shared by the translation units of the block-scaled fake quant
// (qt_block.cu: entry points, generic / affine / table kernels; qt_block_flat.cu, qt_block_cols.cu, qt_block_tile.cu:
// the single-pass microscaling kernels, one family per file so that they compile in parallel).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "qt_fq_common.cuh"

// plain data handed between the translation units (named namespace: one type for all of them)
namespace qtblk {

struct BlockParams {
    float quant_min, quant_max, range;  // range = quant_max - quant_min (affine)
    int32_t pow2;                       // force_scale_power_of_two
    int32_t qmax_exp;                   // floor(log2(quant_max))
    int32_t has_scale_fmt;              // scale / zero point go through the scale_dtype codebook
    int32_t fast_ok;                    // the element format is insensitive to sub-2^-120 quotients (qt_tiny_safe)
    int32_t qmax_short;                 // quant_max has <= 8 significant bits: amax / quant_max by reciprocal multiply
    float rqmax;                        // 1 / quant_max
    QtRound scale_round;
    const uint32_t *pow2_tab;  // QT_POW2_TABLE_WORDS words on the device (pow2 only)
    const uint16_t *scale_table;  // 65 536-entry codebook of the scale given as a TABLE (operator surface), or null
};

struct BlockDims {
    size_t d0, n1, d1, n2, d2;
    size_t nb1, nb2;
    uint32_t bs, bs2;
};


struct BlockJob {
    const qt_block_desc_t *d;
    BlockDims D;
    BlockParams bp;
    cudaStream_t stream;
    bool f32;
};

// single-pass kernels; false = this layout is not theirs (qt_block_*.cu)
bool try_flat(const BlockJob &j, const QtRound &P);
bool try_cols(const BlockJob &j, const QtRound &P);
bool try_tile(const BlockJob &j, const QtRound &P);

}  // namespace qtblk

namespace {
using qtblk::BlockDims;
using qtblk::BlockJob;
using qtblk::BlockParams;

__device__ __forceinline__ float load_elem(const void *x, bool f32, size_t i)
{
    return f32 ? static_cast<const float *>(x)[i]
               : __uint_as_float((uint32_t) static_cast<const uint16_t *>(x)[i] << 16);
}



template <bool F32>
__device__ __forceinline__ float to_dtype(float v)
{
    return F32 ? v : __uint_as_float(bf16_rne_hi(v));
}
// vmap of one value of the tensor's dtype through the scale codebook
template <bool F32>
__device__ __forceinline__ float scale_codebook(const BlockParams &bp, float v)
{
    const uint32_t b = __float_as_uint(v);
    if (bp.scale_table) {
        const uint32_t idx = F32 ? (f32_to_bf16_rto_hi(b) >> 16) : (b >> 16);
        return __uint_as_float((uint32_t)__ldg(bp.scale_table + idx) << 16);
    }
    return __uint_as_float(qt_round_dyn(bp.scale_round, F32 ? f32_to_bf16_rto_hi(b) : b));
}

// scale of one block from the bit pattern of its amax (decomposed.py:391-419); NaN patterns order above Inf
template <bool F32>
__device__ __forceinline__ float mx_scale_of(uint32_t a, const BlockParams &bp)
{
    float s;
    if (bp.pow2) {
        if (a > 0x7F800000u) return 1.0f;  // log2(NaN) -> NaN -> where(scale > 0) picks 1
        if (a == 0x7F800000u) return __uint_as_float(a);
        int E;
        if (a == 0u)
            E = -126;  // amax + FP32_MIN_NORMAL * (amax == 0)
        else if (a >> 23) {
            const uint32_t e = a >> 23;
            E = (int)e - 127 + ((a & 0x7FFFFFu) >= bp.pow2_tab[e] ? 1 : 0);
        } else {
            const int k = 31 - __clz(a);
            E = k - 149 + (a >= bp.pow2_tab[256 + k] ? 1 : 0);
        }
        E -= bp.qmax_exp;
        if (E < (F32 ? -149 : -133)) return 1.0f;  // 2^E rounds to zero in the tensor's dtype
        if (E > 127) return __uint_as_float(0x7F800000u);
        s = __uint_as_float(E >= -126 ? (uint32_t)(E + 127) << 23 : 1u << (E + 149));
    } else {
        s = to_dtype<F32>(__fdiv_rn(__uint_as_float(a), bp.quant_max));
        if (bp.has_scale_fmt) s = scale_codebook<F32>(bp, s);
    }
    return s > 0.0f ? s : 1.0f;
}

__device__ __forceinline__ ScaleBf16 make_scale(float s)
{
    ScaleBf16 sc;
    sc.s = s;
    sc.rs = __frcp_rn(s);
    return sc;
}

// one 16-byte vector with one scale; the scale is already in the tensor's dtype
template <class R, bool F32>
__device__ __forceinline__ uint4 mx_apply_vec(const R &round, const uint4 &v, float s)
{
    const ScaleBf16 sc = make_scale(s);
    uint32_t unused = 0u;
    const int mode = classify_scale(s);
    if (mode == DIV_UNIT) return fq_vec<R, F32, DIV_UNIT, false>(round, v, sc, unused);
    if (F32 || mode == DIV_EXACT) return fq_vec<R, F32, DIV_EXACT, false>(round, v, sc, unused);
    return fq_vec<R, F32, DIV_RECIP, false>(round, v, sc, unused);
}

// The common case of mx_scale_of() with the rare inputs (zero, subnormal, Inf, NaN amax; scales outside the normal
// range) sent to it: the power-of-two branch is an exponent-field lookup in the staged threshold table, the
// amax / quant_max branch a reciprocal multiply (same argument as DIV_RECIP in qt_fq_common.cuh: a bf16 amax over a
// quant_max of at most 8 significant bits is never within 2^-17 of a bf16 rounding tie).
template <bool F32>
__device__ __forceinline__ float mx_scale_fast(uint32_t a, const BlockParams &bp, const uint32_t *tab_smem)
{
    if (bp.pow2) {
        const uint32_t e = a >> 23;
        if (e - 1u < 254u) {
            const int E = (int)e - 127 - bp.qmax_exp + ((a & 0x7FFFFFu) >= tab_smem[e] ? 1 : 0);
            if ((unsigned)(E + 126) <= 253u) return __uint_as_float((uint32_t)(E + 127) << 23);
        }
    } else if (!F32 && bp.qmax_short) {
        const float p = __fmul_rn(__uint_as_float(a), bp.rqmax);
        if (p >= 0x1p-120f && a < 0x7F800000u) {
            float s = __uint_as_float(bf16_rne_hi(p));
            if (bp.has_scale_fmt) s = scale_codebook<false>(bp, s);
            return s > 0.0f ? s : 1.0f;
        }
    }
    return mx_scale_of<F32>(a, bp);
}

// The rounder of the fast path: the same table without the fpN_eXmY NaN-band test (quotients are bounded there).
template <class R>
struct FastOf {
    using type = R;
};
template <bool C, bool M, int REPL>
struct FastOf<TableRounder<C, M, REPL>> {
    using type = TableRounder<C, false, REPL>;
};

// A block may take the fast path when every quotient is finite and below 2^126 and the scale allows the
// reciprocal multiply; sub-2^-120 quotients need no care for formats with bp.fast_ok (see qt_tiny_safe()).
__device__ __forceinline__ bool mx_block_is_fast(uint32_t a, float s, float rs, const BlockParams &bp)
{
    const uint32_t sb = __float_as_uint(s);
    return bp.fast_ok && a < 0x7E800000u && (sb - 0x0D800000u) <= (0x71800000u - 0x0D800000u) &&
           __fmul_rn(__uint_as_float(a), rs) < 0x1p126f;
}
// two bf16 values, each with its own scale
template <class RF>
__device__ __forceinline__ uint32_t mx_word_fast(const RF &round, uint32_t w, float s_lo, float rs_lo, float s_hi,
                                                 float rs_hi)
{
    const uint32_t uq = bf16x2_rne(__fmul_rn(__uint_as_float(w << 16), rs_lo),
                                   __fmul_rn(__uint_as_float(w & 0xFFFF0000u), rs_hi));
    return bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq)), s_lo), __fmul_rn(__uint_as_float(round.hi(uq)), s_hi));
}

__device__ __forceinline__ const uint32_t *stage_pow2_table(const BlockParams &bp, size_t smem_offset)
{
    uint32_t *dst = reinterpret_cast<uint32_t *>(qt_dyn_smem + smem_offset);
    if (bp.pow2) {
        const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.x + threadIdx.y * blockDim.x;
        for (int i = tid; i < QT_POW2_TABLE_WORDS; i += nthreads) dst[i] = bp.pow2_tab[i];
        __syncthreads();
    }
    return dst;
}
constexpr size_t kPow2SmemBytes = QT_POW2_TABLE_WORDS * 4;

}  // namespace
