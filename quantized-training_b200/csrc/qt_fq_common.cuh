// qt_fq_common.cuh -- rounding engines and per-element helpers shared by the fake-quant kernels (qt_fq.cu) and the
// fused producer + fake-quant kernels (qt_fused.cu).  Everything lives in an anonymous namespace: each
// translation unit gets its own copy (the shared-memory table symbol included).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "qt_internal.h"
#include "qt_lut.h"

namespace {

constexpr int kUnroll = 4;  // 16-byte vectors in flight per thread

// ----------------------------------------------------------------------------- rounding engines
// A rounder maps one bf16 value to its rounded value, both as fp32 bits with the low half zero.  Three entry
// points so that the packed bf16 path never pays for an unpack it does not need:
//   operator()(u)   u = fp32 bits (low half zero)
//   lo(w) / hi(w)   the low / high bf16 of a packed 32-bit word

template <int KIND>
struct DirectParams {
    QtRound P;
};
template <int KIND>
struct DirectRounder {
    static constexpr bool kTable = false;
    static constexpr int kThreads = 512, kCtasPerSm = 2, kMinCtas = 2;  // same shape as the table kernels (measured: 256 x 8 was 12 % slower for int)
    static constexpr bool kMxBand = KIND == QTR_FP_MX;
    static constexpr size_t kSmemBytes = 0;
    using Params = DirectParams<KIND>;
    static __device__ __forceinline__ bool tiny_safe(const Params &p) { return p.P.tiny_safe != 0; }
    const QtRound &P;
    __device__ __forceinline__ DirectRounder(const Params &p, const unsigned char *) : P(p.P) {}
    __device__ __forceinline__ uint32_t operator()(uint32_t u) const { return qt_round<KIND>(P, u); }
    __device__ __forceinline__ uint32_t lo(uint32_t w) const { return qt_round<KIND>(P, w << 16); }
    __device__ __forceinline__ uint32_t hi(uint32_t w) const { return qt_round<KIND>(P, w & 0xFFFF0000u); }
};

struct TableParams {
    const QtLutEntry *table;  // global memory, QT_LUT_BYTES
    QtLutCfg cfg;
};
template <bool CLAMP, bool MXBAND, int REPL = QT_LUT_REPLICAS>
struct TableRounder {
    static constexpr bool kTable = true;
    static constexpr int kThreads = 512, kCtasPerSm = 2, kMinCtas = 2;  // 2 x 64 KB of replicated table per SM
    static constexpr bool kMxBand = MXBAND;
    static constexpr int kReplicas = REPL;
    static constexpr size_t kSmemBytes = (size_t)QT_LUT_BYTES * REPL;
    using Params = TableParams;
    static __device__ __forceinline__ bool tiny_safe(const Params &p) { return p.cfg.tiny_safe != 0; }
    const unsigned char *tab;  // shared memory, 8 interleaved replicas (qt_lut.h)
    const uint32_t clamp_bits;
    const uint32_t slot16;
    __device__ __forceinline__ TableRounder(const Params &p, const unsigned char *smem)
        : tab(smem), clamp_bits(p.cfg.clamp_bits), slot16((threadIdx.x & 7u) << 4)
    {
    }
    __device__ __forceinline__ uint32_t go(uint32_t pattern16, uint32_t a) const
    {
        const uint32_t ac = CLAMP ? min(a, clamp_bits) : a;
        return qt_lut_round_smem<MXBAND, REPL>(tab, slot16, pattern16, a, ac);
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t u) const { return go(u >> 16, u & 0x7FFFFFFFu); }
    __device__ __forceinline__ uint32_t lo(uint32_t w) const { return go(w, (w << 16) & 0x7FFFFFFFu); }
    __device__ __forceinline__ uint32_t hi(uint32_t w) const { return go(w >> 16, w & 0x7FFF0000u); }
    // the same for inputs the caller has shown to lie below the fpN_eXmY NaN band (|x| < 0x7F58): no per-element test
    __device__ __forceinline__ uint32_t go_nb(uint32_t pattern16, uint32_t a) const
    {
        const uint32_t ac = CLAMP ? min(a, clamp_bits) : a;
        return qt_lut_round_smem<false, REPL>(tab, slot16, pattern16, a, ac);
    }
    __device__ __forceinline__ uint32_t lo_nb(uint32_t w) const { return go_nb(w, (w << 16) & 0x7FFFFFFFu); }
    __device__ __forceinline__ uint32_t hi_nb(uint32_t w) const { return go_nb(w >> 16, w & 0x7FFF0000u); }
};

extern __shared__ __align__(16) unsigned char qt_dyn_smem[];

// every CTA stages the table once: 8 KB from global (L2-resident after the first CTA) -> 8 replicas
template <class R>
__device__ __forceinline__ const unsigned char *stage_table(const typename R::Params &p)
{
    if constexpr (R::kTable) {
        const float4 *src = reinterpret_cast<const float4 *>(p.table);
        float4 *dst = reinterpret_cast<float4 *>(qt_dyn_smem);
        const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.x + threadIdx.y * blockDim.x;
        for (int i = tid; i < QT_LUT_ENTRIES * R::kReplicas; i += nthreads) dst[i] = src[i / R::kReplicas];
        __syncthreads();
    }
    return qt_dyn_smem;
}

// ----------------------------------------------------------------------------- small helpers

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) { __stcs(p, v); }

// float -> bf16 (RNE, NaN canonical) returned as fp32 bits with the low half zero
__device__ __forceinline__ uint32_t bf16_rne_hi(float f)
{
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f)) << 16;
}
// two floats -> packed bf16x2 (one F2FP instruction)
__device__ __forceinline__ uint32_t bf16x2_rne(float lo, float hi)
{
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&p);
}

// round-to-odd truncation of an fp32 to the bf16 grid: what vmap's index derivation does
// (decomposed.py:151-153); returns fp32 bits with the low half zero
__device__ __forceinline__ uint32_t f32_to_bf16_rto_hi(uint32_t b)
{
    return (b & 0xFFFF0000u) | (((b & 0xFFFFu) != 0u) ? 0x10000u : 0u);
}

// How x / s is evaluated for bf16 tensors.
//   UNIT   s == 1: identity.
//   RECIP  x * (1/s) in fp32, then RNE to bf16.  For bf16 x and s (8-bit significands) the exact quotient
//          is never a bf16 rounding tie and lies at least 2^-17 (relative) away from every tie point, while
//          x * rcp(s) is within 2^-23 of it, so both round to the same bf16 as the reference's
//          bf16(fp32(x / s)).  The argument needs a normal-range quotient: elements whose product is below
//          2^-120 (other than exact zeros) take the true division.
//   EXACT  __fdiv_rn (scale outside [2^-100, 2^100], or not finite).
enum { DIV_UNIT = 0, DIV_RECIP = 1, DIV_EXACT = 2, DIV_RECIP_NOTINY = 3 };

struct ScaleBf16 {
    float s, rs;
};
__device__ __forceinline__ int classify_scale(float s)
{
    const float a = fabsf(s);
    if (s == 1.0f) return DIV_UNIT;
    return (a >= 0x1p-100f && a <= 0x1p100f) ? DIV_RECIP : DIV_EXACT;
}

template <int DIV>
__device__ __forceinline__ float bf16_quotient(uint32_t xh, const ScaleBf16 &sc)
{
    const float x = __uint_as_float(xh);
    if (DIV == DIV_EXACT) return __fdiv_rn(x, sc.s);
    float p = __fmul_rn(x, sc.rs);
    if (fabsf(p) < 0x1p-120f && (xh & 0x7FFFFFFFu) != 0u) p = __fdiv_rn(x, sc.s);
    return p;
}

// one bf16 element held as fp32 bits (low half zero): returns the result in the same form
template <class R, int DIV>
__device__ __forceinline__ uint32_t fq_bf16(const R &round, uint32_t xh, const ScaleBf16 &sc)
{
    if (DIV == DIV_UNIT) return round(xh);
    const uint32_t q = round(bf16_rne_hi(bf16_quotient<DIV>(xh, sc)));
    return bf16_rne_hi(__fmul_rn(__uint_as_float(q), sc.s));  // q * s, rounded to bf16
}

template <class R, bool UNIT>
__device__ __forceinline__ uint32_t fq_f32(const R &round, uint32_t xb, float s)
{
    if (UNIT) return round(f32_to_bf16_rto_hi(xb));
    const float u = __fdiv_rn(__uint_as_float(xb), s);
    const uint32_t q = round(f32_to_bf16_rto_hi(__float_as_uint(u)));
    return __float_as_uint(__fmul_rn(__uint_as_float(q), s));
}

// a 32-bit word holding two bf16 values (amax is taken per vector, see amax_of_vec_bf16)
template <class R, int DIV>
__device__ __forceinline__ uint32_t fq_word_bf16(const R &round, uint32_t w, const ScaleBf16 &sc)
{
    if (DIV == DIV_UNIT) return __byte_perm(round.lo(w), round.hi(w), 0x7632);  // {hi[31:16], lo[31:16]}
    const uint32_t lo = w << 16, hi = w & 0xFFFF0000u;
    // scaled: both conversions are packed (one F2FP per pair each way)
    const uint32_t uq = bf16x2_rne(bf16_quotient<DIV>(lo, sc), bf16_quotient<DIV>(hi, sc));
    const uint32_t qlo = round.lo(uq), qhi = round.hi(uq);
    return bf16x2_rne(__fmul_rn(__uint_as_float(qlo), sc.s), __fmul_rn(__uint_as_float(qhi), sc.s));
}
// Reciprocal-multiply form without a per-element branch: `tiny` collects "some non-zero element has a quotient
// below 2^-120" for the whole vector; the caller then redoes that (rare) vector with the true division.
template <class R>
__device__ __forceinline__ uint32_t fq_word_bf16_recip(const R &round, uint32_t w, const ScaleBf16 &sc, bool &tiny)
{
    const uint32_t lo = w << 16, hi = w & 0xFFFF0000u;
    const float plo = __fmul_rn(__uint_as_float(lo), sc.rs), phi = __fmul_rn(__uint_as_float(hi), sc.rs);
    tiny |= (fabsf(plo) < 0x1p-120f) & ((w & 0x00007FFFu) != 0u);
    tiny |= (fabsf(phi) < 0x1p-120f) & ((w & 0x7FFF0000u) != 0u);
    const uint32_t uq = bf16x2_rne(plo, phi);
    const uint32_t qlo = round.lo(uq), qhi = round.hi(uq);
    return bf16x2_rne(__fmul_rn(__uint_as_float(qlo), sc.s), __fmul_rn(__uint_as_float(qhi), sc.s));
}

// The same without the range test: valid for formats with QtRound::tiny_safe, where a sub-2^-120 quotient only
// has to come out with the right sign and zero-ness (x * rcp(s) underflows to zero exactly when bf16(x / s) does:
// the two differ by a factor within 2^-22 of one, and a quotient of two 8-bit significands is either equal to the
// bf16 underflow boundary 2^-134 -- then the product is exactly on it too -- or at least 2^-8 away from it).
template <class R>
__device__ __forceinline__ uint32_t fq_word_bf16_recip_notiny(const R &round, uint32_t w, const ScaleBf16 &sc)
{
    const uint32_t uq = bf16x2_rne(__fmul_rn(__uint_as_float(w << 16), sc.rs),
                                   __fmul_rn(__uint_as_float(w & 0xFFFF0000u), sc.rs));
    return bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq)), sc.s), __fmul_rn(__uint_as_float(round.hi(uq)), sc.s));
}

// ... and with one scale per half (per-channel along the last axis)
template <class R>
__device__ __forceinline__ uint32_t fq_word_bf16_recip2_notiny(const R &round, uint32_t w, const ScaleBf16 &lo,
                                                               const ScaleBf16 &hi)
{
    const uint32_t uq = bf16x2_rne(__fmul_rn(__uint_as_float(w << 16), lo.rs),
                                   __fmul_rn(__uint_as_float(w & 0xFFFF0000u), hi.rs));
    return bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq)), lo.s), __fmul_rn(__uint_as_float(round.hi(uq)), hi.s));
}

// eight values on that path.  Tables with the fpN_eXmY NaN band test it ONCE per vector on the packed quotients (the
// band is |u| >= 0x7F58, Inf and NaN included) instead of once per element: ~18 instructions per vector less.
template <class R>
__device__ __forceinline__ uint4 fq_vec_bf16_recip_notiny(const R &round, const uint4 &v, const ScaleBf16 &sc)
{
    uint4 r;
    if constexpr (R::kTable && R::kMxBand) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t uq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            uq[k] = bf16x2_rne(__fmul_rn(__uint_as_float(w[k] << 16), sc.rs),
                               __fmul_rn(__uint_as_float(w[k] & 0xFFFF0000u), sc.rs));
        const uint32_t M = 0x7FFF7FFFu;
        const uint32_t m = __vmaxu2(__vmaxu2(uq[0] & M, uq[1] & M), __vmaxu2(uq[2] & M, uq[3] & M));
        uint32_t o[4];
        if ((m >> 16) < 0x7F58u && (m & 0xFFFFu) < 0x7F58u) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = bf16x2_rne(__fmul_rn(__uint_as_float(round.lo_nb(uq[k])), sc.s),
                                  __fmul_rn(__uint_as_float(round.hi_nb(uq[k])), sc.s));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq[k])), sc.s),
                                  __fmul_rn(__uint_as_float(round.hi(uq[k])), sc.s));
        }
        r = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        r.x = fq_word_bf16_recip_notiny<R>(round, v.x, sc);
        r.y = fq_word_bf16_recip_notiny<R>(round, v.y, sc);
        r.z = fq_word_bf16_recip_notiny<R>(round, v.z, sc);
        r.w = fq_word_bf16_recip_notiny<R>(round, v.w, sc);
    }
    return r;
}

__device__ __forceinline__ uint32_t amax_of_vec_f32(uint32_t amax, const uint4 &v)
{
    return max(max(amax, v.x & 0x7FFFFFFFu), max(max(v.y & 0x7FFFFFFFu, v.z & 0x7FFFFFFFu), v.w & 0x7FFFFFFFu));
}
// max |x| over the eight bf16 values of a vector with packed 16-bit maxima (VIMNMX.U16x2): 9 instructions per
// vector instead of 24; |x| bit patterns order like unsigned integers, NaN patterns above Inf.
__device__ __forceinline__ uint32_t amax_of_vec_bf16(uint32_t amax, const uint4 &v)
{
    const uint32_t M = 0x7FFF7FFFu;
    const uint32_t m = __vmaxu2(__vmaxu2(v.x & M, v.y & M), __vmaxu2(v.z & M, v.w & M));
    return max(amax, max(m << 16, m & 0xFFFF0000u));
}

// eight bf16 values by reciprocal multiply; `tiny` is OR-ed with "this vector needs the true division"
template <class R>
__device__ __forceinline__ uint4 fq_vec_bf16_recip_fast(const R &round, const uint4 &v, const ScaleBf16 &sc, bool &tiny)
{
    uint4 r;
    r.x = fq_word_bf16_recip<R>(round, v.x, sc, tiny);
    r.y = fq_word_bf16_recip<R>(round, v.y, sc, tiny);
    r.z = fq_word_bf16_recip<R>(round, v.z, sc, tiny);
    r.w = fq_word_bf16_recip<R>(round, v.w, sc, tiny);
    return r;
}

template <class R, bool F32, int DIV, bool AMAX>
__device__ __forceinline__ uint4 fq_vec(const R &round, uint4 v, const ScaleBf16 &sc, uint32_t &amax)
{
    uint4 r;
    if (F32) {
        if (AMAX) amax = amax_of_vec_f32(amax, v);
        r.x = fq_f32<R, DIV == DIV_UNIT>(round, v.x, sc.s);
        r.y = fq_f32<R, DIV == DIV_UNIT>(round, v.y, sc.s);
        r.z = fq_f32<R, DIV == DIV_UNIT>(round, v.z, sc.s);
        r.w = fq_f32<R, DIV == DIV_UNIT>(round, v.w, sc.s);
    } else {
        if (AMAX) amax = amax_of_vec_bf16(amax, v);
        if (DIV == DIV_RECIP_NOTINY) {
            r = fq_vec_bf16_recip_notiny<R>(round, v, sc);
        } else if (DIV == DIV_RECIP) {
            bool tiny = false;
            r = fq_vec_bf16_recip_fast<R>(round, v, sc, tiny);
            if (tiny) {  // sub-2^-120 quotients: the reference's fp32 division rounds in the denormal range
                r.x = fq_word_bf16<R, DIV_EXACT>(round, v.x, sc);
                r.y = fq_word_bf16<R, DIV_EXACT>(round, v.y, sc);
                r.z = fq_word_bf16<R, DIV_EXACT>(round, v.z, sc);
                r.w = fq_word_bf16<R, DIV_EXACT>(round, v.w, sc);
            }
        } else {
            r.x = fq_word_bf16<R, DIV>(round, v.x, sc);
            r.y = fq_word_bf16<R, DIV>(round, v.y, sc);
            r.z = fq_word_bf16<R, DIV>(round, v.z, sc);
            r.w = fq_word_bf16<R, DIV>(round, v.w, sc);
        }
    }
    return r;
}

// scale.to(x.dtype): bf16 inputs see the scale rounded to bf16 (fake_quantize.py:245)
template <bool F32>
__device__ __forceinline__ ScaleBf16 load_scale(const float *scale, size_t c)
{
    ScaleBf16 sc;
    const float s = scale[c];
    sc.s = F32 ? s : __uint_as_float(bf16_rne_hi(s));
    sc.rs = __frcp_rn(sc.s);
    return sc;
}

// block-wide max of |x| bit patterns, then ONE atomicMax.  Non-negative floats order like
// unsigned ints and NaN patterns sit above Inf, so NaN propagates exactly like torch.amax.
__device__ __forceinline__ void block_amax_commit(uint32_t amax, float *amax_out)
{
    __shared__ uint32_t warp_max[32];
    amax = __reduce_max_sync(0xFFFFFFFFu, amax);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) warp_max[warp] = amax;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < (int)(blockDim.x >> 5) ? warp_max[lane] : 0u;
        v = __reduce_max_sync(0xFFFFFFFFu, v);
        if (lane == 0 && v != 0u) atomicMax(reinterpret_cast<unsigned int *>(amax_out), v);
    }
}

// ----------------------------------------------------------------------------- launch plumbing

int g_num_sms = 0;

int cuda_fail(cudaError_t e, const char *what)
{
    qt_set_error("%s: %s", what, cudaGetErrorString(e));
    return QT_ERR_CUDA;
}

int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        g_num_sms = n;
    }
    return g_num_sms;
}

inline unsigned grid_for(size_t work_items, int ctas_per_sm)
{
    size_t cap = (size_t)num_sms() * ctas_per_sm;
    if (cap == 0) cap = 148u * ctas_per_sm;
    return (unsigned)(work_items < cap ? (work_items ? work_items : 1) : cap);
}

// kernels that stage the replicated table need 64 KB of dynamic shared memory: opt in once per device
template <auto kernel>  // one flag table per kernel instantiation
void allow_smem(size_t bytes)
{
    if (bytes <= 48 * 1024) return;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (dev >= 0 && dev < 64) done[dev] = true;
}

int no_device()
{
    cudaError_t e = cudaGetLastError();
    return cuda_fail(e == cudaSuccess ? cudaErrorNoDevice : e,
                     "qt_b200: no usable CUDA device (there is no CPU fallback)");
}

// Picks the rounding engine for a format: the binade-constant table when the caller supplied one and the format
// has it, else the direct bitwise logic.  `fn` is called with a rounder tag type and its kernel parameters.
template <class R>
struct RounderTag {
    using type = R;
};
template <class Fn>
int dispatch_direct_small(const QtRound &P, Fn &&fn)
{
    if (P.kind == QTR_INT) {
        DirectParams<QTR_INT> p;
        p.P = P;
        fn(RounderTag<DirectRounder<QTR_INT>>{}, p);
    } else {
        DirectParams<QTR_IDENTITY> p;
        p.P = P;
        fn(RounderTag<DirectRounder<QTR_IDENTITY>>{}, p);
    }
    return QT_OK;
}
template <class Fn>
int dispatch_rounder(const QtRound &P, const void *lut, Fn &&fn)
{
    TableParams tp;
    if (lut && qt_lut_config(P, &tp.cfg) == QT_OK) {
        if (reinterpret_cast<uintptr_t>(lut) & 15u) {
            qt_set_error("qt_b200: lut must be 16-byte aligned");
            return QT_ERR_UNALIGNED;
        }
        tp.table = static_cast<const QtLutEntry *>(lut);
        if (tp.cfg.mx_band)
            fn(RounderTag<TableRounder<true, true>>{}, tp);
        else if (tp.cfg.clamp_bits != 0x7FFFFFFFu)
            fn(RounderTag<TableRounder<true, false>>{}, tp);
        else
            fn(RounderTag<TableRounder<false, false>>{}, tp);
        return QT_OK;
    }
    switch (P.kind) {
    case QTR_IDENTITY: { DirectParams<QTR_IDENTITY> p; p.P = P; fn(RounderTag<DirectRounder<QTR_IDENTITY>>{}, p); break; }
    case QTR_INT: { DirectParams<QTR_INT> p; p.P = P; fn(RounderTag<DirectRounder<QTR_INT>>{}, p); break; }
    case QTR_FP_CUSTOM: { DirectParams<QTR_FP_CUSTOM> p; p.P = P; fn(RounderTag<DirectRounder<QTR_FP_CUSTOM>>{}, p); break; }
    case QTR_FP_MX: { DirectParams<QTR_FP_MX> p; p.P = P; fn(RounderTag<DirectRounder<QTR_FP_MX>>{}, p); break; }
    case QTR_POSIT: { DirectParams<QTR_POSIT> p; p.P = P; fn(RounderTag<DirectRounder<QTR_POSIT>>{}, p); break; }
    default: qt_set_error("bad format kind %d", P.kind); return QT_ERR_INVALID_ARGUMENT;
    }
    return QT_OK;
}

}  // namespace
