// qt_fused.cu -- the ops BETWEEN the GEMMs of a quantized transformer block, each fused with the fake-quant step
// that follows it (and, for the hooks that precede it, the one before), one memory pass per op:
//
//   qt_softmax_fq    scores -> [fq] -> * alpha -> + mask -> [fq] -> softmax -> [fq]      (attention probabilities)
//   qt_norm_fq       x -> [fq] -> RMSNorm / LayerNorm -> [fq]                             (block inputs)
//   qt_act_mul_fq    fq(act(gate) * up), fq(act(x)), strided fq                           (MLP)
//   qt_rope_fq       fq(x * cos + rotate_half(x) * sin) for q and k in one launch         (Llama)
//   qt_fq_transpose  v [B, S, H, D] -> fq -> [B, H, D, S]   (the K-major operand of probs x V)
//
// In the reference each of these is a chain of bf16 ATen ops (HF modeling code) followed by the fake-quant hook of
// the consuming module (quantize.py:116-150): 5-15 launches and as many round trips through HBM.  Here the bf16
// roundings of that chain are reproduced in registers -- every intermediate the reference materialises in bf16 is
// rounded to bf16 at the same point -- so the only differences are the order of the row reductions and the last-ulp
// behaviour of exp / rsqrt.  The fake-quant step is the same rounding engine as qt_fq_forward (bit-exact formats);
// per-tensor scales are read from device memory (NULL = bare spec, scale 1).  Observers (amax) are not fused: a
// module with a live observer keeps its own qt_fq_forward pass.
#include <cuda_fp8.h>
#include <math.h>

#include "qt_fq_common.cuh"
#include "qt_launch.cuh"

namespace {

constexpr int FQ_PRE = 1;   // fake-quantize the op's input  (hook of the op's own group: scaling / layernorm)
constexpr int FQ_MID = 2;   // softmax only: fake-quantize scores * alpha + mask (the "activation" hook of nn.Softmax)
constexpr int FQ_CAUSAL = 16;  // qt_softmax_fq: QT_SOFTMAX_CAUSAL
constexpr int FQ_POST = 4;  // fake-quantize the op's output (input hook of the consuming GEMM)
constexpr int FQ_RES_A = 32;  // qt_add_norm_fq: fake-quantize x before the residual add (AddFunctional input 0 hook)
constexpr int FQ_RES_B = 64;  // qt_add_norm_fq: fake-quantize the residual operand (AddFunctional input 1 hook)

// Table rounder of the fused kernels: single-replica 8 KB table (staging 64 KB per CTA would dominate these small
// launches), clamp / NaN-band switches read at run time so that ONE instantiation serves every fp / posit format.
struct TableRounderDyn {
    static constexpr bool kTable = true;
    static constexpr int kReplicas = 1;
    static constexpr bool kMxBand = true;
    static constexpr size_t kSmemBytes = QT_LUT_BYTES;
    using Params = TableParams;
    const unsigned char *tab;
    const uint32_t clamp_bits, mx_band;
    const bool plain;  // no clamp, no NaN band (posit formats): the two-FMA core alone
    __device__ __forceinline__ TableRounderDyn(const Params &p, const unsigned char *smem)
        : tab(smem), clamp_bits(p.cfg.clamp_bits), mx_band(p.cfg.mx_band),
          plain(p.cfg.clamp_bits == 0x7FFFFFFFu && p.cfg.mx_band == 0u)
    {
    }
    template <bool PLAIN>
    __device__ __forceinline__ uint32_t apply(uint32_t u) const
    {
        const uint32_t a = u & 0x7FFFFFFFu;
        uint32_t q = qt_lut_round_smem<false, 1>(tab, 0u, u >> 16, a, PLAIN ? a : min(a, clamp_bits));
        if (!PLAIN && mx_band && a >= 0x7F580000u && a != 0x7F800000u) q = QT_NAN_BITS;
        return q;
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t u) const { return apply<false>(u); }
    __device__ __forceinline__ uint32_t lo(uint32_t w) const { return (*this)(w << 16); }
    __device__ __forceinline__ uint32_t hi(uint32_t w) const { return (*this)(w & 0xFFFF0000u); }
    // eight values; the format switches are uniform over the launch, so they are tested once per vector
    __device__ __forceinline__ void round8(float (&f)[8]) const
    {
        if (plain) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(apply<true>(__float_as_uint(f[k])));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(apply<false>(__float_as_uint(f[k])));
        }
    }
};
template <class R>
__device__ __forceinline__ void rounder8(const R &round, float (&f)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(round(__float_as_uint(f[k])));
}
__device__ __forceinline__ void rounder8(const TableRounderDyn &round, float (&f)[8]) { round.round8(f); }

struct FqPoint {
    ScaleBf16 sc;
    int mode;
};
__device__ __forceinline__ FqPoint load_point(const float *scale)
{
    FqPoint p;
    p.sc.s = p.sc.rs = 1.0f;
    if (scale) p.sc = load_scale<false>(scale, 0);
    p.mode = classify_scale(p.sc.s);
    return p;
}
// ex2.approx.ftz alone (MUFU.EX2): arguments are <= 0 here, results below 2^-126 flush to zero, which is what the
// following rounding to bf16 / to the format would make of them anyway (__expf adds a range fix-up of 3 instructions)
__device__ __forceinline__ float exp2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Rounding to the bf16 grid, eight values at a time through the PACKED conversion (F2FP.BF16.PACK_AB, ALU pipe, one
// instruction per pair + two to unpack).  The scalar cvt.rn.bf16.f32 is an XU-pipe instruction (16 lanes / clock / SM,
// shared with MUFU.EX2): with three roundings and one exp per element the softmax kernel was XU-bound (ncu: XU 47 %,
// issue 54 %, DRAM 14 %).
__device__ __forceinline__ void round8(float (&f)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t p = bf16x2_rne(f[2 * k], f[2 * k + 1]);
        f[2 * k] = __uint_as_float(p << 16);
        f[2 * k + 1] = __uint_as_float(p & 0xFFFF0000u);
    }
}

// Eight bf16 values (as floats) through quantize-dequantize.  SCALED = false (bare specs: every BASELINE config):
// the rounding alone.  SCALED = true: the frozen per-tensor scale, divide / round / multiply as qt_fq_forward does;
// the mode is uniform over the launch, so the three-way branch is taken per vector, not per element.
template <class R, bool SCALED>
__device__ __forceinline__ void fq8(const R &round, float (&f)[8], const FqPoint &p)
{
    if (!SCALED || p.mode == DIV_UNIT) {
        rounder8(round, f);
    } else if (p.mode == DIV_RECIP) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(fq_bf16<R, DIV_RECIP>(round, __float_as_uint(f[k]), p.sc));
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(fq_bf16<R, DIV_EXACT>(round, __float_as_uint(f[k]), p.sc));
    }
}
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8])
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(w[k] << 16);
        f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
    }
}
// values already on the bf16 grid -> packed
__device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    uint4 o;
    o.x = __byte_perm(__float_as_uint(f[0]), __float_as_uint(f[1]), 0x7632);
    o.y = __byte_perm(__float_as_uint(f[2]), __float_as_uint(f[3]), 0x7632);
    o.z = __byte_perm(__float_as_uint(f[4]), __float_as_uint(f[5]), 0x7632);
    o.w = __byte_perm(__float_as_uint(f[6]), __float_as_uint(f[7]), 0x7632);
    return o;
}
// fp32 values -> RNE to bf16 -> packed (one F2FP per pair)
__device__ __forceinline__ uint4 round_pack8(const float (&f)[8])
{
    return make_uint4(bf16x2_rne(f[0], f[1]), bf16x2_rne(f[2], f[3]), bf16x2_rne(f[4], f[5]), bf16x2_rne(f[6], f[7]));
}

// Output forms.  OUT_BF16: the fake-quantized values.  OUT_E4M3 / OUT_E5M2: their one-byte OCP fp8 encodings (8 bytes
// per vector) for the FP8 tensor-core GEMM -- only after an UNSCALED fake-quant step of that very format, so that
// decode(code) == value exactly (the launchers enforce it).  +-Inf (which only the fpN_eXmY flavour lets through)
// gets the format's Inf / NaN code instead of the saturated maximum.
enum { OUT_BF16 = 0, OUT_E4M3 = 1, OUT_E5M2 = 2, OUT_INF_POSSIBLE = 4 /* flag, set by the launcher for fpN_eXmY */ };
template <bool INF>
__device__ __forceinline__ uint32_t fp8x2_dyn(float lo, float hi, int out_type)
{
    const bool e5m2 = (out_type & 3) == OUT_E5M2;
    uint32_t c = e5m2 ? (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E5M2)
                      : (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3);
    if (INF) {
        const uint32_t inf_code = e5m2 ? 0x7Cu : 0x7Fu;
        const uint32_t bl = __float_as_uint(lo), bh = __float_as_uint(hi);
        if ((bl & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0xFF00u) | ((bl >> 24) & 0x80u) | inf_code;
        if ((bh & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0x00FFu) | ((((bh >> 24) & 0x80u) | inf_code) << 8);
    }
    return c;
}
__device__ __forceinline__ uint2 codes8(const float (&f)[8], int out_type)
{
    uint2 o;
    if (out_type & OUT_INF_POSSIBLE) {
        o.x = fp8x2_dyn<true>(f[0], f[1], out_type) | (fp8x2_dyn<true>(f[2], f[3], out_type) << 16);
        o.y = fp8x2_dyn<true>(f[4], f[5], out_type) | (fp8x2_dyn<true>(f[6], f[7], out_type) << 16);
    } else {
        o.x = fp8x2_dyn<false>(f[0], f[1], out_type) | (fp8x2_dyn<false>(f[2], f[3], out_type) << 16);
        o.y = fp8x2_dyn<false>(f[4], f[5], out_type) | (fp8x2_dyn<false>(f[6], f[7], out_type) << 16);
    }
    return o;
}
// vector `idx` (8 elements) of an output whose base is `base`: 16 bytes of bf16 or 8 bytes of codes
__device__ __forceinline__ void store8(void *base, size_t idx, const float (&f)[8], int out_type)
{
    if (out_type == OUT_BF16) {
        __stcs(static_cast<uint4 *>(base) + idx, pack8(f));
    } else {
        __stcs(static_cast<uint2 *>(base) + idx, codes8(f, out_type));
    }
}

// Launch shapes.  Row kernels: 256-thread CTAs; a row is owned by TPR = 32 threads (rows up to 1024 elements: eight
// rows per CTA, warp-shuffle reductions only) or by 128 (longer rows: two rows per CTA), VPL 16-byte vectors per
// thread, i.e. 64 bytes in flight per thread -- these launches are latency-bound unless every SM keeps ~40 KB of
// loads in flight.  Elementwise kernels: 256-thread CTAs.
constexpr int ROW_THREADS = 256, ROW_MIN_CTAS = 4;
constexpr int EW_THREADS = 256, EW_MIN_CTAS = 4;
// Measured (scripts/fused_micro.py, Llama window): 4 resident CTAs (64 registers) instead of 3 (85) take the row and
// elementwise kernels from 33 % to 50 % occupancy -- rmsnorm 9.4 -> 7.6 us, rope 16.7 -> 13.7, transpose 10.1 -> 8.7,
// softmax 55 -> 51 -- except the gated activation product, which spills and is better off with 3.
constexpr int ACT_MIN_CTAS = 3;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// reductions over the TPR threads that own a row (TPR == 32: one warp; TPR == 128: four warps through smem)
template <int TPR, bool MAX>
__device__ __forceinline__ float row_reduce(float v)
{
    v = MAX ? warp_max(v) : warp_sum(v);
    if constexpr (TPR == 32) {
        return v;
    } else {
        static_assert(TPR == 128, "a long row is owned by four warps");
        __shared__ float part[ROW_THREADS / 32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) part[warp] = v;
        __syncthreads();
        const int g = (warp >> 2) << 2;  // first warp of this row's group
        const float a = part[g], b = part[g + 1], c = part[g + 2], d = part[g + 3];
        __syncthreads();  // the next reduction may overwrite part[]
        return MAX ? fmaxf(fmaxf(a, b), fmaxf(c, d)) : (a + b) + (c + d);
    }
}

// ----------------------------------------------------------------------------- softmax
// TPR threads per row of `cols` (<= TPR * VPL * 8) scores; the row lives in registers between the passes.
// Reference chain (modules/quantizable/modeling_bert.py:118-158, modeling_llama.py:228-246), all bf16 tensors:
//   s = qk_matmul(q, k^T); [s = fq(s)]; s = s * scaling; s = s + mask; [s = fq(s)]; p = softmax(s, -1); [p = fq(p)]
// Which steps exist is a run-time mask tested once per 8-element vector (a per-element test multiplies the
// unrolled code by the number of combinations and thrashes the instruction cache).
// SIMPLE: the step set of a gemm-only model (Llama perplexity, BASELINE configs[1]) -- no hook on the scores and no mid
// point: those branches and the prefetched mask registers are compiled out
// and the kernel keeps its row in registers without spills (the general kernel is capped at 64 registers).
template <class R, bool SCALED, int TPR, int VPL, bool SIMPLE = false>
__global__ void __launch_bounds__(ROW_THREADS, SIMPLE ? 3 : ROW_MIN_CTAS)
softmax_fq_kernel(const uint4 *__restrict__ scores, void *__restrict__ probs, int out_type, size_t rows, int cols, float alpha,
                  int has_alpha, const uint4 *__restrict__ mask, size_t rows_per_batch, size_t mask_rows,
                  size_t mask_batch_stride_vec, int flags, const __grid_constant__ typename R::Params params,
                  const float *__restrict__ scale_pre, const float *__restrict__ scale_mid,
                  const float *__restrict__ scale_post, const int32_t *__restrict__ causal_flag)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the table is constant data; everything below reads what the predecessor wrote
    griddep_launch_dependents();
    const FqPoint pre = load_point(scale_pre), mid = load_point(scale_mid), post = load_point(scale_post);
    if (causal_flag && *causal_flag == 0) flags &= ~FQ_CAUSAL;  // decided on the device (graph-safe)
    constexpr int RPC = ROW_THREADS / TPR;  // rows per CTA pass
    const int lane = threadIdx.x % TPR;
    const int nvec = cols >> 3;
    const size_t row_stride = (size_t)gridDim.x * RPC;
    // every thread of a CTA runs the same number of passes (the long-row reductions contain block barriers)
    for (size_t row0 = (size_t)blockIdx.x * RPC; row0 < rows; row0 += row_stride) {
        const size_t row = row0 + threadIdx.x / TPR;
        const bool live = row < rows;
        // QT_SOFTMAX_CAUSAL: columns past the row's own position are masked by contract -- not read, probability 0;
        // columns past the row tile's diagonal block are not even written (the P x V product stops there)
        const int r_seq = (flags & FQ_CAUSAL) ? (int)(row % mask_rows) : 0x7FFFFFF0;
        const int wr_end = (flags & FQ_CAUSAL) ? ((r_seq >> 7) + 1) << 7 : cols;
        const uint4 *srow = scores + (live ? row : 0) * nvec;
        const uint4 *mrow = nullptr;
        if (mask && live) mrow = mask + (row / rows_per_batch) * mask_batch_stride_vec + (row % mask_rows) * (size_t)nvec;
        uint4 raw[VPL], mraw[SIMPLE ? 1 : VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {  // all loads first: VPL (x2 with a mask) 16-byte requests in flight per thread
            const int i = lane + TPR * j;
            const bool in = i < nvec && live && i * 8 <= r_seq;
            raw[j] = in ? __ldcs(srow + i) : make_uint4(0u, 0u, 0u, 0u);
            // causal by contract: the mask is known (0 up to the diagonal, finfo.min above it) and is not read
            if constexpr (!SIMPLE)
                mraw[j] = (in && mrow && !(flags & FQ_CAUSAL)) ? __ldg(mrow + i) : make_uint4(0u, 0u, 0u, 0u);
        }
        float f[VPL][8];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int i = lane + TPR * j;
            if (i < nvec && live && i * 8 <= r_seq) {
                unpack8(raw[j], f[j]);
                if (!SIMPLE && (flags & FQ_PRE)) fq8<R, SCALED>(round, f[j], pre);
                if (has_alpha) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[j][k] *= alpha;
                    round8(f[j]);
                }
                if (flags & FQ_CAUSAL) {
                    // s + 0 == s up to the diagonal; bf16(s + finfo.min) is so far below the row maximum that its
                    // exp is exactly 0, like -inf: only the vector holding the diagonal needs any work
                    if (i * 8 + 7 > r_seq) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (i * 8 + k > r_seq) f[j][k] = -INFINITY;
                    }
                } else if (mrow) {
                    float m8[8];
                    // SIMPLE: a mask that turned out not to be causal (device flag) is read here, unprefetched
                    unpack8(SIMPLE ? __ldg(mrow + i) : mraw[SIMPLE ? 0 : j], m8);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[j][k] += m8[k];
                    round8(f[j]);
                }
                if (!SIMPLE && (flags & FQ_MID)) fq8<R, SCALED>(round, f[j], mid);
#pragma unroll
                for (int k = 0; k < 8; ++k) mx = fmaxf(mx, f[j][k]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[j][k] = -INFINITY;
            }
        }
        mx = row_reduce<TPR, true>(mx);
        float sum = 0.0f;
        // A vector whose eight inputs are all below mx - 128 contributes exp() == 0 exactly (the fp32 exp underflows
        // below -103.97): masked positions (s + finfo.min) take this path -- half of a causal row -- and skip the
        // exp, the normalisation and the fake quant, whose result for 0 is the same 0 for every format.
        const float dead = mx - 128.0f;
        unsigned alive = 0u;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            float vmax = f[j][0];
#pragma unroll
            for (int k = 1; k < 8; ++k) vmax = fmaxf(vmax, f[j][k]);
            if (vmax > dead || !(mx == mx) || mx == -INFINITY) {  // NaN / all-masked rows: keep torch's NaN semantics
                alive |= 1u << j;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // s and mx are bf16 values: the difference is exact; exp through ex2.approx (relative error
                    // ~2^-21 for the arguments that survive the following rounding to bf16)
                    f[j][k] = exp2_approx((f[j][k] - mx) * 1.4426950408889634f);
                    sum += f[j][k];
                }
            }
        }
        sum = row_reduce<TPR, false>(sum);
        const float inv = __frcp_rn(sum);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int i = lane + TPR * j;
            if (i < nvec && live && i * 8 < wr_end) {
                if (alive & (1u << j)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[j][k] *= inv;
                    round8(f[j]);
                    if (flags & FQ_POST) fq8<R, SCALED>(round, f[j], post);
                    store8(probs, row * nvec + i, f[j], out_type);
                } else if (out_type == OUT_BF16) {
                    __stcs(static_cast<uint4 *>(probs) + row * nvec + i, make_uint4(0u, 0u, 0u, 0u));
                } else {
                    __stcs(static_cast<uint2 *>(probs) + row * nvec + i, make_uint2(0u, 0u));
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- RMSNorm / LayerNorm
// TPR threads per row of `cols` (<= TPR * VPL * 8).  kind 0: LlamaRMSNorm (HF modeling_llama.py): n = bf16(x *
// rsqrt(mean(x^2) + eps)) computed in fp32, y = bf16(weight * n).  kind 1: nn.LayerNorm: y = bf16((x - mean) * rstd * w
// + b), fp32 inside.
template <class R, bool SCALED, int TPR, int VPL>
__global__ void __launch_bounds__(ROW_THREADS, ROW_MIN_CTAS)
norm_fq_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ res, void *__restrict__ y,
               uint4 *__restrict__ y_raw, int out_type, size_t rows, int cols, int kind,
               const uint4 *__restrict__ weight, const uint4 *__restrict__ bias, float eps, int flags,
               const __grid_constant__ typename R::Params params, const float *__restrict__ scale_pre,
               const float *__restrict__ scale_post)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the table is constant data; everything below reads what the predecessor wrote
    griddep_launch_dependents();
    const FqPoint pre = load_point(scale_pre), post = load_point(scale_post);
    constexpr int RPC = ROW_THREADS / TPR;
    const int lane = threadIdx.x % TPR;
    const int nvec = cols >> 3;
    const float inv_n = 1.0f / (float)cols;
    const size_t row_stride = (size_t)gridDim.x * RPC;
    for (size_t row0 = (size_t)blockIdx.x * RPC; row0 < rows; row0 += row_stride) {
        const size_t row = row0 + threadIdx.x / TPR;
        const bool live = row < rows;
        uint4 v[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int i = lane + TPR * j;
            v[j] = (i < nvec && live) ? __ldcs(x + row * nvec + i) : make_uint4(0u, 0u, 0u, 0u);
        }
        if (res) {  // the residual add in front of the norm (AddFunctional), with the hooks on its two inputs
            FqPoint bare;
            bare.sc.s = bare.sc.rs = 1.0f;
            bare.mode = DIV_UNIT;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                const int i = lane + TPR * j;
                if (i < nvec && live) {
                    float a[8], b[8];
                    unpack8(v[j], a);
                    unpack8(__ldcs(res + row * nvec + i), b);
                    if (flags & FQ_RES_A) fq8<R, false>(round, a, bare);
                    if (flags & FQ_RES_B) fq8<R, false>(round, b, bare);
#pragma unroll
                    for (int k = 0; k < 8; ++k) a[k] += b[k];
                    v[j] = round_pack8(a);
                }
            }
        }
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            float f[8];
            unpack8(v[j], f);
            if (flags & FQ_PRE) {
                fq8<R, SCALED>(round, f, pre);  // zeros of the padding stay zeros for every format
                v[j] = pack8(f);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s1 += f[k];
                s2 += f[k] * f[k];
            }
        }
        float mean = 0.0f, rstd;
        if (kind == 0) {
            rstd = rsqrtf(row_reduce<TPR, false>(s2) * inv_n + eps);
        } else {
            mean = row_reduce<TPR, false>(s1) * inv_n;
            float ss = 0.0f;  // second pass over the registers: sum (x - mean)^2 over the row's real elements
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                if (lane + TPR * j < nvec) {
                    float f[8];
                    unpack8(v[j], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) ss += (f[k] - mean) * (f[k] - mean);
                }
            }
            rstd = rsqrtf(row_reduce<TPR, false>(ss) * inv_n + eps);
        }
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const int i = lane + TPR * j;
            if (i < nvec && live) {
                float f[8], w[8];
                unpack8(v[j], f);
                if (kind != 3) unpack8(__ldg(weight + i), w);
                if (kind == 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] *= rstd;
                    round8(f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] *= w[k];
                    round8(f);
                } else if (kind == 3) {  // no norm at all: the (hooked) residual add alone
                } else if (kind == 2) {  // MobileBERT NoNorm: input * weight + bias, two bf16 ops, no statistics
                    float b[8];
                    unpack8(__ldg(bias + i), b);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] *= w[k];
                    round8(f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] += b[k];
                    round8(f);
                } else {
                    float b[8];
                    if (bias) unpack8(__ldg(bias + i), b);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = (f[k] - mean) * rstd * w[k] + (bias ? b[k] : 0.0f);
                    round8(f);
                }
                if (y_raw) __stcs(y_raw + row * nvec + i, pack8(f));  // the normalised values before the output step
                if (flags & FQ_POST) fq8<R, SCALED>(round, f, post);
                store8(y, row * nvec + i, f, out_type);
            }
        }
    }
}

// ----------------------------------------------------------------------------- activation (x up) + fq, strided rows
// out[r, c] = fq( bf16( bf16(act(gate[r, c])) * up[r, c] ) )   (up == NULL: no product).  ACT: compile time.
// HF LlamaMLP: down_proj(act_fn(gate_proj(x)) * up_proj(x)); each op rounds to bf16.
enum { FACT_NONE = 0, FACT_RELU = 1, FACT_GELU = 2, FACT_SILU = 3 };
template <class R, bool SCALED, int ACT>
__global__ void __launch_bounds__(EW_THREADS, ACT_MIN_CTAS)
act_mul_fq_kernel(const uint4 *__restrict__ gate, const uint4 *__restrict__ up, void *__restrict__ out, int out_type,
                  size_t rows,
                  int vec_per_row, size_t ld_gate_vec, size_t ld_up_vec, size_t ld_out_vec, int flags,
                  const __grid_constant__ typename R::Params params, const float *__restrict__ scale_post)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the table is constant data; everything below reads what the predecessor wrote
    griddep_launch_dependents();
    const FqPoint post = load_point(scale_post);
    const size_t total = rows * (size_t)vec_per_row;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int U = 2;  // independent vectors in flight per thread
    for (size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
        uint4 gv[U], uv[U];
        size_t off_out[U];
        bool live[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const size_t t = t0 + (size_t)i * stride;
            live[i] = t < total;
            const size_t r = live[i] ? t / vec_per_row : 0, c = live[i] ? t - r * vec_per_row : 0;
            off_out[i] = r * ld_out_vec + c;
            gv[i] = live[i] ? __ldcs(gate + r * ld_gate_vec + c) : make_uint4(0u, 0u, 0u, 0u);
            uv[i] = (live[i] && up) ? __ldcs(up + r * ld_up_vec + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
            float g[8];
            unpack8(gv[i], g);
            if (flags & FQ_PRE) {  // the activation module's own input hook (bare spec)
                FqPoint bare;
                bare.sc.s = bare.sc.rs = 1.0f;
                bare.mode = DIV_UNIT;
                fq8<R, false>(round, g, bare);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float a = g[k];
                if (ACT == FACT_SILU)
                    a = __fdividef(a, 1.0f + __expf(-a));
                else if (ACT == FACT_GELU)
                    a = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));
                else if (ACT == FACT_RELU)
                    a = fmaxf(a, 0.0f);
                g[k] = a;
            }
            if (ACT == FACT_SILU || ACT == FACT_GELU) round8(g);
            if (up) {
                float u[8];
                unpack8(uv[i], u);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] *= u[k];
                round8(g);
            }
            if (flags & FQ_POST) fq8<R, SCALED>(round, g, post);
            if (live[i]) store8(out, off_out[i], g, out_type);
        }
    }
}

// ----------------------------------------------------------------------------- LoRA merged weight + fq
// qat.LoraLinear (reference modules/qat/lora.py:44-52):  W' = fq( W + transpose(fq(B) @ fq(A)) * scaling ), as one pass
// over W.  The reference's chain is clone -> fq(A) -> fq(B) -> matmul (bf16 out, fp32 accumulate) -> * scaling (bf16)
// -> += (bf16) -> fq: the three bf16 roundings are reproduced in registers.  A [r, K] is staged once per CTA in shared
// memory (fake-quantized on the way in when FQ_PRE is set); B [N, r] is read per output row.
template <class R, bool SCALED>
__global__ void __launch_bounds__(EW_THREADS, EW_MIN_CTAS)
lora_merge_fq_kernel(const uint4 *__restrict__ W, const uint4 *__restrict__ A, const uint16_t *__restrict__ B,
                     uint4 *__restrict__ out, size_t N, int k_vecs, int r, float scaling, int flags,
                     const __grid_constant__ typename R::Params params, const float *__restrict__ scale_post)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();
    griddep_launch_dependents();
    const FqPoint post = load_point(scale_post);
    FqPoint bare;
    bare.sc.s = bare.sc.rs = 1.0f;
    bare.mode = DIV_UNIT;
    uint4 *sA = reinterpret_cast<uint4 *>(qt_dyn_smem + R::kSmemBytes);
    for (int i = threadIdx.x; i < r * k_vecs; i += blockDim.x) {
        uint4 v = __ldg(A + i);
        if (flags & FQ_PRE) {
            float f[8];
            unpack8(v, f);
            fq8<R, false>(round, f, bare);
            v = pack8(f);
        }
        sA[i] = v;
    }
    __syncthreads();
    const size_t total = N * (size_t)k_vecs;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t n = t / k_vecs;
        const int kv = (int)(t - n * k_vecs);
        const uint4 wv = __ldcs(W + t);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j0 = 0; j0 < r; j0 += 8) {
            float b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                b[j] = (j0 + j < r) ? __uint_as_float((uint32_t)__ldg(B + n * r + j0 + j) << 16) : 0.0f;
            if (flags & FQ_PRE) fq8<R, false>(round, b, bare);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j0 + j < r) {
                    float a[8];
                    unpack8(sA[(j0 + j) * k_vecs + kv], a);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(b[j], a[k], acc[k]);
                }
            }
        }
        round8(acc);  // the matmul's bf16 output
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= scaling;
        round8(acc);  // * scaling
        float w[8];
        unpack8(wv, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += w[k];
        round8(acc);  // += into the cloned weight
        if (flags & FQ_POST) fq8<R, SCALED>(round, acc, post);
        __stcs(out + t, pack8(acc));
    }
}

// ----------------------------------------------------------------------------- rotary embedding + fq
// HF apply_rotary_pos_emb: x_embed = (x * cos) + (rotate_half(x) * sin), three bf16 ops.  x: [tokens, heads, D] with a
// token stride (projection output [B*S, heads*D], possibly a slice of a fused QKV buffer); cos/sin: [cos_rows, D],
// token t uses row t % cos_rows.  Up to two tensors (q and k) per launch.  A thread owns columns [8c, 8c+8) and their
// partners [8c + D/2, ...).
struct RopeTensor {
    const uint4 *x;
    void *y;
    size_t ld_x_vec, ld_y_vec;  // token strides, 16-byte vectors
    int heads;
};
template <class R, bool SCALED>
__global__ void __launch_bounds__(EW_THREADS, EW_MIN_CTAS)
rope_fq_kernel(RopeTensor t0, RopeTensor t1, int out_type, size_t tokens, int head_dim, const uint4 *__restrict__ cos_t,
               const uint4 *__restrict__ sin_t, size_t cos_rows, int flags,
               const __grid_constant__ typename R::Params params, const float *__restrict__ scale0,
               const float *__restrict__ scale1)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the table is constant data; everything below reads what the predecessor wrote
    griddep_launch_dependents();
    const FqPoint p0 = load_point(scale0), p1 = load_point(scale1);
    const int hv = head_dim >> 4;  // vector pairs per head
    const int dv = head_dim >> 3;  // vectors per head
    const size_t work0 = tokens * (size_t)t0.heads * hv, work1 = tokens * (size_t)t1.heads * hv;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < work0 + work1; w += stride) {
        const bool second = w >= work0;
        const RopeTensor &t = second ? t1 : t0;
        const FqPoint &pt = second ? p1 : p0;
        const size_t i = second ? w - work0 : w;
        const int c = (int)(i % hv);
        const size_t th = i / hv;
        const int h = (int)(th % t.heads);
        const size_t tok = th / t.heads;
        const uint4 *xp = t.x + tok * t.ld_x_vec + (size_t)h * dv;
        const size_t crow = (tok % cos_rows) * dv;
        const uint4 xa = __ldcs(xp + c), xb = __ldcs(xp + c + hv);
        const uint4 ca4 = __ldg(cos_t + crow + c), cb4 = __ldg(cos_t + crow + c + hv);
        const uint4 sa4 = __ldg(sin_t + crow + c), sb4 = __ldg(sin_t + crow + c + hv);
        float a[8], b[8], ra[8], rb[8];
        unpack8(xa, a);
        unpack8(xb, b);
        {
            float cw[8], sw[8], t2[8];
            unpack8(ca4, cw);
            unpack8(sa4, sw);
            // first half: x1 * cos - x2 * sin (rotate_half gives -x2)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ra[k] = a[k] * cw[k];
                t2[k] = -b[k] * sw[k];
            }
            round8(ra);
            round8(t2);
#pragma unroll
            for (int k = 0; k < 8; ++k) ra[k] += t2[k];
            round8(ra);
            unpack8(cb4, cw);
            unpack8(sb4, sw);
            // second half: x2 * cos + x1 * sin
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                rb[k] = b[k] * cw[k];
                t2[k] = a[k] * sw[k];
            }
            round8(rb);
            round8(t2);
#pragma unroll
            for (int k = 0; k < 8; ++k) rb[k] += t2[k];
            round8(rb);
        }
        if (flags & FQ_POST) {
            fq8<R, SCALED>(round, ra, pt);
            fq8<R, SCALED>(round, rb, pt);
        }
        const size_t yv = tok * t.ld_y_vec + (size_t)h * dv + c;
        store8(t.y, yv, ra, out_type);
        store8(t.y, yv + hv, rb, out_type);
    }
}

// ----------------------------------------------------------------------------- fq + transpose
// v: [B, S, H, D] (token stride ld_tok, heads contiguous) -> out [B, H, D, S] contiguous, fake-quantized: the K-major
// operand of probabilities x values.  One CTA per (b, h, 64 tokens): coalesced loads, fq, padded smem tile, then
// 128-byte rows of the transposed output.
constexpr int TR_TOK = 64;
template <class R, bool SCALED>
__global__ void __launch_bounds__(EW_THREADS, EW_MIN_CTAS)
fq_transpose_kernel(const uint16_t *__restrict__ v, void *__restrict__ out_v, int out_type, int B, int S, int H, int D,
                    size_t ld_tok, size_t batch_stride, int flags, const __grid_constant__ typename R::Params params,
                    const float *__restrict__ scale_post)
{
    const unsigned char *tab = stage_table<R>(params);
    const R round(params, tab);
    griddep_wait();  // PDL
    griddep_launch_dependents();
    const FqPoint post = load_point(scale_post);
    // tile after the (optional) rounding table in dynamic shared memory: [TR_TOK][D + 2] bf16
    uint16_t *tile = reinterpret_cast<uint16_t *>(qt_dyn_smem + R::kSmemBytes);
    const int pitch = D + 2;
    const int s_tiles = (S + TR_TOK - 1) / TR_TOK;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int job = blockIdx.x; job < B * H * s_tiles; job += gridDim.x) {
        const int st = job % s_tiles, bh = job / s_tiles, h = bh % H, b = bh / H;
        const int s0 = st * TR_TOK;
        const int dvec = D >> 3;
        __syncthreads();  // previous tile fully written out
        for (int i = tid; i < TR_TOK * dvec; i += nthr) {
            const int r = i / dvec, c = i - r * dvec;
            float f[8];
            if (s0 + r < S) {
                const uint4 *src = reinterpret_cast<const uint4 *>(v + (size_t)b * batch_stride + (size_t)(s0 + r) * ld_tok +
                                                                   (size_t)h * D) + c;
                unpack8(__ldcs(src), f);
                if (flags & FQ_POST) fq8<R, SCALED>(round, f, post);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = 0.0f;
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(tile + r * pitch + c * 8);  // (r * pitch + c * 8) is even
            const uint4 pk = pack8(f);
            dst[0] = pk.x;
            dst[1] = pk.y;
            dst[2] = pk.z;
            dst[3] = pk.w;
        }
        __syncthreads();
        // out[b, h, d, s0 + 8 g .. + 8): lane -> (d_sub = lane / 8, g = lane % 8)
        const bool full = (s0 + TR_TOK <= S) && (S % 8 == 0);
        for (int i = tid; i < D * (TR_TOK / 8); i += nthr) {
            const int d = i >> 3, g = i & 7;
            uint16_t e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = tile[(g * 8 + k) * pitch + d];
            const size_t o0 = ((size_t)bh * D + d) * S + s0 + g * 8;
            if (out_type == OUT_BF16) {
                uint16_t *dst = static_cast<uint16_t *>(out_v) + o0;
                if (full) {
                    uint4 o;
                    o.x = e[0] | ((uint32_t)e[1] << 16);
                    o.y = e[2] | ((uint32_t)e[3] << 16);
                    o.z = e[4] | ((uint32_t)e[5] << 16);
                    o.w = e[6] | ((uint32_t)e[7] << 16);
                    *reinterpret_cast<uint4 *>(dst) = o;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (s0 + g * 8 + k < S) dst[k] = e[k];
                }
            } else {
                uint8_t *dst = static_cast<uint8_t *>(out_v) + o0;
                float fe[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) fe[k] = __uint_as_float((uint32_t)e[k] << 16);
                const uint2 cc = codes8(fe, out_type);
                const uint32_t c2[4] = {cc.x & 0xFFFFu, cc.x >> 16, cc.y & 0xFFFFu, cc.y >> 16};
                if (full) {
                    uint2 o;
                    o.x = c2[0] | (c2[1] << 16);
                    o.y = c2[2] | (c2[3] << 16);
                    *reinterpret_cast<uint2 *>(dst) = o;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (s0 + g * 8 + k < S) dst[k] = (uint8_t)((c2[k >> 1] >> ((k & 1) * 8)) & 0xFFu);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- launch helpers
// Rounding engines of the fused kernels: the binade-constant table for fp / posit formats (the caller always has it:
// qt_lut_build_host), the direct logic for intN / uintN and the identity ("bfloat16": the op without fake quant).
template <class Fn>
int dispatch_fused(const QtRound &P, const void *lut, Fn &&fn)
{
    if (P.kind == QTR_INT || P.kind == QTR_IDENTITY) return dispatch_direct_small(P, fn);
    QtLutCfg cfg;
    if (!lut || qt_lut_config(P, &cfg) != QT_OK) {
        qt_set_error("qt_b200 fused ops: fp / posit formats need the device table from qt_lut_build_host(fmt)");
        return QT_ERR_INVALID_ARGUMENT;
    }
    TableParams tp;
    tp.cfg = cfg;
    if (reinterpret_cast<uintptr_t>(lut) & 15u) {
        qt_set_error("qt_b200: lut must be 16-byte aligned");
        return QT_ERR_UNALIGNED;
    }
    tp.table = static_cast<const QtLutEntry *>(lut);
    fn(RounderTag<TableRounderDyn>{}, tp);
    return QT_OK;
}
int check_common(const char *fn, const qt_format_t *fmt, QtRound *P)
{
    if (!fmt) {
        qt_set_error("%s: fmt is NULL", fn);
        return QT_ERR_INVALID_ARGUMENT;
    }
    int rc = qt_make_round(fmt, P);
    if (rc != QT_OK) return rc;
    if (num_sms() == 0) return no_device();
    return QT_OK;
}
bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// fp8 codes may only follow an unscaled fake-quant step of the same fp8 format (decode(code) == value)
int check_out_type(const char *fn, int *out_type_io, int fq_points, const qt_format_t *fmt, const float *scale_post)
{
    const int out_type = *out_type_io;
    if (out_type == OUT_BF16) return QT_OK;
    const bool e4m3 = fmt->kind == QT_KIND_FP && fmt->ebits == 4 && fmt->mbits == 3 && !fmt->is_unsigned;
    const bool e5m2 = fmt->kind == QT_KIND_FP && fmt->ebits == 5 && fmt->mbits == 2 && !fmt->is_unsigned;
    if ((out_type != OUT_E4M3 && out_type != OUT_E5M2) || !(fq_points & FQ_POST) || scale_post ||
        (out_type == OUT_E4M3 && !e4m3) || (out_type == OUT_E5M2 && !e5m2)) {
        qt_set_error("%s: fp8 code output needs an unscaled output fake-quant step of the same fp8 format", fn);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (fmt->flavour == QT_FP_MX) *out_type_io |= OUT_INF_POSSIBLE;  // fpN_eXmY lets +-Inf through
    return QT_OK;
}

int finish(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return QT_OK;
}

}  // namespace

extern "C" int qt_softmax_fq(const void *scores, void *probs, size_t rows, size_t cols, float alpha, const void *mask,
                             size_t rows_per_batch, size_t mask_rows, size_t mask_batches, int fq_points,
                             int out_type, const qt_format_t *fmt, const float *scale_pre, const float *scale_mid,
                             const float *scale_post, const void *lut, const int32_t *causal_flag, void *stream)
{
    QtRound P;
    int rc = check_common("qt_softmax_fq", fmt, &P);
    if (rc != QT_OK) return rc;
    rc = check_out_type("qt_softmax_fq", &out_type, fq_points, fmt, scale_post);
    if (rc != QT_OK) return rc;
    if (rows == 0 || cols == 0) return QT_OK;
    if ((fq_points & FQ_CAUSAL) && (!mask || mask_rows != cols || (fq_points & FQ_MID))) {
        qt_set_error("qt_softmax_fq: QT_SOFTMAX_CAUSAL needs the causal mask of a square attention (mask_rows == cols) "
                     "and no QT_FQ_MID step");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (!scores || !probs || cols % 8 || cols > 4096 || !aligned16(scores) || !aligned16(probs) ||
        (mask && (!aligned16(mask) || mask_rows == 0 || rows_per_batch == 0 || mask_batches == 0))) {
        qt_set_error("qt_softmax_fq: needs 16-byte aligned contiguous bf16 rows, cols %% 8 == 0, cols <= 4096 (got %zu)",
                     cols);
        return QT_ERR_INVALID_ARGUMENT;
    }
    const size_t mask_batch_stride_vec = mask_batches > 1 ? mask_rows * (cols / 8) : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int has_alpha = alpha != 1.0f;
    const bool scaled = scale_pre || scale_mid || scale_post;
    const bool simple = !(fq_points & (FQ_PRE | FQ_MID)) && getenv("QT_SOFTMAX_GENERAL") == nullptr;
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
#define QT_SOFTMAX_LAUNCH(TPR, VPL)                                                                                  \
    do {                                                                                                             \
        const size_t ctas = (rows + (ROW_THREADS / TPR) - 1) / (ROW_THREADS / TPR);                                  \
        auto kernel = scaled ? softmax_fq_kernel<R, true, TPR, VPL> : softmax_fq_kernel<R, false, TPR, VPL>;         \
        if (simple && TPR == 32 && VPL >= 2)                                                                         \
            kernel = scaled ? softmax_fq_kernel<R, true, TPR, (VPL >= 2 ? VPL : 2), true>                            \
                            : softmax_fq_kernel<R, false, TPR, (VPL >= 2 ? VPL : 2), true>;                          \
        qt_launch(kernel, dim3(grid_for(ctas, ROW_MIN_CTAS * 2)), dim3(ROW_THREADS), R::kSmemBytes, st,                                \
            static_cast<const uint4 *>(scores), probs, out_type, rows, (int)cols, alpha, has_alpha,                  \
            static_cast<const uint4 *>(mask), rows_per_batch, mask_rows, mask_batch_stride_vec, fq_points, params,   \
            scale_pre, scale_mid, scale_post, causal_flag);                                                          \
    } while (0)
        if (cols <= 256)
            QT_SOFTMAX_LAUNCH(32, 1);
        else if (cols <= 512)
            QT_SOFTMAX_LAUNCH(32, 2);
        else if (cols <= 1024)
            QT_SOFTMAX_LAUNCH(32, 4);
        else if (cols <= 2048)
            QT_SOFTMAX_LAUNCH(128, 2);
        else
            QT_SOFTMAX_LAUNCH(128, 4);
#undef QT_SOFTMAX_LAUNCH
    });
    if (rc != QT_OK) return rc;
    return finish("softmax_fq kernel launch");
}

extern "C" int qt_norm_fq(const void *x, void *y, void *y_raw, size_t rows, size_t cols, int kind, const void *weight,
                          const void *bias, float eps, int fq_points, int out_type, const qt_format_t *fmt,
                          const float *scale_pre, const float *scale_post, const void *lut, void *stream)
{
    if (kind == 3 || (fq_points & (FQ_RES_A | FQ_RES_B))) {
        qt_set_error("qt_norm_fq: kind 3 and QT_FQ_RES_* belong to qt_add_norm_fq");
        return QT_ERR_INVALID_ARGUMENT;
    }
    return qt_add_norm_fq(x, nullptr, y, y_raw, rows, cols, kind, weight, bias, eps, fq_points, out_type, fmt, scale_pre,
                          scale_post, lut, stream);
}

extern "C" int qt_add_norm_fq(const void *x, const void *res, void *y, void *y_raw, size_t rows, size_t cols, int kind,
                              const void *weight, const void *bias, float eps, int fq_points, int out_type,
                              const qt_format_t *fmt, const float *scale_pre, const float *scale_post, const void *lut,
                              void *stream)
{
    QtRound P;
    int rc = check_common("qt_norm_fq", fmt, &P);
    if (rc != QT_OK) return rc;
    rc = check_out_type("qt_norm_fq", &out_type, fq_points, fmt, scale_post);
    if (rc != QT_OK) return rc;
    if (rows == 0 || cols == 0) return QT_OK;
    if (!x || !y || (!weight && kind != 3) || cols % 8 || cols > 8192 || kind < 0 || kind > 3 || (kind == 2 && !bias) ||
        (kind == 3 && !res) || (!res && (fq_points & (FQ_RES_A | FQ_RES_B))) || (res && !aligned16(res)) || !aligned16(x) ||
        !aligned16(y) ||
        (y_raw && !aligned16(y_raw)) ||
        (weight && !aligned16(weight)) || (bias && !aligned16(bias))) {
        qt_set_error("qt_norm_fq: needs 16-byte aligned contiguous bf16 rows, cols %% 8 == 0, cols <= 8192 (got %zu), "
                     "kind 0 (RMSNorm) or 1 (LayerNorm)", cols);
        return QT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool scaled = scale_pre || scale_post;
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
#define QT_NORM_LAUNCH(TPR, VPL)                                                                                 \
    do {                                                                                                         \
        const size_t ctas = (rows + (ROW_THREADS / TPR) - 1) / (ROW_THREADS / TPR);                              \
        auto kernel = scaled ? norm_fq_kernel<R, true, TPR, VPL> : norm_fq_kernel<R, false, TPR, VPL>;           \
        qt_launch(kernel, dim3(grid_for(ctas, ROW_MIN_CTAS * 2)), dim3(ROW_THREADS), R::kSmemBytes, st,                            \
            static_cast<const uint4 *>(x), static_cast<const uint4 *>(res), y, static_cast<uint4 *>(y_raw), out_type, rows, (int)cols, kind, \
            static_cast<const uint4 *>(weight), static_cast<const uint4 *>(bias), eps, fq_points, params,        \
            scale_pre, scale_post);                                                                              \
    } while (0)
        if (cols <= 256)
            QT_NORM_LAUNCH(32, 1);
        else if (cols <= 512)
            QT_NORM_LAUNCH(32, 2);
        else if (cols <= 1024)
            QT_NORM_LAUNCH(32, 4);
        else if (cols <= 2048)
            QT_NORM_LAUNCH(128, 2);
        else if (cols <= 4096)
            QT_NORM_LAUNCH(128, 4);
        else
            QT_NORM_LAUNCH(128, 8);
#undef QT_NORM_LAUNCH
    });
    if (rc != QT_OK) return rc;
    return finish("norm_fq kernel launch");
}

extern "C" int qt_act_mul_fq(const void *gate, const void *up, void *out, size_t rows, size_t cols, size_t ld_gate,
                             size_t ld_up, size_t ld_out, int activation, int fq_points, int out_type,
                             const qt_format_t *fmt, const float *scale_post, const void *lut, void *stream)
{
    QtRound P;
    int rc = check_common("qt_act_mul_fq", fmt, &P);
    if (rc != QT_OK) return rc;
    rc = check_out_type("qt_act_mul_fq", &out_type, fq_points, fmt, scale_post);
    if (rc != QT_OK) return rc;
    if (rows == 0 || cols == 0) return QT_OK;
    if (!gate || !out || cols % 8 || ld_gate % 8 || ld_out % 8 || (up && ld_up % 8) || !aligned16(gate) ||
        !aligned16(out) || (up && !aligned16(up)) || activation < FACT_NONE || activation > FACT_SILU) {
        qt_set_error("qt_act_mul_fq: needs 16-byte aligned bf16 rows (cols, row strides multiples of 8) and a QT_ACT_* "
                     "activation");
        return QT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
        const size_t total = rows * (cols / 8);
        const unsigned grid = grid_for((total + EW_THREADS * 2 - 1) / (EW_THREADS * 2), ACT_MIN_CTAS * 2);
        void (*kernel)(const uint4 *, const uint4 *, void *, int, size_t, int, size_t, size_t, size_t, int,
                       const typename R::Params, const float *) = nullptr;
        const bool scaled = scale_post != nullptr;
        switch (activation * 2 + (scaled ? 1 : 0)) {
        case 0: kernel = act_mul_fq_kernel<R, false, FACT_NONE>; break;
        case 1: kernel = act_mul_fq_kernel<R, true, FACT_NONE>; break;
        case 2: kernel = act_mul_fq_kernel<R, false, FACT_RELU>; break;
        case 3: kernel = act_mul_fq_kernel<R, true, FACT_RELU>; break;
        case 4: kernel = act_mul_fq_kernel<R, false, FACT_GELU>; break;
        case 5: kernel = act_mul_fq_kernel<R, true, FACT_GELU>; break;
        case 6: kernel = act_mul_fq_kernel<R, false, FACT_SILU>; break;
        default: kernel = act_mul_fq_kernel<R, true, FACT_SILU>; break;
        }
        qt_launch(kernel, dim3(grid), dim3(EW_THREADS), R::kSmemBytes, st,
            static_cast<const uint4 *>(gate), static_cast<const uint4 *>(up), out, out_type, rows, (int)(cols / 8),
            ld_gate / 8, ld_up / 8, ld_out / 8, fq_points, params, scale_post);
    });
    if (rc != QT_OK) return rc;
    return finish("act_mul_fq kernel launch");
}

extern "C" int qt_lora_merge_fq(const void *W, const void *A, const void *B, void *out, size_t N, size_t K, int r,
                                float scaling, int fq_points, const qt_format_t *fmt, const float *scale_post,
                                const void *lut, void *stream)
{
    QtRound P;
    int rc = check_common("qt_lora_merge_fq", fmt, &P);
    if (rc != QT_OK) return rc;
    if (N == 0 || K == 0) return QT_OK;
    const size_t a_bytes = (size_t)r * K * 2;
    if (!W || !A || !B || !out || K % 8 || r < 1 || r > 256 || a_bytes > 160 * 1024 || !aligned16(W) || !aligned16(A) ||
        !aligned16(out) || (reinterpret_cast<uintptr_t>(B) & 1u) || (fq_points & ~(FQ_PRE | FQ_POST))) {
        qt_set_error("qt_lora_merge_fq: needs contiguous 16-byte aligned bf16 W / out [N, K], A [r, K], B [N, r], "
                     "K %% 8 == 0, r * K * 2 <= 160 KB; fq_points in {QT_FQ_PRE, QT_FQ_POST}");
        return QT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
        const size_t total = N * (K / 8);
        const size_t smem = R::kSmemBytes + a_bytes;
        const unsigned grid = grid_for((total + EW_THREADS - 1) / EW_THREADS, EW_MIN_CTAS);
        auto kernel = scale_post ? lora_merge_fq_kernel<R, true> : lora_merge_fq_kernel<R, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        qt_launch(kernel, dim3(grid), dim3(EW_THREADS), smem, st, static_cast<const uint4 *>(W),
                  static_cast<const uint4 *>(A), static_cast<const uint16_t *>(B), static_cast<uint4 *>(out), N,
                  (int)(K / 8), r, scaling, fq_points, params, scale_post);
    });
    if (rc != QT_OK) return rc;
    return finish("lora_merge_fq kernel launch");
}

extern "C" int qt_rope_fq(const void *q, void *q_out, size_t ld_q, size_t ld_q_out, int q_heads, const void *k,
                          void *k_out, size_t ld_k, size_t ld_k_out, int k_heads, size_t tokens, int head_dim,
                          const void *cos_table, const void *sin_table, size_t cos_rows, int fq_points, int out_type,
                          const qt_format_t *fmt, const float *scale_q, const float *scale_k, const void *lut,
                          void *stream)
{
    QtRound P;
    int rc = check_common("qt_rope_fq", fmt, &P);
    if (rc != QT_OK) return rc;
    rc = check_out_type("qt_rope_fq", &out_type, fq_points, fmt, scale_q ? scale_q : scale_k);
    if (rc != QT_OK) return rc;
    if (tokens == 0) return QT_OK;
    if (!q || !q_out || !cos_table || !sin_table || cos_rows == 0 || head_dim % 16 || head_dim <= 0 || ld_q % 8 ||
        ld_q_out % 8 || q_heads < 1 || !aligned16(q) || !aligned16(q_out) || !aligned16(cos_table) ||
        !aligned16(sin_table) ||
        (k && (!k_out || ld_k % 8 || ld_k_out % 8 || k_heads < 1 || !aligned16(k) || !aligned16(k_out)))) {
        qt_set_error("qt_rope_fq: needs 16-byte aligned bf16 tensors, head_dim %% 16 == 0, token strides %% 8 == 0");
        return QT_ERR_INVALID_ARGUMENT;
    }
    RopeTensor t0 = {static_cast<const uint4 *>(q), q_out, ld_q / 8, ld_q_out / 8, q_heads};
    RopeTensor t1 = {static_cast<const uint4 *>(k), k_out, ld_k / 8, ld_k_out / 8,
                     k ? k_heads : 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
        const size_t total = tokens * (size_t)(t0.heads + t1.heads) * (head_dim / 16);
        const unsigned grid = grid_for((total + EW_THREADS - 1) / EW_THREADS, EW_MIN_CTAS * 2);
        auto kernel = (scale_q || scale_k) ? rope_fq_kernel<R, true> : rope_fq_kernel<R, false>;
        qt_launch(kernel, dim3(grid), dim3(EW_THREADS), R::kSmemBytes, st,
            t0, t1, out_type, tokens, head_dim, static_cast<const uint4 *>(cos_table),
            static_cast<const uint4 *>(sin_table), cos_rows, fq_points, params, scale_q, scale_k);
    });
    if (rc != QT_OK) return rc;
    return finish("rope_fq kernel launch");
}

extern "C" int qt_fq_transpose(const void *v, void *out, int batch, int seq, int heads, int head_dim, size_t ld_tok,
                               size_t batch_stride, int fq_points, int out_type, const qt_format_t *fmt,
                               const float *scale_post, const void *lut, void *stream)
{
    QtRound P;
    int rc = check_common("qt_fq_transpose", fmt, &P);
    if (rc != QT_OK) return rc;
    rc = check_out_type("qt_fq_transpose", &out_type, fq_points, fmt, scale_post);
    if (rc != QT_OK) return rc;
    if (batch <= 0 || seq <= 0 || heads <= 0) return QT_OK;
    if (!v || !out || head_dim % 8 || head_dim <= 0 || head_dim > 256 || ld_tok % 8 || batch_stride % 8 ||
        !aligned16(v) || !aligned16(out)) {
        qt_set_error("qt_fq_transpose: needs 16-byte aligned bf16 tensors, head_dim %% 8 == 0 (<= 256), strides %% 8 == 0");
        return QT_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = dispatch_fused(P, lut, [&](auto tag, const auto &params) {
        using R = typename decltype(tag)::type;
        const size_t jobs = (size_t)batch * heads * ((seq + TR_TOK - 1) / TR_TOK);
        const size_t smem = R::kSmemBytes + (size_t)TR_TOK * (head_dim + 2) * 2;
        const unsigned grid = grid_for(jobs, EW_MIN_CTAS * 2);
        auto kernel = scale_post ? fq_transpose_kernel<R, true> : fq_transpose_kernel<R, false>;
        qt_launch(kernel, dim3(grid), dim3(EW_THREADS), smem, st, static_cast<const uint16_t *>(v), out, out_type, batch, seq, heads,
                                               head_dim, ld_tok, batch_stride, fq_points, params,
                                                                scale_post);
    });
    if (rc != QT_OK) return rc;
    return finish("fq_transpose kernel launch");
}

// ----------------------------------------------------------------------------- stream capture identity
// 0 when `stream` is not capturing, else the id of the capture in progress (cudaStreamGetCaptureInfo): lets the host
// side cache "one launch per captured graph" work (the causal-mask check) without leaking it across graphs.
extern "C" unsigned long long qt_stream_capture_id(void *stream)
{
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(static_cast<cudaStream_t>(stream), &status, &id) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return status == cudaStreamCaptureStatusActive ? (id ? id : 1ull) : 0ull;
}

// ----------------------------------------------------------------------------- causal mask detection
namespace {
// mask [batches, rows, rows] bf16: standard causal = finfo(bf16).min (0xFF7F) strictly above the diagonal, +-0 elsewhere
__global__ void __launch_bounds__(256)
causal_mask_check_kernel(const uint16_t *__restrict__ mask, size_t batches, size_t rows, int32_t *__restrict__ flag)
{
    const size_t total = batches * rows * rows;
    bool ok = true;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i % rows, r = (i / rows) % rows;
        const uint16_t v = mask[i];
        ok = ok && (c > r ? v == 0xFF7Fu : (v & 0x7FFFu) == 0u);
    }
    if (!__all_sync(0xFFFFFFFFu, ok) && (threadIdx.x & 31) == 0) *flag = 0;
}
__global__ void set_flag_kernel(int32_t *flag, int32_t v) { *flag = v; }
}  // namespace

extern "C" int qt_causal_mask_check(const void *mask, size_t batches, size_t rows, int32_t *flag_out, void *stream)
{
    if (!flag_out) {
        qt_set_error("qt_causal_mask_check: flag_out is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (num_sms() == 0) return no_device();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool possible = mask != nullptr && batches > 0 && rows > 0;
    set_flag_kernel<<<1, 1, 0, st>>>(flag_out, possible ? 1 : 0);
    if (possible) {
        const size_t total = batches * rows * rows;
        causal_mask_check_kernel<<<grid_for((total + 255) / 256, 8), 256, 0, st>>>(static_cast<const uint16_t *>(mask),
                                                                                   batches, rows, flag_out);
    }
    return finish("causal mask check launch");
}
