// qt_fq.cu -- fused quantize-dequantize + amax kernels for sm_100a, and their C-ABI launchers.
//
// One pass over HBM per call: 128-bit coalesced streaming loads -> (x / s) -> round to the format
// -> (* s) -> 128-bit streaming stores, while max|x| of the UNSCALED input is reduced warp-wide
// (redux.sync) and block-wide and merged with one atomicMax per CTA.  The reference's scaling is
// delayed (the scale applied now comes from earlier calls, fake_quantize.py:230-242), so no grid-wide
// dependency exists and one pass suffices.
//
// Replaces (reference, src/quantized_training/): fake_quantize.py:217-223 (amax),
// :244-246 (divide, vmap, multiply), decomposed.py:146-163 (vmap's Python chunk loop).
//
// Rounding engines (same results, checked against each other and the reference's tables):
//   Direct   qt_round.h -- bitwise logic, integer ALU.  int/uint use it always (a handful of ops).
//   Table    qt_lut.h   -- per-binade constants in shared memory, 2 FFMA + 1 LDS.128 per element.
//            Used for fp and posit formats when the caller supplies the 8 KB table: the direct posit
//            path needs ~30 ALU-pipe ops per element and caps the kernel at ~40 % of HBM bandwidth.
//
// Layouts:  x, y contiguous [outer, channels, inner].
//   flat  kernel: channels == 1 (per-tensor or bare spec); persistent grid-stride over 16-byte vectors.
//   rows  kernel: inner > 1 per-channel (e.g. weights [out, in], ax=0): one scale per row segment.
//   cols  kernel: inner == 1 per-channel (ax = last dim): one scale per column, register-resident.
//   scalar kernel: tails, misaligned views, odd per-channel shapes.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <string.h>

#include "qt_fq_common.cuh"
#include "qt_launch.cuh"

namespace {

// ----------------------------------------------------------------------------- flat kernel
// channels == 1.  nvec 16-byte vectors; tile = kThreads * kUnroll vectors; persistent grid.
template <class R, bool F32, int DIV, bool AMAX>
__device__ __forceinline__ void fq_span(const R &round, const uint4 *__restrict__ x, uint4 *__restrict__ y,
                                        size_t nvec, size_t first_tile, size_t tile_stride, const ScaleBf16 &sc,
                                        uint32_t &amax)
{
    const size_t nthr = blockDim.x;
    const size_t tile = nthr * kUnroll;
    const size_t ntiles = (nvec + tile - 1) / tile;
    for (size_t t = first_tile; t < ntiles; t += tile_stride) {
        const size_t base = t * tile + threadIdx.x;
        uint4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * nthr;
            v[j] = i < nvec ? ld_stream(x + i) : make_uint4(0u, 0u, 0u, 0u);
        }
        if constexpr (!F32 && DIV == DIV_RECIP_NOTINY) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const size_t i = base + (size_t)j * nthr;
                if (AMAX) amax = amax_of_vec_bf16(amax, v[j]);
                const uint4 r = fq_vec_bf16_recip_notiny<R>(round, v[j], sc);
                if (i < nvec) st_stream(y + i, r);
            }
        } else if constexpr (!F32 && DIV == DIV_RECIP) {
            // Reciprocal path: the inputs are consumed as they are processed (no register is held back for a
            // fallback); the rare tile with a sub-2^-120 quotient is redone from memory with the true division.
            bool tiny = false;
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const size_t i = base + (size_t)j * nthr;
                if (AMAX) amax = amax_of_vec_bf16(amax, v[j]);
                const uint4 r = fq_vec_bf16_recip_fast<R>(round, v[j], sc, tiny);
                if (i < nvec) st_stream(y + i, r);
            }
            if (tiny) {
                for (int j = 0; j < kUnroll; ++j) {
                    const size_t i = base + (size_t)j * nthr;
                    if (i < nvec) {
                        uint32_t unused = 0u;
                        st_stream(y + i, fq_vec<R, false, DIV_EXACT, false>(round, ld_stream(x + i), sc, unused));
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const size_t i = base + (size_t)j * nthr;
                const uint4 r = fq_vec<R, F32, DIV, AMAX>(round, v[j], sc, amax);
                if (i < nvec) st_stream(y + i, r);
            }
        }
    }
}

// The scale is one number for the whole launch.  When it is exactly 1 (bare specs such as "e4m3" or
// "posit8_1", whose `scale` buffer is never written) x / 1 and q * 1 are identities and the CTA takes a
// path without divide / multiply / re-rounding.  The branch is grid-uniform and made on the device, so
// the host never reads the scale back.
template <class R, bool F32, bool AMAX>
__global__ void __launch_bounds__(R::kThreads, R::kMinCtas)
fq_flat_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t nvec,
               const __grid_constant__ typename R::Params params, const float *__restrict__ scale,
               float *__restrict__ amax_out)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the scale and the input may come from the kernel before this one
    griddep_launch_dependents();
    ScaleBf16 sc = {1.0f, 1.0f};
    if (scale) sc = load_scale<F32>(scale, 0);
    uint32_t amax = 0u;
    const int mode = classify_scale(sc.s);
    if (mode == DIV_UNIT)
        fq_span<R, F32, DIV_UNIT, AMAX>(round, x, y, nvec, blockIdx.x, gridDim.x, sc, amax);
    else if (F32 || mode == DIV_EXACT)
        fq_span<R, F32, DIV_EXACT, AMAX>(round, x, y, nvec, blockIdx.x, gridDim.x, sc, amax);
    else if (R::tiny_safe(params))
        fq_span<R, F32, DIV_RECIP_NOTINY, AMAX>(round, x, y, nvec, blockIdx.x, gridDim.x, sc, amax);
    else
        fq_span<R, F32, DIV_RECIP, AMAX>(round, x, y, nvec, blockIdx.x, gridDim.x, sc, amax);
    if (AMAX) block_amax_commit(amax, amax_out);
}

// observer only (fake quant disabled): read, reduce, no store
template <bool F32>
__global__ void __launch_bounds__(256)
amax_flat_kernel(const uint4 *__restrict__ x, size_t nvec, float *__restrict__ amax_out)
{
    uint32_t amax = 0u;
    const size_t nthr = blockDim.x;
    const size_t tile = nthr * kUnroll;
    const size_t ntiles = (nvec + tile - 1) / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        uint4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * nthr;
            v[j] = i < nvec ? ld_stream(x + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) amax = F32 ? amax_of_vec_f32(amax, v[j]) : amax_of_vec_bf16(amax, v[j]);
    }
    block_amax_commit(amax, amax_out);
}

// ----------------------------------------------------------------------------- scalar kernel
// Any layout, any alignment: element i belongs to channel (i / inner) % channels.  Used for tails,
// misaligned views and odd per-channel shapes.  Elements [first, first + count).  WRITE = false: observe only.
template <class R, bool F32, bool AMAX, bool WRITE>
__global__ void __launch_bounds__(R::kThreads)
fq_scalar_kernel(const void *__restrict__ xv, void *__restrict__ yv, size_t first, size_t count, size_t channels,
                 size_t inner, const __grid_constant__ typename R::Params params, const float *__restrict__ scale,
                 float *__restrict__ amax_out)
{
    const R round(params, stage_table<R>(params));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t amax1 = 0u;  // channels == 1: reduce in registers, one atomic per CTA
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        const size_t i = first + j;
        const size_t c = channels == 1 ? 0 : (i / inner) % channels;
        ScaleBf16 sc = {1.0f, 1.0f};
        if (WRITE && scale) sc = load_scale<F32>(scale, c);
        uint32_t ab;
        if (F32) {
            const uint32_t bits = static_cast<const uint32_t *>(xv)[i];
            ab = bits & 0x7FFFFFFFu;
            if (WRITE) static_cast<uint32_t *>(yv)[i] = fq_f32<R, false>(round, bits, sc.s);
        } else {
            const uint32_t bits = (uint32_t) static_cast<const uint16_t *>(xv)[i] << 16;
            ab = bits & 0x7FFFFFFFu;
            if (WRITE) static_cast<uint16_t *>(yv)[i] = (uint16_t)(fq_bf16<R, DIV_EXACT>(round, bits, sc) >> 16);
        }
        if (AMAX) {
            if (channels == 1)
                amax1 = max(amax1, ab);
            else if (ab != 0u)
                atomicMax(reinterpret_cast<unsigned int *>(amax_out) + c, ab);
        }
    }
    if (AMAX && channels == 1) block_amax_commit(amax1, amax_out);
}

// ----------------------------------------------------------------------------- rows kernel
// inner > 1 per-channel, inner % VEC == 0, 16-byte aligned (weights [out, in] with ax = 0, or [B, C, HW]).
// One WARP per row segment: row = o * channels + c, a segment is 32 * kUnroll * kRowIters vectors of that row.
// Scale load, mode choice and the amax merge (redux + one atomicMax) are warp-local: no block barrier, and
// short rows (a 4096-wide bf16 row is 512 vectors) keep every lane busy.
constexpr int kRowIters = 4;
template <class R, bool F32, int DIV, bool AMAX>
__device__ __forceinline__ void fq_row_segment(const R &round, const uint4 *__restrict__ xr, uint4 *__restrict__ yr,
                                               size_t v0, size_t v1, const ScaleBf16 &sc, uint32_t &amax)
{
    const int lane = threadIdx.x & 31;
    for (size_t base = v0 + lane; base < v1; base += 32 * kUnroll) {
        uint4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * 32;
            v[j] = i < v1 ? ld_stream(xr + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * 32;
            const uint4 r = fq_vec<R, F32, DIV, AMAX>(round, v[j], sc, amax);
            if (i < v1) st_stream(yr + i, r);
        }
    }
}

template <class R, bool F32, bool AMAX, bool WRITE>
__global__ void __launch_bounds__(R::kThreads, R::kMinCtas)
fq_rows_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t rows, size_t channels, size_t vec_per_row,
               size_t segs_per_row, const __grid_constant__ typename R::Params params,
               const float *__restrict__ scale, float *__restrict__ amax_out)
{
    const R round(params, stage_table<R>(params));
    const size_t seg_vecs = (size_t)32 * kUnroll * kRowIters;
    const size_t work = rows * segs_per_row;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t wi = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < work; wi += nwarps) {
        const size_t row = wi / segs_per_row, seg = wi - row * segs_per_row;
        const size_t c = row % channels;
        const uint4 *xr = x + row * vec_per_row;
        const size_t v0 = seg * seg_vecs, v1 = min(vec_per_row, v0 + seg_vecs);
        uint32_t amax = 0u;
        if (WRITE) {
            const ScaleBf16 sc = load_scale<F32>(scale, c);
            const int mode = classify_scale(sc.s);
            uint4 *yr = y + row * vec_per_row;
            if (mode == DIV_UNIT)
                fq_row_segment<R, F32, DIV_UNIT, AMAX>(round, xr, yr, v0, v1, sc, amax);
            else if (F32 || mode == DIV_EXACT)
                fq_row_segment<R, F32, DIV_EXACT, AMAX>(round, xr, yr, v0, v1, sc, amax);
            else if (R::tiny_safe(params))
                fq_row_segment<R, F32, DIV_RECIP_NOTINY, AMAX>(round, xr, yr, v0, v1, sc, amax);
            else
                fq_row_segment<R, F32, DIV_RECIP, AMAX>(round, xr, yr, v0, v1, sc, amax);
        } else {
            for (size_t i = v0 + (threadIdx.x & 31); i < v1; i += 32) {
                const uint4 v = ld_stream(xr + i);
                amax = F32 ? amax_of_vec_f32(amax, v) : amax_of_vec_bf16(amax, v);
            }
        }
        if (AMAX) {
            amax = __reduce_max_sync(0xFFFFFFFFu, amax);
            if ((threadIdx.x & 31) == 0 && amax != 0u) atomicMax(reinterpret_cast<unsigned int *>(amax_out) + c, amax);
        }
    }
}

// ----------------------------------------------------------------------------- cols kernel
// inner == 1 per-channel (channel = last dim), channels % VEC == 0, 16-byte aligned rows.
// blockDim = (32, 8): threadIdx.x -> a 16-byte column group, threadIdx.y -> row phase.
// Scales and running maxima for the thread's VEC columns stay in registers over all rows.
// kColsY row phases per CTA: 16 with a table rounder (512 threads, two CTAs per SM next to 2 x 64 KB of table = 1024
// resident threads instead of 3 x 256), 8 otherwise.
template <class R>
constexpr int kColsY = R::kTable ? 16 : 8;
template <class R, bool F32, bool AMAX, bool WRITE>
__global__ void __launch_bounds__(32 * kColsY<R>, R::kTable ? 2 : 1)
fq_cols_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t rows, size_t vec_per_row,
               const __grid_constant__ typename R::Params params, const float *__restrict__ scale,
               float *__restrict__ amax_out)
{
    constexpr int VEC = F32 ? 4 : 8;
    const R round(params, stage_table<R>(params));
    const size_t cg = (size_t)blockIdx.x * 32 + threadIdx.x;  // column group
    const bool active = cg < vec_per_row;
    ScaleBf16 sc[VEC];
    bool recip_ok[VEC];
    uint32_t am[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        sc[k].s = sc[k].rs = 1.0f;
        if (WRITE && active) sc[k] = load_scale<F32>(scale, cg * VEC + k);
        recip_ok[k] = classify_scale(sc[k].s) != DIV_EXACT;  // x * rcp(1) is exact too
        am[k] = 0u;
    }
    // every column of every lane on the reciprocal path, and the format indifferent to sub-2^-120 quotients:
    // the warp takes the branch without per-element tests
    bool mine_fast = !F32 && WRITE && R::tiny_safe(params);
#pragma unroll
    for (int k = 0; k < VEC; ++k) mine_fast = mine_fast && recip_ok[k];
    const bool fast = __all_sync(0xFFFFFFFFu, mine_fast);
    uint32_t amp[4] = {0u, 0u, 0u, 0u};  // bf16: packed running maxima of the eight columns
    if (active) {
        const size_t rstep = (size_t)gridDim.y * blockDim.y;
        for (size_t r0 = (size_t)blockIdx.y * blockDim.y + threadIdx.y; r0 < rows; r0 += rstep * kUnroll) {
            uint4 vin[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {  // kUnroll rows in flight per thread
                const size_t r = r0 + (size_t)u * rstep;
                vin[u] = r < rows ? ld_stream(x + r * vec_per_row + cg) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t r = r0 + (size_t)u * rstep;
                const uint32_t w[4] = {vin[u].x, vin[u].y, vin[u].z, vin[u].w};
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (F32) {
                        if (AMAX) am[k] = max(am[k], w[k] & 0x7FFFFFFFu);
                        if (WRITE) o[k] = fq_f32<R, false>(round, w[k], sc[k].s);
                    } else {
                        const uint32_t lo = w[k] << 16, hi = w[k] & 0xFFFF0000u;
                        if (AMAX) amp[k] = __vmaxu2(amp[k], w[k] & 0x7FFF7FFFu);  // two 16-bit maxima per word
                        if (WRITE && fast) {
                            o[k] = fq_word_bf16_recip2_notiny<R>(round, w[k], sc[2 * k], sc[2 * k + 1]);
                        } else if (WRITE) {
                            // per column: reciprocal multiply when its scale allows it, true division otherwise
                            const float qlo = recip_ok[2 * k] ? bf16_quotient<DIV_RECIP>(lo, sc[2 * k])
                                                              : bf16_quotient<DIV_EXACT>(lo, sc[2 * k]);
                            const float qhi = recip_ok[2 * k + 1] ? bf16_quotient<DIV_RECIP>(hi, sc[2 * k + 1])
                                                                  : bf16_quotient<DIV_EXACT>(hi, sc[2 * k + 1]);
                            const uint32_t uq = bf16x2_rne(qlo, qhi);
                            o[k] = bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq)), sc[2 * k].s),
                                              __fmul_rn(__uint_as_float(round.hi(uq)), sc[2 * k + 1].s));
                        }
                    }
                }
                if (WRITE && r < rows) st_stream(y + r * vec_per_row + cg, make_uint4(o[0], o[1], o[2], o[3]));
            }
        }
    }
    if (AMAX) {
        if (!F32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                am[(2 * k) % VEC] = amp[k] << 16;
                am[(2 * k + 1) % VEC] = amp[k] & 0xFFFF0000u;
            }
        }
        __shared__ uint32_t red[kColsY<R>][32][VEC + 1];
#pragma unroll
        for (int k = 0; k < VEC; ++k) red[threadIdx.y][threadIdx.x][k] = am[k];
        __syncthreads();
        if (threadIdx.y == 0 && active) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                uint32_t m = 0u;
                for (int yy = 0; yy < kColsY<R>; ++yy) m = max(m, red[yy][threadIdx.x][k]);
                if (m != 0u) atomicMax(reinterpret_cast<unsigned int *>(amax_out) + cg * VEC + k, m);
            }
        }
    }
}

// ----------------------------------------------------------------------------- quantize to fp8 codes
// Same rounding as the fake-quant pass, but the pass ends at q = round_fmt(x / s) and stores its one-byte OCP
// encoding (e4m3fn / e5m2) instead of q * s: 2 + 1 bytes per bf16 element instead of 2 + 2, and the result
// feeds the FP8 tensor-core GEMM directly (decode(code) == q exactly; NaN -> NaN code; no saturation is needed
// because q is already clamped to the format's range).
// Hardware conversion (cvt.rn.satfinite.{e4m3,e5m2}x2.f32, one instruction per pair; the non-saturating
// flavour of the intrinsic is a ~50-instruction software routine).  q is already on the format's grid, so
// the only inputs "satfinite" changes are +-Inf, which only the fpN_eXmY flavour can produce (Inf passes
// through it): patched to the Inf code for e5m2; e4m3 has no Inf and gets the NaN code.
template <bool E5M2, bool INF_POSSIBLE>
__device__ __forceinline__ uint32_t fp8x2_of(uint32_t qlo, uint32_t qhi)
{
    uint32_t c = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(__uint_as_float(qlo), __uint_as_float(qhi)),
                                                    __NV_SATFINITE, E5M2 ? __NV_E5M2 : __NV_E4M3);
    if (INF_POSSIBLE) {
        const uint32_t inf_code = E5M2 ? 0x7Cu : 0x7Fu;
        if ((qlo & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0xFF00u) | ((qlo >> 24) & 0x80u) | inf_code;
        if ((qhi & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0x00FFu) | ((((qhi >> 24) & 0x80u) | inf_code) << 8);
    }
    return c;
}

template <class R, bool F32, int DIV, bool AMAX, bool E5M2>
__device__ __forceinline__ void codes_span(const R &round, const uint4 *__restrict__ x, uint32_t *__restrict__ y,
                                           size_t nvec, const ScaleBf16 &sc, uint32_t &amax)
{
    // y is indexed in 32-bit words: a bf16 vector (8 elements) yields 2 words, an fp32 vector (4 elements) 1 word
    const size_t nthr = blockDim.x;
    const size_t tile = nthr * kUnroll;
    const size_t ntiles = (nvec + tile - 1) / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        uint4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * nthr;
            v[j] = i < nvec ? ld_stream(x + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            const size_t i = base + (size_t)j * nthr;
            const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint32_t q[8];
            if (F32) {
                if (AMAX) amax = amax_of_vec_f32(amax, v[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t u = DIV == DIV_UNIT ? w[k] : __float_as_uint(__fdiv_rn(__uint_as_float(w[k]), sc.s));
                    q[k] = round(f32_to_bf16_rto_hi(u));
                }
                const uint32_t word = fp8x2_of<E5M2, R::kMxBand>(q[0], q[1]) | (fp8x2_of<E5M2, R::kMxBand>(q[2], q[3]) << 16);
                if (i < nvec) y[i] = word;
            } else {
                if (AMAX) amax = amax_of_vec_bf16(amax, v[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (DIV == DIV_UNIT) {
                        q[2 * k] = round.lo(w[k]);
                        q[2 * k + 1] = round.hi(w[k]);
                    } else {
                        const uint32_t uq = bf16x2_rne(bf16_quotient<DIV>(w[k] << 16, sc),
                                                       bf16_quotient<DIV>(w[k] & 0xFFFF0000u, sc));
                        q[2 * k] = round.lo(uq);
                        q[2 * k + 1] = round.hi(uq);
                    }
                }
                uint2 o;
                o.x = fp8x2_of<E5M2, R::kMxBand>(q[0], q[1]) | (fp8x2_of<E5M2, R::kMxBand>(q[2], q[3]) << 16);
                o.y = fp8x2_of<E5M2, R::kMxBand>(q[4], q[5]) | (fp8x2_of<E5M2, R::kMxBand>(q[6], q[7]) << 16);
                if (i < nvec) reinterpret_cast<uint2 *>(y)[i] = o;
            }
        }
    }
}

template <class R, bool F32, bool AMAX, bool E5M2>
__global__ void __launch_bounds__(R::kThreads, R::kMinCtas)
codes_flat_kernel(const uint4 *__restrict__ x, uint32_t *__restrict__ y, size_t nvec,
                  const __grid_constant__ typename R::Params params, const float *__restrict__ scale,
                  float *__restrict__ amax_out)
{
    const R round(params, stage_table<R>(params));
    griddep_wait();  // PDL: the scale and the input may come from the kernel before this one
    griddep_launch_dependents();
    ScaleBf16 sc = {1.0f, 1.0f};
    if (scale) sc = load_scale<F32>(scale, 0);
    uint32_t amax = 0u;
    const int mode = classify_scale(sc.s);
    if (mode == DIV_UNIT)
        codes_span<R, F32, DIV_UNIT, AMAX, E5M2>(round, x, y, nvec, sc, amax);
    else if (F32 || mode == DIV_EXACT)
        codes_span<R, F32, DIV_EXACT, AMAX, E5M2>(round, x, y, nvec, sc, amax);
    else
        codes_span<R, F32, DIV_RECIP, AMAX, E5M2>(round, x, y, nvec, sc, amax);
    if (AMAX) block_amax_commit(amax, amax_out);
}

// tails and misaligned views
template <class R, bool F32, bool AMAX, bool E5M2>
__global__ void __launch_bounds__(R::kThreads)
codes_scalar_kernel(const void *__restrict__ xv, uint8_t *__restrict__ y, size_t first, size_t count,
                    const __grid_constant__ typename R::Params params, const float *__restrict__ scale,
                    float *__restrict__ amax_out)
{
    const R round(params, stage_table<R>(params));
    ScaleBf16 sc = {1.0f, 1.0f};
    if (scale) sc = load_scale<F32>(scale, 0);
    uint32_t amax = 0u;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = first + j;
        uint32_t q, ab;
        if (F32) {
            const uint32_t bits = static_cast<const uint32_t *>(xv)[i];
            ab = bits & 0x7FFFFFFFu;
            q = round(f32_to_bf16_rto_hi(__float_as_uint(__fdiv_rn(__uint_as_float(bits), sc.s))));
        } else {
            const uint32_t bits = (uint32_t) static_cast<const uint16_t *>(xv)[i] << 16;
            ab = bits & 0x7FFFFFFFu;
            q = round(bf16_rne_hi(bf16_quotient<DIV_EXACT>(bits, sc)));
        }
        y[i] = (uint8_t)(fp8x2_of<E5M2, R::kMxBand>(q, 0u) & 0xFFu);
        if (AMAX) amax = max(amax, ab);
    }
    if (AMAX) block_amax_commit(amax, amax_out);
}

// ----------------------------------------------------------------------------- scale update
__global__ void scale_update_kernel(float *__restrict__ history, int ahl, size_t channels, float *__restrict__ scale,
                                    float quant_max, int pow2)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    // history entries are max|x| bit patterns: non-negative or NaN, so an unsigned max is torch.amax
    uint32_t m = 0u;
    for (int i = 0; i < ahl; ++i) m = max(m, __float_as_uint(history[(size_t)i * channels + c]) & 0x7FFFFFFFu);
    const float amax = __uint_as_float(m);
    if (ahl > 1) {  // torch.roll(history, -1, 0)
        const float first = history[c];
        for (int i = 0; i + 1 < ahl; ++i) history[(size_t)i * channels + c] = history[(size_t)(i + 1) * channels + c];
        history[(size_t)(ahl - 1) * channels + c] = first;
    }
    history[c] = 0.0f;  // slot 0 <- amax of the current tensor, max-accumulated by the next kernel
    float sf = __fdiv_rn(amax, quant_max);
    const bool keep = (amax > 0.0f) && (m < 0x7F800000u);
    if (!keep) sf = scale[c];
    if (pow2) {
        const float p2 = qt_pow2_ceil(sf);
        sf = p2 >= 0.0f ? p2 : exp2f(ceilf(log2f(sf)));  // zero / subnormal / non-finite scales: libm semantics
    }
    scale[c] = sf;
}

// ----------------------------------------------------------------------------- launch plumbing

struct Job {
    const void *x;
    void *y;
    size_t outer, channels, inner;
    bool f32;
    const float *scale;
    float *amax;
    cudaStream_t stream;
};

template <class R, bool F32, bool AMAX, bool WRITE>
void launch_scalar(const Job &j, const typename R::Params &p, size_t first, size_t count)
{
    if (count == 0) return;
    auto kernel = fq_scalar_kernel<R, F32, AMAX, WRITE>;
    allow_smem<fq_scalar_kernel<R, F32, AMAX, WRITE>>(R::kSmemBytes);
    const unsigned grid = grid_for((count + R::kThreads - 1) / R::kThreads, R::kCtasPerSm * 2);
    kernel<<<grid, R::kThreads, R::kSmemBytes, j.stream>>>(j.x, j.y, first, count, j.channels, j.inner, p, j.scale,
                                                           j.amax);
}

template <class R, bool F32, bool AMAX, bool WRITE>
void launch_layout(const Job &j, const typename R::Params &p)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const size_t n = j.outer * j.channels * j.inner;
    const bool aligned = ((reinterpret_cast<uintptr_t>(j.x) | reinterpret_cast<uintptr_t>(j.y)) & 15u) == 0;
    const uint4 *xv = static_cast<const uint4 *>(j.x);
    uint4 *yv = static_cast<uint4 *>(j.y);

    if (j.channels == 1) {
        const size_t nvec = aligned ? n / VEC : 0;
        if (nvec) {
            if constexpr (WRITE) {
                auto kernel = fq_flat_kernel<R, F32, AMAX>;
                allow_smem<fq_flat_kernel<R, F32, AMAX>>(R::kSmemBytes);
                const size_t tile = (size_t)R::kThreads * kUnroll;
                const unsigned grid = grid_for((nvec + tile - 1) / tile, R::kCtasPerSm);
                qt_launch(kernel, dim3(grid), dim3(R::kThreads), R::kSmemBytes, j.stream, xv, yv, nvec, p, j.scale, j.amax);
            } else {
                const size_t tile = (size_t)256 * kUnroll;
                amax_flat_kernel<F32><<<grid_for((nvec + tile - 1) / tile, 8), 256, 0, j.stream>>>(xv, nvec, j.amax);
            }
        }
        launch_scalar<R, F32, AMAX, WRITE>(j, p, nvec * VEC, n - nvec * VEC);
        return;
    }
    // per channel (scale is never NULL here when WRITE)
    if (aligned && j.inner == 1 && j.channels % VEC == 0) {
        const size_t rows = j.outer, vec_per_row = j.channels / VEC;
        const unsigned gx = (unsigned)((vec_per_row + 31) / 32);
        size_t want_y = ((size_t)num_sms() * (R::kTable ? 2 : 8) + gx - 1) / gx;
        const size_t max_y = (rows + kColsY<R> - 1) / kColsY<R>;
        if (want_y > max_y) want_y = max_y;
        if (want_y < 1) want_y = 1;
        if (want_y > 65535) want_y = 65535;
        auto kernel = fq_cols_kernel<R, F32, AMAX, WRITE>;
        allow_smem<fq_cols_kernel<R, F32, AMAX, WRITE>>(R::kSmemBytes);
        kernel<<<dim3(gx, (unsigned)want_y), dim3(32, kColsY<R>), R::kSmemBytes, j.stream>>>(xv, yv, rows, vec_per_row, p,
                                                                                  j.scale, j.amax);
        return;
    }
    if (aligned && j.inner % VEC == 0 && j.inner >= 32 * VEC) {
        const size_t rows = j.outer * j.channels, vec_per_row = j.inner / VEC;
        const size_t seg_vecs = (size_t)32 * kUnroll * kRowIters;
        const size_t segs = (vec_per_row + seg_vecs - 1) / seg_vecs;
        const size_t warps_per_cta = R::kThreads / 32;
        auto kernel = fq_rows_kernel<R, F32, AMAX, WRITE>;
        allow_smem<fq_rows_kernel<R, F32, AMAX, WRITE>>(R::kSmemBytes);
        const unsigned grid = grid_for((rows * segs + warps_per_cta - 1) / warps_per_cta, R::kCtasPerSm);
        kernel<<<grid, R::kThreads, R::kSmemBytes, j.stream>>>(xv, yv, rows, j.channels, vec_per_row, segs, p, j.scale,
                                                               j.amax);
        return;
    }
    launch_scalar<R, F32, AMAX, WRITE>(j, p, 0, n);
}

template <class R>
void launch_flags(const Job &j, const typename R::Params &p)
{
    const bool amax = j.amax != nullptr;
    if (j.f32)
        amax ? launch_layout<R, true, true, true>(j, p) : launch_layout<R, true, false, true>(j, p);
    else
        amax ? launch_layout<R, false, true, true>(j, p) : launch_layout<R, false, false, true>(j, p);
}

template <int KIND>
void launch_direct(const Job &j, const QtRound &P)
{
    DirectParams<KIND> p;
    p.P = P;
    launch_flags<DirectRounder<KIND>>(j, p);
}

int check_layout(const char *fn, const void *x, size_t outer, size_t channels, size_t inner, int elem_type)
{
    if (elem_type != QT_BF16 && elem_type != QT_F32) {
        qt_set_error("%s: elem_type must be QT_BF16 or QT_F32, got %d", fn, elem_type);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (channels == 0) {
        qt_set_error("%s: channels must be >= 1", fn);
        return QT_ERR_INVALID_ARGUMENT;
    }
    const size_t esz = elem_type == QT_F32 ? 4 : 2;
    if (outer * channels * inner != 0 && (x == nullptr || (reinterpret_cast<uintptr_t>(x) % esz) != 0)) {
        qt_set_error("%s: x is NULL or not aligned to its element size", fn);
        return x ? QT_ERR_UNALIGNED : QT_ERR_INVALID_ARGUMENT;
    }
    return QT_OK;
}

}  // namespace

extern "C" int qt_fq_forward(const void *x, void *y, size_t outer, size_t channels, size_t inner, int elem_type,
                             const qt_format_t *fmt, const float *scale, float *amax_out, const void *lut,
                             void *stream)
{
    int rc = check_layout("qt_fq_forward", x, outer, channels, inner, elem_type);
    if (rc != QT_OK) return rc;
    if (!fmt) {
        qt_set_error("qt_fq_forward: fmt is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (channels > 1 && !scale) {
        qt_set_error("qt_fq_forward: per-channel call (channels=%zu) needs a scale array", channels);
        return QT_ERR_INVALID_ARGUMENT;
    }
    QtRound P;
    rc = qt_make_round(fmt, &P);
    if (rc != QT_OK) return rc;
    if (outer * channels * inner == 0) return QT_OK;
    if (!y) {
        qt_set_error("qt_fq_forward: y is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (num_sms() == 0) return no_device();
    Job j;
    j.x = x;
    j.y = y;
    j.outer = outer;
    j.channels = channels;
    j.inner = inner;
    j.f32 = elem_type == QT_F32;
    j.scale = scale;
    j.amax = amax_out;
    j.stream = static_cast<cudaStream_t>(stream);

    TableParams tp;
    if (lut && qt_lut_config(P, &tp.cfg) == QT_OK) {
        if (reinterpret_cast<uintptr_t>(lut) & 15u) {
            qt_set_error("qt_fq_forward: lut must be 16-byte aligned");
            return QT_ERR_UNALIGNED;
        }
        tp.table = static_cast<const QtLutEntry *>(lut);
        if (tp.cfg.mx_band)
            launch_flags<TableRounder<true, true>>(j, tp);
        else if (tp.cfg.clamp_bits != 0x7FFFFFFFu)
            launch_flags<TableRounder<true, false>>(j, tp);
        else
            launch_flags<TableRounder<false, false>>(j, tp);
    } else {
        switch (P.kind) {
        case QTR_IDENTITY: launch_direct<QTR_IDENTITY>(j, P); break;
        case QTR_INT: launch_direct<QTR_INT>(j, P); break;
        case QTR_FP_CUSTOM: launch_direct<QTR_FP_CUSTOM>(j, P); break;
        case QTR_FP_MX: launch_direct<QTR_FP_MX>(j, P); break;
        case QTR_POSIT: launch_direct<QTR_POSIT>(j, P); break;
        default: qt_set_error("bad format kind %d", P.kind); return QT_ERR_INVALID_ARGUMENT;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fake-quant kernel launch");
    return QT_OK;
}

extern "C" int qt_amax(const void *x, size_t outer, size_t channels, size_t inner, int elem_type, float *amax_out,
                       void *stream)
{
    int rc = check_layout("qt_amax", x, outer, channels, inner, elem_type);
    if (rc != QT_OK) return rc;
    if (!amax_out) {
        qt_set_error("qt_amax: amax_out is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (outer * channels * inner == 0) return QT_OK;
    if (num_sms() == 0) return no_device();
    Job j;
    j.x = x;
    j.y = nullptr;
    j.outer = outer;
    j.channels = channels;
    j.inner = inner;
    j.f32 = elem_type == QT_F32;
    j.scale = nullptr;
    j.amax = amax_out;
    j.stream = static_cast<cudaStream_t>(stream);
    DirectParams<QTR_IDENTITY> p;
    memset(&p, 0, sizeof(p));
    using R = DirectRounder<QTR_IDENTITY>;
    j.f32 ? launch_layout<R, true, true, false>(j, p) : launch_layout<R, false, true, false>(j, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "amax kernel launch");
    return QT_OK;
}

extern "C" int qt_scale_update(float *history, int amax_history_len, size_t channels, float *scale, float quant_max,
                               int force_scale_power_of_two, void *stream)
{
    if (!history || !scale || amax_history_len < 1 || channels == 0) {
        qt_set_error("qt_scale_update: invalid argument (history=%p scale=%p ahl=%d channels=%zu)", (void *)history,
                     (void *)scale, amax_history_len, channels);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (num_sms() == 0) return no_device();
    const unsigned grid = (unsigned)((channels + 127) / 128);
    scale_update_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(history, amax_history_len, channels, scale,
                                                                             quant_max, force_scale_power_of_two);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "scale_update kernel launch");
    return QT_OK;
}

namespace {
template <class R, bool F32, bool AMAX, bool E5M2>
void launch_codes(const void *x, void *y, size_t n, const typename R::Params &p, const float *scale, float *amax,
                  cudaStream_t stream)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0;
    const size_t nvec = aligned ? n / VEC : 0;
    if (nvec) {
        allow_smem<codes_flat_kernel<R, F32, AMAX, E5M2>>(R::kSmemBytes);
        const size_t tile = (size_t)R::kThreads * kUnroll;
        const unsigned grid = grid_for((nvec + tile - 1) / tile, R::kCtasPerSm);
        qt_launch(codes_flat_kernel<R, F32, AMAX, E5M2>, dim3(grid), dim3(R::kThreads), R::kSmemBytes, stream,
                  static_cast<const uint4 *>(x), static_cast<uint32_t *>(y), nvec, p, scale, amax);
    }
    const size_t rest = n - nvec * VEC;
    if (rest) {
        allow_smem<codes_scalar_kernel<R, F32, AMAX, E5M2>>(R::kSmemBytes);
        const unsigned grid = grid_for((rest + R::kThreads - 1) / R::kThreads, R::kCtasPerSm * 2);
        codes_scalar_kernel<R, F32, AMAX, E5M2><<<grid, R::kThreads, R::kSmemBytes, stream>>>(
            x, static_cast<uint8_t *>(y), nvec * VEC, rest, p, scale, amax);
    }
}
template <class R, bool E5M2>
void launch_codes_flags(const void *x, void *y, size_t n, bool f32, const typename R::Params &p, const float *scale,
                        float *amax, cudaStream_t stream)
{
    if (f32)
        amax ? launch_codes<R, true, true, E5M2>(x, y, n, p, scale, amax, stream)
             : launch_codes<R, true, false, E5M2>(x, y, n, p, scale, amax, stream);
    else
        amax ? launch_codes<R, false, true, E5M2>(x, y, n, p, scale, amax, stream)
             : launch_codes<R, false, false, E5M2>(x, y, n, p, scale, amax, stream);
}
}  // namespace

extern "C" int qt_quantize_codes(const void *x, void *codes, size_t n, int elem_type, const qt_format_t *fmt,
                                 const float *scale, float *amax_out, const void *lut, void *stream)
{
    int rc = check_layout("qt_quantize_codes", x, 1, 1, n, elem_type);
    if (rc != QT_OK) return rc;
    if (!fmt || (n && !codes)) {
        qt_set_error("qt_quantize_codes: NULL argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const bool e4m3 = fmt->kind == QT_KIND_FP && fmt->ebits == 4 && fmt->mbits == 3 && !fmt->is_unsigned;
    const bool e5m2 = fmt->kind == QT_KIND_FP && fmt->ebits == 5 && fmt->mbits == 2 && !fmt->is_unsigned;
    if (!e4m3 && !e5m2) {
        qt_set_error("qt_quantize_codes: one-byte codes exist for e4m3 / e5m2 / fp8_e4m3 / fp8_e5m2 only");
        return QT_ERR_UNSUPPORTED_DTYPE;
    }
    QtRound P;
    rc = qt_make_round(fmt, &P);
    if (rc != QT_OK) return rc;
    if (n == 0) return QT_OK;
    if (num_sms() == 0) return no_device();
    TableParams tp;
    if (!lut || qt_lut_config(P, &tp.cfg) != QT_OK || (reinterpret_cast<uintptr_t>(lut) & 15u)) {
        qt_set_error("qt_quantize_codes: needs the 16-byte aligned device table from qt_lut_build_host(fmt)");
        return QT_ERR_INVALID_ARGUMENT;
    }
    tp.table = static_cast<const QtLutEntry *>(lut);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool f32 = elem_type == QT_F32;
    if (tp.cfg.mx_band)
        e5m2 ? launch_codes_flags<TableRounder<true, true>, true>(x, codes, n, f32, tp, scale, amax_out, st)
             : launch_codes_flags<TableRounder<true, true>, false>(x, codes, n, f32, tp, scale, amax_out, st);
    else
        e5m2 ? launch_codes_flags<TableRounder<true, false>, true>(x, codes, n, f32, tp, scale, amax_out, st)
             : launch_codes_flags<TableRounder<true, false>, false>(x, codes, n, f32, tp, scale, amax_out, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "quantize-to-codes kernel launch");
    return QT_OK;
}
