// qt_block.cu -- block-scaled fake quant for sm_100a: the `microscaling` and `group_wise_affine` qschemes.
//
// Replaces (reference, src/quantized_training/): MXFakeQuantFunction.forward fake_quantize.py:105-129,
// GroupWiseAffineFakeQuantFunction.forward :138-190, calculate_mx_qparam decomposed.py:372-419, quantize :171-210,
// expand :127-140 and the pad / reshape of mx_utils.py:62-121 -- about 25 ATen launches with padded, tiled and
// repeat_interleave-d temporaries -- with ONE pass over HBM: the block statistic, the scale, (x / s) -> round to the
// format -> (* s) and the store happen while the block sits in registers.
//
// Unlike the per-tensor scheme (qt_fq.cu) the scale is NOT delayed: it is a function of the block being quantized,
// so the statistic has to be complete before the first element is rounded.  A block is small (block_size elements,
// or block_size^2 for two tiled axes), which is what makes the single pass possible:
//   (the three single-pass families live in qt_block_flat.cu / _cols.cu / _tile.cu so that they compile in parallel)
//   flat  kernel  blocks along the last axis (unit stride): block_size / VEC lanes of a warp hold one block, the
//                 maximum is a log2(lanes)-step xor-shuffle.
//   cols  kernel  blocks along an inner axis (stride = inner elements): a (32, 8) CTA holds 32 column groups of
//                 one block row; every column is its own block, maxima of the 8 row phases meet in shared memory.
//   tile  kernel  two tiled axes (the last two): the cols geometry with a block bs / VEC lanes wide.
//   generic       anything else (ragged last axis, odd block sizes, misaligned views, non-adjacent tiled axes):
//                 a statistic kernel (one warp per block) and an element-wise apply kernel.
//
// Scale arithmetic follows the reference op by op in the tensor's dtype (bf16 tensors: every intermediate is
// rounded to bf16; python scalars enter as fp32).  force_scale_power_of_two evaluates floor(log2(amax)) *in the
// tensor's dtype* (mx_utils.py:44-48): for a bf16 tensor log2() is rounded to 8 significant bits before the floor,
// so e.g. amax = 1.98 * 2^60 has floor(bf16(60.98)) = 61.  That function of amax is one mantissa threshold per
// exponent, tabulated on the host (qt_block_pow2_table_host) with the same libm call sequence and checked against
// the reference on every bf16 value.
#include "qt_block_common.cuh"

namespace {

// ----------------------------------------------------------------------------- affine scheme, single pass
// sf = (max - min) / (quant_max - quant_min); sf = where(sf > 0, sf, 1); zp = -min / sf + quant_min, each op in
// the tensor's dtype, both optionally through the scale codebook (fake_quantize.py:165-174).  NaN in the block:
// amin / amax propagate it.
template <bool F32>
__device__ __forceinline__ void gwa_params(float mn, float mx, bool nan, const BlockParams &bp, float &sf, float &zp)
{
    if (nan) mn = mx = __uint_as_float(QT_NAN_BITS);
    sf = to_dtype<F32>(__fdiv_rn(to_dtype<F32>(__fsub_rn(mx, mn)), bp.range));
    sf = sf > 0.0f ? sf : 1.0f;
    zp = to_dtype<F32>(__fadd_rn(to_dtype<F32>(__fdiv_rn(-mn, sf)), bp.quant_min));
    if (bp.has_scale_fmt) {
        sf = scale_codebook<F32>(bp, sf);
        zp = scale_codebook<F32>(bp, zp);
    }
}
// q = clamp(round(x / sf + zp), qmin, qmax); y = (q - zp) * sf -- every op rounded to the tensor's dtype
template <bool F32>
__device__ __forceinline__ float gwa_elem(float x, float sf, float zp, const BlockParams &bp)
{
    float q = to_dtype<F32>(__fadd_rn(to_dtype<F32>(__fdiv_rn(x, sf)), zp));
    q = rintf(q);
    q = (q < bp.quant_min) ? bp.quant_min : q;  // compare-select keeps NaN like torch.clamp
    q = (q > bp.quant_max) ? bp.quant_max : q;
    return to_dtype<F32>(__fmul_rn(to_dtype<F32>(__fsub_rn(q, zp)), sf));
}
// x / sf as a reciprocal multiply is legal when sf is a normal-range bf16 value, the quotients stay finite, and a
// sub-2^-120 quotient cannot matter: it is absorbed by a zero point of ordinary size, or (zp == 0) only its sign
// survives round().
__device__ __forceinline__ bool gwa_block_is_fast(uint32_t amax_bits, float sf, float rsf, float zp)
{
    const uint32_t sb = __float_as_uint(sf), zb = __float_as_uint(zp) & 0x7FFFFFFFu;
    return amax_bits < 0x7E800000u && (sb - 0x0D800000u) <= (0x71800000u - 0x0D800000u) &&
           (zb == 0u || (zb - 0x0D800000u) <= (0x71800000u - 0x0D800000u)) &&
           __fmul_rn(__uint_as_float(amax_bits), rsf) < 0x1p126f;
}
// two bf16 values of one word, each with its own parameters (fast path: no NaN, bounded quotients)
__device__ __forceinline__ uint32_t gwa_word_fast(uint32_t w, float sf_lo, float rsf_lo, float zp_lo, float sf_hi,
                                                 float rsf_hi, float zp_hi, const BlockParams &bp)
{
    uint32_t u = bf16x2_rne(__fmul_rn(__uint_as_float(w << 16), rsf_lo),
                            __fmul_rn(__uint_as_float(w & 0xFFFF0000u), rsf_hi));
    u = bf16x2_rne(__fadd_rn(__uint_as_float(u << 16), zp_lo), __fadd_rn(__uint_as_float(u & 0xFFFF0000u), zp_hi));
    float qlo = rintf(__uint_as_float(u << 16)), qhi = rintf(__uint_as_float(u & 0xFFFF0000u));
    qlo = fminf(fmaxf(qlo, bp.quant_min), bp.quant_max);
    qhi = fminf(fmaxf(qhi, bp.quant_min), bp.quant_max);
    u = bf16x2_rne(__fsub_rn(qlo, zp_lo), __fsub_rn(qhi, zp_hi));
    return bf16x2_rne(__fmul_rn(__uint_as_float(u << 16), sf_lo), __fmul_rn(__uint_as_float(u & 0xFFFF0000u), sf_hi));
}
// fminf / fmaxf may return either zero for (-0, +0), unlike the compare-select clamp of the careful path.  The sign of
// a zero q matters only when it survives (q - zp), i.e. when zp == 0 and a clamp bound is zero: careful path then.
__device__ __forceinline__ bool gwa_fast_allowed(float zp, const BlockParams &bp)
{
    return !(zp == 0.0f && (bp.quant_min == 0.0f || bp.quant_max == 0.0f));
}

__device__ __forceinline__ uint32_t bf16x2_min(uint32_t a, uint32_t b)
{
    __nv_bfloat162 r = __hmin2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b)
{
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}

// Blocks of LANES consecutive 16-byte vectors along the last axis (same geometry as mx_flat_kernel).
template <bool F32, int LANES>
__global__ void __launch_bounds__(256, 4)
gwa_flat_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t nvec, const __grid_constant__ BlockParams bp,
                float *__restrict__ scale_out, float *__restrict__ zp_out)
{
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
            uint32_t a = F32 ? amax_of_vec_f32(0u, v[j]) : amax_of_vec_bf16(0u, v[j]);
            float mn, mx;
            if (F32) {
                mn = fminf(fminf(__uint_as_float(v[j].x), __uint_as_float(v[j].y)),
                           fminf(__uint_as_float(v[j].z), __uint_as_float(v[j].w)));
                mx = fmaxf(fmaxf(__uint_as_float(v[j].x), __uint_as_float(v[j].y)),
                           fmaxf(__uint_as_float(v[j].z), __uint_as_float(v[j].w)));
            } else {
                const uint32_t pmn = bf16x2_min(bf16x2_min(v[j].x, v[j].y), bf16x2_min(v[j].z, v[j].w));
                const uint32_t pmx = bf16x2_max(bf16x2_max(v[j].x, v[j].y), bf16x2_max(v[j].z, v[j].w));
                mn = fminf(__uint_as_float(pmn << 16), __uint_as_float(pmn & 0xFFFF0000u));
                mx = fmaxf(__uint_as_float(pmx << 16), __uint_as_float(pmx & 0xFFFF0000u));
            }
#pragma unroll
            for (int o = 1; o < LANES; o <<= 1) {
                a = max(a, __shfl_xor_sync(0xFFFFFFFFu, a, o));
                mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
            }
            float sf, zp;
            gwa_params<F32>(mn, mx, a > 0x7F800000u, bp, sf, zp);
            const float rsf = __frcp_rn(sf);
            const bool fast = !F32 && __all_sync(0xFFFFFFFFu, gwa_block_is_fast(a, sf, rsf, zp) && gwa_fast_allowed(zp, bp));
            uint4 r;
            if (fast) {
                r.x = gwa_word_fast(v[j].x, sf, rsf, zp, sf, rsf, zp, bp);
                r.y = gwa_word_fast(v[j].y, sf, rsf, zp, sf, rsf, zp, bp);
                r.z = gwa_word_fast(v[j].z, sf, rsf, zp, sf, rsf, zp, bp);
                r.w = gwa_word_fast(v[j].w, sf, rsf, zp, sf, rsf, zp, bp);
            } else {
                const uint32_t win[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                uint32_t out[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (F32) {
                        out[q] = __float_as_uint(gwa_elem<true>(__uint_as_float(win[q]), sf, zp, bp));
                    } else {
                        const float lo = gwa_elem<false>(__uint_as_float(win[q] << 16), sf, zp, bp);
                        const float hi = gwa_elem<false>(__uint_as_float(win[q] & 0xFFFF0000u), sf, zp, bp);
                        out[q] = __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
                    }
                }
                r = make_uint4(out[0], out[1], out[2], out[3]);
            }
            if (i < nvec) {
                if ((i & (size_t)(LANES - 1)) == 0) {
                    scale_out[i / LANES] = sf;
                    zp_out[i / LANES] = zp;
                }
                st_stream(y + i, r);
            }
        }
    }
}

// Blocks along an inner axis (same geometry as mx_cols_kernel): rows past n are the reference's zero padding and
// DO enter min / max.
template <bool F32, int RPT>
__global__ void __launch_bounds__(256, RPT <= 4 ? 4 : (RPT <= 8 ? 3 : 2))
gwa_cols_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t outer, size_t n, size_t inner_vec,
                size_t nblk, const __grid_constant__ BlockParams bp, float *__restrict__ scale_out,
                float *__restrict__ zp_out)
{
    constexpr int VEC = F32 ? 4 : 8;
    constexpr int BS = 8 * RPT;
    __shared__ uint4 red_mn[8][33], red_mx[8][33], red_am[8][33];
    __shared__ float col_sf[32][VEC + 1], col_zp[32][VEC + 1];
    __shared__ unsigned char col_fast[32][VEC];
    const size_t G = outer * inner_vec;
    const size_t gchunks = (G + 31) / 32;
    const size_t work = nblk * gchunks;
    for (size_t w = blockIdx.x; w < work; w += gridDim.x) {
        const size_t b = w / gchunks, gc = w - b * gchunks;
        const size_t g = gc * 32 + threadIdx.x;
        const bool active = g < G;
        const size_t o = active ? g / inner_vec : 0, cv = active ? g - o * inner_vec : 0;
        const size_t row0 = b * BS + threadIdx.y;
        const uint4 *xp = x + (o * n) * inner_vec + cv;
        uint4 v[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const size_t r = row0 + (size_t)k * 8;
            v[k] = (active && r < n) ? ld_stream(xp + r * inner_vec) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint4 mn = v[0], mx = v[0], am = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            if (F32) {
                mn.x = __float_as_uint(fminf(__uint_as_float(mn.x), __uint_as_float(v[k].x)));
                mn.y = __float_as_uint(fminf(__uint_as_float(mn.y), __uint_as_float(v[k].y)));
                mn.z = __float_as_uint(fminf(__uint_as_float(mn.z), __uint_as_float(v[k].z)));
                mn.w = __float_as_uint(fminf(__uint_as_float(mn.w), __uint_as_float(v[k].w)));
                mx.x = __float_as_uint(fmaxf(__uint_as_float(mx.x), __uint_as_float(v[k].x)));
                mx.y = __float_as_uint(fmaxf(__uint_as_float(mx.y), __uint_as_float(v[k].y)));
                mx.z = __float_as_uint(fmaxf(__uint_as_float(mx.z), __uint_as_float(v[k].z)));
                mx.w = __float_as_uint(fmaxf(__uint_as_float(mx.w), __uint_as_float(v[k].w)));
                am.x = max(am.x, v[k].x & 0x7FFFFFFFu);
                am.y = max(am.y, v[k].y & 0x7FFFFFFFu);
                am.z = max(am.z, v[k].z & 0x7FFFFFFFu);
                am.w = max(am.w, v[k].w & 0x7FFFFFFFu);
            } else {
                mn.x = bf16x2_min(mn.x, v[k].x); mn.y = bf16x2_min(mn.y, v[k].y);
                mn.z = bf16x2_min(mn.z, v[k].z); mn.w = bf16x2_min(mn.w, v[k].w);
                mx.x = bf16x2_max(mx.x, v[k].x); mx.y = bf16x2_max(mx.y, v[k].y);
                mx.z = bf16x2_max(mx.z, v[k].z); mx.w = bf16x2_max(mx.w, v[k].w);
                am.x = __vmaxu2(am.x, v[k].x & 0x7FFF7FFFu); am.y = __vmaxu2(am.y, v[k].y & 0x7FFF7FFFu);
                am.z = __vmaxu2(am.z, v[k].z & 0x7FFF7FFFu); am.w = __vmaxu2(am.w, v[k].w & 0x7FFF7FFFu);
            }
        }
        red_mn[threadIdx.y][threadIdx.x] = mn;
        red_mx[threadIdx.y][threadIdx.x] = mx;
        red_am[threadIdx.y][threadIdx.x] = am;
        __syncthreads();
        if (threadIdx.y < VEC) {
            const int c = threadIdx.y;
            float fmn = __uint_as_float(0x7F800000u), fmx = __uint_as_float(0xFF800000u);
            uint32_t a = 0u;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const uint4 qn = red_mn[p][threadIdx.x], qx = red_mx[p][threadIdx.x], qa = red_am[p][threadIdx.x];
                const int wi = F32 ? c : (c >> 1);
                const uint32_t wn = wi == 0 ? qn.x : wi == 1 ? qn.y : wi == 2 ? qn.z : qn.w;
                const uint32_t wx = wi == 0 ? qx.x : wi == 1 ? qx.y : wi == 2 ? qx.z : qx.w;
                const uint32_t wa = wi == 0 ? qa.x : wi == 1 ? qa.y : wi == 2 ? qa.z : qa.w;
                const bool hi = !F32 && (c & 1);
                fmn = fminf(fmn, __uint_as_float(F32 ? wn : (hi ? (wn & 0xFFFF0000u) : (wn << 16))));
                fmx = fmaxf(fmx, __uint_as_float(F32 ? wx : (hi ? (wx & 0xFFFF0000u) : (wx << 16))));
                a = max(a, F32 ? wa : (hi ? (wa & 0xFFFF0000u) : (wa << 16)));
            }
            float sf, zp;
            gwa_params<F32>(fmn, fmx, a > 0x7F800000u, bp, sf, zp);
            if (active) {
                const size_t si = (o * nblk + b) * (inner_vec * VEC) + cv * VEC + c;
                scale_out[si] = sf;
                zp_out[si] = zp;
            }
            col_sf[threadIdx.x][c] = sf;
            col_zp[threadIdx.x][c] = zp;
            col_fast[threadIdx.x][c] =
                !F32 && gwa_block_is_fast(a, sf, __frcp_rn(sf), zp) && gwa_fast_allowed(zp, bp) ? 1 : 0;
        }
        __syncthreads();
        float sf[VEC], rsf[VEC], zp[VEC];
        bool mine_fast = !F32;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            mine_fast = mine_fast && col_fast[threadIdx.x][c] != 0;
            sf[c] = col_sf[threadIdx.x][c];
            zp[c] = col_zp[threadIdx.x][c];
        }
        const bool fast = __all_sync(0xFFFFFFFFu, mine_fast);
        uint4 *yp = y + (o * n) * inner_vec + cv;
#pragma unroll
        for (int c = 0; c < VEC; ++c) rsf[c] = __frcp_rn(sf[c]);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const size_t r = row0 + (size_t)k * 8;
            const uint32_t win[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (F32) {
                    out[q] = __float_as_uint(gwa_elem<true>(__uint_as_float(win[q]), sf[q % VEC], zp[q % VEC], bp));
                } else if (fast) {
                    out[q] = gwa_word_fast(win[q], sf[(2 * q) % VEC], rsf[(2 * q) % VEC], zp[(2 * q) % VEC],
                                           sf[(2 * q + 1) % VEC], rsf[(2 * q + 1) % VEC], zp[(2 * q + 1) % VEC], bp);
                } else {
                    const float lo = gwa_elem<false>(__uint_as_float(win[q] << 16), sf[(2 * q) % VEC],
                                                     zp[(2 * q) % VEC], bp);
                    const float hi = gwa_elem<false>(__uint_as_float(win[q] & 0xFFFF0000u),
                                                     sf[(2 * q + 1) % VEC], zp[(2 * q + 1) % VEC], bp);
                    out[q] = __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
                }
            }
            if (active && r < n) st_stream(yp + r * inner_vec, make_uint4(out[0], out[1], out[2], out[3]));
        }
    }
}

// ----------------------------------------------------------------------------- generic kernels
// Tensor [d0, n1, d1, n2, d2]; n1 is tiled with bs, n2 is tiled with bs2 (bs or 1).  Block grid
// [d0, nb1, d1, nb2, d2] row-major is the layout of scale / zero_point.
// AFFINE = false: amax -> scale.  AFFINE = true: min / max (padding zeros of cut blocks included) -> scale, zp.
template <bool F32, bool AFFINE>
__global__ void __launch_bounds__(256)
block_stat_kernel(const void *__restrict__ x, const __grid_constant__ BlockDims D,
                  const __grid_constant__ BlockParams bp, float *__restrict__ scale_out,
                  float *__restrict__ zp_out)
{
    const size_t nblocks = D.d0 * D.nb1 * D.d1 * D.nb2 * D.d2;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const int lane = threadIdx.x & 31;
    const uint32_t per_block = D.bs * D.bs2;
    for (size_t bi = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bi < nblocks; bi += nwarps) {
        size_t r = bi;
        const size_t i4 = r % D.d2; r /= D.d2;
        const size_t b2 = r % D.nb2; r /= D.nb2;
        const size_t i2 = r % D.d1; r /= D.d1;
        const size_t b1 = r % D.nb1;
        const size_t i0 = r / D.nb1;
        uint32_t amax = 0u;
        float mn = __uint_as_float(0x7F800000u), mx = __uint_as_float(0xFF800000u);
        bool nan = false, cut = false;
        for (uint32_t e = lane; e < per_block; e += 32) {
            const size_t j1 = b1 * D.bs + e / D.bs2, j2 = b2 * D.bs2 + e % D.bs2;
            if (j1 < D.n1 && j2 < D.n2) {
                const float v = load_elem(x, F32, (((i0 * D.n1 + j1) * D.d1 + i2) * D.n2 + j2) * D.d2 + i4);
                if (AFFINE) {
                    nan |= v != v;
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                } else {
                    amax = max(amax, __float_as_uint(v) & 0x7FFFFFFFu);
                }
            } else {
                cut = true;
            }
        }
        if (AFFINE) {
            if (cut) {
                mn = fminf(mn, 0.0f);
                mx = fmaxf(mx, 0.0f);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
            }
            nan = __any_sync(0xFFFFFFFFu, nan);
            if (lane == 0) {
                float sf, zp;
                gwa_params<F32>(mn, mx, nan, bp, sf, zp);
                scale_out[bi] = sf;
                zp_out[bi] = zp;
            }
        } else {
            amax = __reduce_max_sync(0xFFFFFFFFu, amax);
            if (lane == 0) scale_out[bi] = mx_scale_of<F32>(amax, bp);
        }
    }
}

template <class R, bool F32, bool AFFINE>
__global__ void __launch_bounds__(R::kThreads)
block_apply_kernel(const void *__restrict__ x, void *__restrict__ y, size_t total,
                   const __grid_constant__ BlockDims D, const __grid_constant__ typename R::Params params,
                   const __grid_constant__ BlockParams bp, const float *__restrict__ scale,
                   const float *__restrict__ zp)
{
    const R round(params, stage_table<R>(params));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        size_t r = i;
        const size_t i4 = r % D.d2; r /= D.d2;
        const size_t j2 = r % D.n2; r /= D.n2;
        const size_t i2 = r % D.d1; r /= D.d1;
        const size_t j1 = r % D.n1;
        const size_t i0 = r / D.n1;
        const size_t bi = (((i0 * D.nb1 + j1 / D.bs) * D.d1 + i2) * D.nb2 + j2 / D.bs2) * D.d2 + i4;
        const float s = scale[bi];
        uint32_t out;
        if (AFFINE) {
            // q = clamp(round(x / sf + zp), qmin, qmax); y = (q - zp) * sf -- every op rounded to the tensor's dtype
            out = __float_as_uint(gwa_elem<F32>(load_elem(x, F32, i), s, zp[bi], bp));
        } else if (F32) {
            out = fq_f32<R, false>(round, static_cast<const uint32_t *>(x)[i], s);
        } else {
            out = fq_bf16<R, DIV_EXACT>(round, (uint32_t) static_cast<const uint16_t *>(x)[i] << 16, make_scale(s));
        }
        if (F32)
            static_cast<uint32_t *>(y)[i] = out;
        else
            static_cast<uint16_t *>(y)[i] = (uint16_t)(out >> 16);
    }
}

// ----------------------------------------------------------------------------- launch

template <class R, bool F32, bool AFFINE>
void launch_generic(const BlockJob &j, const typename R::Params &p)
{
    const BlockDims &D = j.D;
    const size_t nblocks = D.d0 * D.nb1 * D.d1 * D.nb2 * D.d2;
    const size_t total = D.d0 * D.n1 * D.d1 * D.n2 * D.d2;
    block_stat_kernel<F32, AFFINE><<<grid_for((nblocks + 7) / 8, 8), 256, 0, j.stream>>>(j.d->x, D, j.bp, j.d->scale,
                                                                                      j.d->zero_point);
    if (!j.d->y) return;  // parameters only
    allow_smem<block_apply_kernel<R, F32, AFFINE>>(R::kSmemBytes);
    const unsigned grid = grid_for((total + R::kThreads - 1) / R::kThreads, R::kCtasPerSm);
    block_apply_kernel<R, F32, AFFINE><<<grid, R::kThreads, R::kSmemBytes, j.stream>>>(
        j.d->x, j.d->y, total, D, p, j.bp, j.d->scale, j.d->zero_point);
}

template <bool F32, int LANES>
void launch_gwa_flat(const BlockJob &j, size_t nvec)
{
    const size_t tile = (size_t)256 * kUnroll;
    gwa_flat_kernel<F32, LANES><<<grid_for((nvec + tile - 1) / tile, 4), 256, 0, j.stream>>>(
        static_cast<const uint4 *>(j.d->x), static_cast<uint4 *>(j.d->y), nvec, j.bp, j.d->scale, j.d->zero_point);
}
template <bool F32, int RPT>
void launch_gwa_cols(const BlockJob &j, size_t outer, size_t n, size_t inner_vec)
{
    const size_t nblk = (n + 8 * RPT - 1) / (8 * RPT);
    const size_t work = nblk * ((outer * inner_vec + 31) / 32);
    gwa_cols_kernel<F32, RPT><<<grid_for(work, 6), dim3(32, 8), 0, j.stream>>>(
        static_cast<const uint4 *>(j.d->x), static_cast<uint4 *>(j.d->y), outer, n, inner_vec, nblk, j.bp, j.d->scale,
        j.d->zero_point);
}
// single-pass kernels for the affine scheme; false = not expressible, use the generic pair
template <bool F32>
bool try_gwa_fast(const BlockJob &j)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const BlockDims &D = j.D;
    if ((reinterpret_cast<uintptr_t>(j.d->x) | reinterpret_cast<uintptr_t>(j.d->y)) & 15u) return false;
    if (D.n2 != 1 || D.d2 != 1 || D.bs2 != 1) return false;
    if (D.d1 == 1) {
        if (D.n1 % D.bs != 0 || D.bs % VEC != 0) return false;
        const size_t nvec = D.d0 * D.n1 / VEC;
        switch (D.bs / VEC) {
        case 1: launch_gwa_flat<F32, 1>(j, nvec); return true;
        case 2: launch_gwa_flat<F32, 2>(j, nvec); return true;
        case 4: launch_gwa_flat<F32, 4>(j, nvec); return true;
        case 8: launch_gwa_flat<F32, 8>(j, nvec); return true;
        case 16: launch_gwa_flat<F32, 16>(j, nvec); return true;
        case 32: launch_gwa_flat<F32, 32>(j, nvec); return true;
        default: return false;
        }
    }
    if (D.d1 < VEC || D.d1 % VEC != 0) return false;
    const size_t inner_vec = D.d1 / VEC;
    switch (D.bs) {
    case 8: launch_gwa_cols<F32, 1>(j, D.d0, D.n1, inner_vec); return true;
    case 16: launch_gwa_cols<F32, 2>(j, D.d0, D.n1, inner_vec); return true;
    case 32: launch_gwa_cols<F32, 4>(j, D.d0, D.n1, inner_vec); return true;
    case 64: launch_gwa_cols<F32, 8>(j, D.d0, D.n1, inner_vec); return true;
    case 128: launch_gwa_cols<F32, 16>(j, D.d0, D.n1, inner_vec); return true;
    default: return false;
    }
}


// ----------------------------------------------------------------------------- table ops (operator surface)
// torch.ops.quantized_ops.{vmap, quantize, dequantize} of the reference (decomposed.py:143-262) with a caller-supplied
// 65 536-entry table (any codebook, e.g. one this library has no bitwise rounder for):
//   QUANT    u = x / s [+ zp];  y = table ? table[idx(u)] : u
//   DEQUANT  v = table_in ? table_in[idx(x)] : x;  d = zp ? (v - zp) * s : v * s;  y = table_out ? table_out[idx(d)] : d
// idx() = the bf16 bits (fp32: truncated with round-to-odd); s / zp per block of the grid `D` (or one value).
// The 128 KB table is read through L1 / L2 (a gather; the formats of this library never need it).
__device__ __forceinline__ float table_lookup(const uint16_t *table, bool f32, float v)
{
    if (!table) return v;
    const uint32_t b = __float_as_uint(v);
    const uint32_t idx = f32 ? (f32_to_bf16_rto_hi(b) >> 16) : (b >> 16);
    return __uint_as_float((uint32_t)__ldg(table + idx) << 16);
}

// One arithmetic result in the tensor's dtype with the NaN the reference's CPU run produces, so that even a table
// that distinguishes NaN patterns is indexed identically: bf16 tensors -- c10::BFloat16's conversion turns every
// NaN into 0x7FC0; fp32 tensors -- SSE semantics: the first NaN operand, quieted, else the default NaN 0xFFC00000.
template <bool F32>
__device__ __forceinline__ float op_result(float a, float b, float r)
{
    if (r == r) return to_dtype<F32>(r);
    if (!F32) return __uint_as_float(0x7FC00000u);
    if (a != a) return __uint_as_float(__float_as_uint(a) | 0x00400000u);
    if (b != b) return __uint_as_float(__float_as_uint(b) | 0x00400000u);
    return __uint_as_float(0xFFC00000u);
}

template <bool F32, int OP>
__global__ void __launch_bounds__(256)
table_op_kernel(const void *__restrict__ x, void *__restrict__ y, size_t total, const __grid_constant__ BlockDims D,
                const void *__restrict__ scale, const void *__restrict__ zp, int scalar_params,
                const uint16_t *__restrict__ table_a, const uint16_t *__restrict__ table_b)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float v = load_elem(x, F32, i);
        float out;
        if (OP == QT_TABLE_LOOKUP) {
            out = table_lookup(table_a, F32, v);
        } else {
            size_t bi = 0;
            if (!scalar_params) {
                size_t r = i;
                const size_t i4 = r % D.d2; r /= D.d2;
                const size_t j2 = r % D.n2; r /= D.n2;
                const size_t i2 = r % D.d1; r /= D.d1;
                const size_t j1 = r % D.n1;
                const size_t i0 = r / D.n1;
                bi = (((i0 * D.nb1 + j1 / D.bs) * D.d1 + i2) * D.nb2 + j2 / D.bs2) * D.d2 + i4;
            }
            const float s = load_elem(scale, F32, bi);
            if (OP == QT_TABLE_DEQUANTIZE) {
                float d = table_lookup(table_a, F32, v);
                if (zp) {
                    const float z = load_elem(zp, F32, bi);
                    d = op_result<F32>(d, z, __fsub_rn(d, z));
                }
                d = op_result<F32>(d, s, __fmul_rn(d, s));
                out = table_lookup(table_b, F32, d);
            } else {
                float u = op_result<F32>(v, s, __fdiv_rn(v, s));
                if (zp) {
                    const float z = load_elem(zp, F32, bi);
                    u = op_result<F32>(u, z, __fadd_rn(u, z));
                }
                out = table_lookup(table_a, F32, u);
            }
        }
        if (F32)
            static_cast<float *>(y)[i] = out;
        else
            static_cast<uint16_t *>(y)[i] = (uint16_t)(__float_as_uint(out) >> 16);
    }
}

}  // namespace

extern "C" int qt_fq_block(const qt_block_desc_t *d, void *stream)
{
    if (!d) {
        qt_set_error("qt_fq_block: desc is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (d->elem_type != QT_BF16 && d->elem_type != QT_F32) {
        qt_set_error("qt_fq_block: elem_type must be QT_BF16 or QT_F32, got %d", d->elem_type);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (d->qscheme != QT_BLOCK_MX && d->qscheme != QT_BLOCK_AFFINE) {
        qt_set_error("qt_fq_block: qscheme must be QT_BLOCK_MX or QT_BLOCK_AFFINE, got %d", d->qscheme);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (d->block_size < 1) {
        qt_set_error("qt_fq_block: block_size must be >= 1, got %d", d->block_size);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (d->d0 < 0 || d->n1 < 0 || d->d1 < 0 || d->n2 < 0 || d->d2 < 0) {
        qt_set_error("qt_fq_block: negative dimension");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const bool affine = d->qscheme == QT_BLOCK_AFFINE;
    BlockJob j;
    j.d = d;
    j.stream = static_cast<cudaStream_t>(stream);
    j.f32 = d->elem_type == QT_F32;
    BlockDims &D = j.D;
    D.d0 = (size_t)d->d0; D.n1 = (size_t)d->n1; D.d1 = (size_t)d->d1; D.n2 = (size_t)d->n2; D.d2 = (size_t)d->d2;
    D.bs = (uint32_t)d->block_size;
    D.bs2 = d->block_axis2 ? D.bs : 1u;
    D.nb1 = (D.n1 + D.bs - 1) / D.bs;
    D.nb2 = (D.n2 + D.bs2 - 1) / D.bs2;
    const size_t total = D.d0 * D.n1 * D.d1 * D.n2 * D.d2;

    QtRound P;
    memset(&P, 0, sizeof(P));
    if (!affine) {
        if (!d->fmt) {
            qt_set_error("qt_fq_block: fmt is NULL");
            return QT_ERR_INVALID_ARGUMENT;
        }
        int rc = qt_make_round(d->fmt, &P);
        if (rc != QT_OK) return rc;
    }
    BlockParams &bp = j.bp;
    memset(&bp, 0, sizeof(bp));
    bp.quant_min = d->quant_min;
    bp.quant_max = d->quant_max;
    bp.range = d->quant_max - d->quant_min;
    bp.pow2 = (!affine && d->force_scale_power_of_two) ? 1 : 0;
    bp.qmax_exp = (int32_t)floor(log2((double)d->quant_max));
    bp.pow2_tab = static_cast<const uint32_t *>(d->pow2_table);
    bp.fast_ok = (!affine && P.tiny_safe) ? 1 : 0;
    {
        uint32_t qb;
        memcpy(&qb, &d->quant_max, 4);
        bp.qmax_short = ((qb & 0xFFFFu) == 0u && d->quant_max > 0x1p-60f && d->quant_max < 0x1p60f) ? 1 : 0;
        bp.rqmax = 1.0f / d->quant_max;
    }
    if (d->scale_fmt) {
        int rc = qt_make_round(d->scale_fmt, &bp.scale_round);
        if (rc != QT_OK) return rc;
        bp.has_scale_fmt = 1;
    }
    if (d->scale_table) {  // the codebook as a table (torch.ops.quantized_ops.calculate_mx_qparam's scale_qmap)
        bp.scale_table = static_cast<const uint16_t *>(d->scale_table);
        bp.has_scale_fmt = 1;
    }
    if (!(d->quant_max > 0.0f) && !affine) {
        qt_set_error("qt_fq_block: quant_max must be positive, got %g", (double)d->quant_max);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (total == 0) return QT_OK;
    const bool stat_only = d->y == nullptr;  // parameters only (calculate_mx_qparam)
    if (!d->x || !d->scale || (affine && !d->zero_point)) {
        qt_set_error("qt_fq_block: x, scale%s must not be NULL", affine ? ", zero_point" : "");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (bp.pow2 && !bp.pow2_tab) {
        qt_set_error("qt_fq_block: force_scale_power_of_two needs the device table of qt_block_pow2_table_host()");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const size_t esz = j.f32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->y)) % esz) {  // y == NULL passes
        qt_set_error("qt_fq_block: x / y not aligned to the element size");
        return QT_ERR_UNALIGNED;
    }
    if (num_sms() == 0) return no_device();

    int rc = QT_OK;
    if (affine) {
        rc = dispatch_direct_small(P, [&](auto tag, const auto &p) {
            using R = typename decltype(tag)::type;
            if (j.d->y && !j.bp.scale_table && (j.f32 ? try_gwa_fast<true>(j) : try_gwa_fast<false>(j))) return;
            j.f32 ? launch_generic<R, true, true>(j, p) : launch_generic<R, false, true>(j, p);
        });
    } else {
        // single-pass kernels (one translation unit per family); anything they cannot express takes the generic pair
        const bool single_pass_ok = d->y != nullptr && bp.scale_table == nullptr;
        if (!(single_pass_ok && (qtblk::try_flat(j, P) || qtblk::try_cols(j, P) || qtblk::try_tile(j, P))))
            rc = dispatch_rounder(P, d->lut, [&](auto tag, const auto &p) {
                using R = typename decltype(tag)::type;
                j.f32 ? launch_generic<R, true, false>(j, p) : launch_generic<R, false, false>(j, p);
            });
    }
    if (rc != QT_OK) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "block-scaled fake-quant kernel launch");
    return QT_OK;
}

extern "C" int qt_table_op(const qt_table_op_desc_t *d, void *stream)
{
    if (!d) {
        qt_set_error("qt_table_op: desc is NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if ((d->elem_type != QT_BF16 && d->elem_type != QT_F32) || (d->op != QT_TABLE_QUANTIZE && d->op != QT_TABLE_DEQUANTIZE && d->op != QT_TABLE_LOOKUP)) {
        qt_set_error("qt_table_op: bad elem_type %d or op %d", d->elem_type, d->op);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (d->d0 < 0 || d->n1 < 0 || d->d1 < 0 || d->n2 < 0 || d->d2 < 0 || (!d->scalar_params && d->block_size < 1)) {
        qt_set_error("qt_table_op: negative dimension or block_size < 1");
        return QT_ERR_INVALID_ARGUMENT;
    }
    BlockDims D;
    D.d0 = (size_t)d->d0; D.n1 = (size_t)d->n1; D.d1 = (size_t)d->d1; D.n2 = (size_t)d->n2; D.d2 = (size_t)d->d2;
    D.bs = d->scalar_params ? 1u : (uint32_t)d->block_size;
    D.bs2 = d->block_axis2 ? D.bs : 1u;
    D.nb1 = (D.n1 + D.bs - 1) / D.bs;
    D.nb2 = (D.n2 + D.bs2 - 1) / D.bs2;
    const size_t total = D.d0 * D.n1 * D.d1 * D.n2 * D.d2;
    if (total == 0) return QT_OK;
    if (!d->x || !d->y || (!d->scale && d->op != QT_TABLE_LOOKUP)) {
        qt_set_error("qt_table_op: x, y and scale must not be NULL");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (num_sms() == 0) return no_device();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = grid_for((total + 255) / 256, 8);
    const uint16_t *ta = static_cast<const uint16_t *>(d->table_a), *tb = static_cast<const uint16_t *>(d->table_b);
    const bool f32 = d->elem_type == QT_F32;
#define QT_TABLE_LAUNCH(F, O) \
    table_op_kernel<F, O><<<grid, 256, 0, st>>>(d->x, d->y, total, D, d->scale, d->zero_point, d->scalar_params, ta, tb)
    switch (d->op) {
    case QT_TABLE_QUANTIZE: f32 ? QT_TABLE_LAUNCH(true, QT_TABLE_QUANTIZE) : QT_TABLE_LAUNCH(false, QT_TABLE_QUANTIZE); break;
    case QT_TABLE_DEQUANTIZE: f32 ? QT_TABLE_LAUNCH(true, QT_TABLE_DEQUANTIZE) : QT_TABLE_LAUNCH(false, QT_TABLE_DEQUANTIZE); break;
    default: f32 ? QT_TABLE_LAUNCH(true, QT_TABLE_LOOKUP) : QT_TABLE_LAUNCH(false, QT_TABLE_LOOKUP); break;
    }
#undef QT_TABLE_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "table op kernel launch");
    return QT_OK;
}
