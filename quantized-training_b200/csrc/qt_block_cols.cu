// qt_block_cols.cu -- microscaling, blocks along an inner axis (every column its own block) (see qt_block.cu for the overview).
#include "qt_block_common.cuh"

namespace {

// ----------------------------------------------------------------------------- cols kernel
// Tensor [outer, n, inner], blocks of BS = 8 * RPT rows along n, inner % VEC == 0: every column is a block of its
// own.  blockDim = (32, 8): threadIdx.x -> a 16-byte column group g of the flattened (outer, inner / VEC) space,
// threadIdx.y -> row phase; a thread keeps its RPT rows in registers.  Rows past n read as zero (the reference pads).
template <class R, bool F32, int RPT>
__global__ void __launch_bounds__(256)
mx_cols_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t outer, size_t n, size_t inner_vec,
               size_t nblk, const __grid_constant__ typename R::Params params,
               const __grid_constant__ BlockParams bp, float *__restrict__ scale_out)
{
    constexpr int VEC = F32 ? 4 : 8;
    constexpr int BS = 8 * RPT;
    const unsigned char *lut_smem = stage_table<R>(params);
    const R round(params, lut_smem);
    const typename FastOf<R>::type fast_round(params, lut_smem);
    const uint32_t *tab = stage_pow2_table(bp, R::kSmemBytes);
    __shared__ uint4 red[8][33];           // per row phase: packed maxima of a column group
    __shared__ float col_scale[32][VEC + 1];
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
        // per-column maxima of |x| bit patterns: fp32 four 32-bit maxima, bf16 eight packed 16-bit maxima
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            if (F32) {
                m.x = max(m.x, v[k].x & 0x7FFFFFFFu);
                m.y = max(m.y, v[k].y & 0x7FFFFFFFu);
                m.z = max(m.z, v[k].z & 0x7FFFFFFFu);
                m.w = max(m.w, v[k].w & 0x7FFFFFFFu);
            } else {
                m.x = __vmaxu2(m.x, v[k].x & 0x7FFF7FFFu);
                m.y = __vmaxu2(m.y, v[k].y & 0x7FFF7FFFu);
                m.z = __vmaxu2(m.z, v[k].z & 0x7FFF7FFFu);
                m.w = __vmaxu2(m.w, v[k].w & 0x7FFF7FFFu);
            }
        }
        red[threadIdx.y][threadIdx.x] = m;
        __syncthreads();
        if (threadIdx.y < VEC) {  // row phase c computes the scale of column c of the group
            const int c = threadIdx.y;
            uint32_t a = 0u;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const uint4 q = red[p][threadIdx.x];
                const uint32_t wsel = F32 ? (c == 0 ? q.x : c == 1 ? q.y : c == 2 ? q.z : q.w)
                                          : ((c >> 1) == 0 ? q.x : (c >> 1) == 1 ? q.y : (c >> 1) == 2 ? q.z : q.w);
                a = max(a, F32 ? wsel : ((c & 1) ? (wsel & 0xFFFF0000u) : (wsel << 16)));
            }
            const float s = mx_scale_fast<F32>(a, bp, tab);
            if (active) scale_out[(o * nblk + b) * (inner_vec * VEC) + cv * VEC + c] = s;
            // the scale the column is applied with (1 for an all-zero block, see mx_flat_tile); scales are
            // positive, so the sign bit is free to carry "this column needs the careful path"
            const float sa = a == 0u ? 1.0f : s;
            const bool f = !F32 && mx_block_is_fast(a, sa, __frcp_rn(sa), bp);
            col_scale[threadIdx.x][c] = f ? sa : -sa;
        }
        __syncthreads();
        ScaleBf16 sc[VEC];
        bool recip_ok[VEC];
        bool mine_fast = !F32;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const float sv = col_scale[threadIdx.x][c];
            mine_fast = mine_fast && !(__float_as_uint(sv) >> 31);
            sc[c] = make_scale(fabsf(sv));
            recip_ok[c] = classify_scale(sc[c].s) != DIV_EXACT;
        }
        const bool fast = __all_sync(0xFFFFFFFFu, mine_fast);
        uint4 *yp = y + (o * n) * inner_vec + cv;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const size_t r = row0 + (size_t)k * 8;
            const uint32_t win[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!F32 && fast) {
                    out[q] = mx_word_fast(fast_round, win[q], sc[(2 * q) % VEC].s, sc[(2 * q) % VEC].rs,
                                          sc[(2 * q + 1) % VEC].s, sc[(2 * q + 1) % VEC].rs);
                } else if (F32) {
                    out[q] = fq_f32<R, false>(round, win[q], sc[q].s);
                } else {
                    const uint32_t lo = win[q] << 16, hi = win[q] & 0xFFFF0000u;
                    const float qlo = recip_ok[2 * q] ? bf16_quotient<DIV_RECIP>(lo, sc[2 * q])
                                                      : bf16_quotient<DIV_EXACT>(lo, sc[2 * q]);
                    const float qhi = recip_ok[2 * q + 1] ? bf16_quotient<DIV_RECIP>(hi, sc[2 * q + 1])
                                                          : bf16_quotient<DIV_EXACT>(hi, sc[2 * q + 1]);
                    const uint32_t uq = bf16x2_rne(qlo, qhi);
                    out[q] = bf16x2_rne(__fmul_rn(__uint_as_float(round.lo(uq)), sc[2 * q].s),
                                        __fmul_rn(__uint_as_float(round.hi(uq)), sc[2 * q + 1].s));
                }
            }
            if (active && r < n) st_stream(yp + r * inner_vec, make_uint4(out[0], out[1], out[2], out[3]));
        }
    }
}

template <class R, bool F32, int RPT>
void launch_cols(const BlockJob &j, const typename R::Params &p, size_t outer, size_t n, size_t inner_vec)
{
    allow_smem<mx_cols_kernel<R, F32, RPT>>(R::kSmemBytes + kPow2SmemBytes);
    const size_t nblk = (n + 8 * RPT - 1) / (8 * RPT);
    const size_t work = nblk * ((outer * inner_vec + 31) / 32);
    const unsigned grid = grid_for(work, R::kTable ? 3 : 8);
    mx_cols_kernel<R, F32, RPT><<<grid, dim3(32, 8), R::kSmemBytes + kPow2SmemBytes, j.stream>>>(
        static_cast<const uint4 *>(j.d->x), static_cast<uint4 *>(j.d->y), outer, n, inner_vec, nblk, p, j.bp,
        j.d->scale);
}
template <class R, bool F32>
bool try_cols_t(const BlockJob &j, const typename R::Params &p)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const BlockDims &D = j.D;
    // one tiled axis with a dense inner part: [d0, n1, d1] (n2 = d2 = 1), d1 % VEC == 0
    if (D.n2 != 1 || D.d2 != 1 || D.bs2 != 1 || D.d1 < VEC || D.d1 % VEC != 0) return false;
    if ((reinterpret_cast<uintptr_t>(j.d->x) | reinterpret_cast<uintptr_t>(j.d->y)) & 15u) return false;
    const size_t inner_vec = D.d1 / VEC;
    switch (D.bs) {
    case 8: launch_cols<R, F32, 1>(j, p, D.d0, D.n1, inner_vec); return true;
    case 16: launch_cols<R, F32, 2>(j, p, D.d0, D.n1, inner_vec); return true;
    case 32: launch_cols<R, F32, 4>(j, p, D.d0, D.n1, inner_vec); return true;
    case 64: launch_cols<R, F32, 8>(j, p, D.d0, D.n1, inner_vec); return true;
    case 128: launch_cols<R, F32, 16>(j, p, D.d0, D.n1, inner_vec); return true;
    default: return false;
    }
}

}  // namespace

bool qtblk::try_cols(const BlockJob &j, const QtRound &P)
{
    bool taken = false;
    dispatch_rounder(P, j.d->lut, [&](auto tag, const auto &p) {
        using R = typename decltype(tag)::type;
        taken = j.f32 ? try_cols_t<R, true>(j, p) : try_cols_t<R, false>(j, p);
    });
    return taken;
}
