// qt_block_tile.cu -- microscaling, square blocks over the last two axes (see qt_block.cu for the overview).
#include "qt_block_common.cuh"

namespace {

// ----------------------------------------------------------------------------- tile kernel (two tiled axes)
// Tensor [outer, n1, n2], square blocks of BS x BS (BS = 8 * RPT) over the last two axes (ax = (-2, -1)),
// n2 % VEC == 0.  blockDim = (32, 8): a CTA holds BS rows x 32 column groups; a block is BS / VEC adjacent lanes wide
// (xor-shuffle) and 8 row phases x RPT register rows high (shared memory).  Rows / columns past the edge read as zero.
template <class R, bool F32, int RPT>
__global__ void __launch_bounds__(256)
mx_tile_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t outer, size_t n1, size_t n2_vec, size_t nb1,
               size_t nb2, const __grid_constant__ typename R::Params params, const __grid_constant__ BlockParams bp,
               float *__restrict__ scale_out)
{
    constexpr int VEC = F32 ? 4 : 8;
    constexpr int BS = 8 * RPT;
    constexpr int LANES = BS / VEC >= 1 ? BS / VEC : 1;
    static_assert(BS % VEC == 0 && LANES <= 32, "block width must be whole 16-byte vectors within a warp");
    const unsigned char *lut_smem = stage_table<R>(params);
    const R round(params, lut_smem);
    const typename FastOf<R>::type fast_round(params, lut_smem);
    const uint32_t *tab = stage_pow2_table(bp, R::kSmemBytes);
    __shared__ uint32_t red[8][33];
    const size_t cchunks = (n2_vec + 31) / 32;
    const size_t work = outer * nb1 * cchunks;
    for (size_t w = blockIdx.x; w < work; w += gridDim.x) {
        const size_t cc = w % cchunks, rest = w / cchunks;
        const size_t b1 = rest % nb1, o = rest / nb1;
        const size_t cv = cc * 32 + threadIdx.x;
        const bool active = cv < n2_vec;
        const size_t row0 = b1 * BS + threadIdx.y;
        const uint4 *xp = x + (o * n1) * n2_vec + cv;
        uint4 v[RPT];
        uint32_t a = 0u;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const size_t r = row0 + (size_t)k * 8;
            v[k] = (active && r < n1) ? ld_stream(xp + r * n2_vec) : make_uint4(0u, 0u, 0u, 0u);
            a = F32 ? amax_of_vec_f32(a, v[k]) : amax_of_vec_bf16(a, v[k]);
        }
#pragma unroll
        for (int off = 1; off < LANES; off <<= 1) a = max(a, __shfl_xor_sync(0xFFFFFFFFu, a, off));
        red[threadIdx.y][threadIdx.x] = a;
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 8; ++p) a = max(a, red[p][threadIdx.x]);
        __syncthreads();  // the next work item overwrites red[]
        const float s = mx_scale_fast<F32>(a, bp, tab);
        if (threadIdx.y == 0 && active && (threadIdx.x & (LANES - 1)) == 0)
            scale_out[(o * nb1 + b1) * nb2 + cv / LANES] = s;
        const float sa = a == 0u ? 1.0f : s;
        float rs = 0.0f;
        bool fast = false;
        if (!F32) {
            rs = __frcp_rn(sa);
            fast = __all_sync(0xFFFFFFFFu, mx_block_is_fast(a, sa, rs, bp));
        }
        uint4 *yp = y + (o * n1) * n2_vec + cv;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const size_t r = row0 + (size_t)k * 8;
            uint4 out;
            if (fast) {
                out.x = mx_word_fast(fast_round, v[k].x, sa, rs, sa, rs);
                out.y = mx_word_fast(fast_round, v[k].y, sa, rs, sa, rs);
                out.z = mx_word_fast(fast_round, v[k].z, sa, rs, sa, rs);
                out.w = mx_word_fast(fast_round, v[k].w, sa, rs, sa, rs);
            } else {
                out = mx_apply_vec<R, F32>(round, v[k], sa);
            }
            if (active && r < n1) st_stream(yp + r * n2_vec, out);
        }
    }
}

template <class R, bool F32, int RPT>
void launch_tile(const BlockJob &j, const typename R::Params &p, size_t n2_vec)
{
    allow_smem<mx_tile_kernel<R, F32, RPT>>(R::kSmemBytes + kPow2SmemBytes);
    const BlockDims &D = j.D;
    const size_t work = D.d0 * D.nb1 * ((n2_vec + 31) / 32);
    const unsigned grid = grid_for(work, R::kTable ? 3 : 8);
    mx_tile_kernel<R, F32, RPT><<<grid, dim3(32, 8), R::kSmemBytes + kPow2SmemBytes, j.stream>>>(
        static_cast<const uint4 *>(j.d->x), static_cast<uint4 *>(j.d->y), D.d0, D.n1, n2_vec, D.nb1, D.nb2, p, j.bp,
        j.d->scale);
}
template <class R, bool F32>
bool try_tile_t(const BlockJob &j, const typename R::Params &p)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const BlockDims &D = j.D;
    // two tiled axes, adjacent and last: [d0, n1, n2] (d1 = d2 = 1), n2 in whole vectors
    if (D.bs2 != D.bs || D.d1 != 1 || D.d2 != 1 || D.n2 % VEC != 0 || D.bs % VEC != 0) return false;
    if ((reinterpret_cast<uintptr_t>(j.d->x) | reinterpret_cast<uintptr_t>(j.d->y)) & 15u) return false;
    const size_t n2_vec = D.n2 / VEC;
    switch (D.bs) {
    case 8: launch_tile<R, F32, 1>(j, p, n2_vec); return true;
    case 16: launch_tile<R, F32, 2>(j, p, n2_vec); return true;
    case 32: launch_tile<R, F32, 4>(j, p, n2_vec); return true;
    case 64: launch_tile<R, F32, 8>(j, p, n2_vec); return true;
    case 128: launch_tile<R, F32, 16>(j, p, n2_vec); return true;
    default: return false;
    }
}

}  // namespace

bool qtblk::try_tile(const BlockJob &j, const QtRound &P)
{
    bool taken = false;
    dispatch_rounder(P, j.d->lut, [&](auto tag, const auto &p) {
        using R = typename decltype(tag)::type;
        taken = j.f32 ? try_tile_t<R, true>(j, p) : try_tile_t<R, false>(j, p);
    });
    return taken;
}
