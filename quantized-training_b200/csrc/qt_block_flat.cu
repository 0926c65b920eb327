// qt_block_flat.cu -- microscaling, blocks along the unit-stride axis (see qt_block.cu for the overview).
#include "qt_block_common.cuh"

namespace {

// ----------------------------------------------------------------------------- flat kernel
// Blocks of LANES consecutive 16-byte vectors (last axis, shape[-1] % block_size == 0): block b = vector i / LANES,
// which is also its index in the row-major block grid.  nvec % LANES == 0.
template <class R, bool F32, int LANES, bool CHECK>
__device__ __forceinline__ void mx_flat_tile(const R &round, const typename FastOf<R>::type &fast_round,
                                             const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t nvec,
                                             size_t base, const BlockParams &bp, const uint32_t *tab,
                                             float *__restrict__ scale_out)
{
    const size_t nthr = blockDim.x;
    uint4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
        const size_t i = base + (size_t)j * nthr;
        v[j] = (!CHECK || i < nvec) ? ld_stream(x + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
        const size_t i = base + (size_t)j * nthr;
        uint32_t a = F32 ? amax_of_vec_f32(0u, v[j]) : amax_of_vec_bf16(0u, v[j]);
#pragma unroll
        for (int o = 1; o < LANES; o <<= 1) a = max(a, __shfl_xor_sync(0xFFFFFFFFu, a, o));
        const float s = mx_scale_fast<F32>(a, bp, tab);
        // an all-zero block gives +-0 whatever its scale is: apply it with 1 (its true scale may be below 2^-126,
        // whose reciprocal overflows)
        const float sa = a == 0u ? 1.0f : s;
        uint4 r;
        bool fast = false;
        float rs = 0.0f;
        if (!F32) {
            rs = __frcp_rn(sa);
            fast = __all_sync(0xFFFFFFFFu, mx_block_is_fast(a, sa, rs, bp));
        }
        if (fast) {
            r.x = mx_word_fast(fast_round, v[j].x, sa, rs, sa, rs);
            r.y = mx_word_fast(fast_round, v[j].y, sa, rs, sa, rs);
            r.z = mx_word_fast(fast_round, v[j].z, sa, rs, sa, rs);
            r.w = mx_word_fast(fast_round, v[j].w, sa, rs, sa, rs);
        } else {
            r = mx_apply_vec<R, F32>(round, v[j], sa);
        }
        if (!CHECK || i < nvec) {
            if ((i & (size_t)(LANES - 1)) == 0) scale_out[i / LANES] = s;
            st_stream(y + i, r);
        }
    }
}

template <class R, bool F32, int LANES>
__global__ void __launch_bounds__(R::kThreads, R::kMinCtas)
mx_flat_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t nvec,
               const __grid_constant__ typename R::Params params, const __grid_constant__ BlockParams bp,
               float *__restrict__ scale_out)
{
    const unsigned char *lut_smem = stage_table<R>(params);
    const R round(params, lut_smem);
    const typename FastOf<R>::type fast_round(params, lut_smem);
    const uint32_t *tab = stage_pow2_table(bp, R::kSmemBytes);
    const size_t tile = (size_t)blockDim.x * kUnroll;
    const size_t full_tiles = nvec / tile;
    for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x)
        mx_flat_tile<R, F32, LANES, false>(round, fast_round, x, y, nvec, t * tile + threadIdx.x, bp, tab, scale_out);
    if (full_tiles * tile < nvec && blockIdx.x == full_tiles % gridDim.x)
        mx_flat_tile<R, F32, LANES, true>(round, fast_round, x, y, nvec, full_tiles * tile + threadIdx.x, bp, tab,
                                          scale_out);
}

template <class R, bool F32, int LANES>
void launch_flat(const BlockJob &j, const typename R::Params &p, size_t nvec)
{
    allow_smem<mx_flat_kernel<R, F32, LANES>>(R::kSmemBytes + kPow2SmemBytes);
    const size_t tile = (size_t)R::kThreads * kUnroll;
    const unsigned grid = grid_for((nvec + tile - 1) / tile, R::kCtasPerSm);
    mx_flat_kernel<R, F32, LANES><<<grid, R::kThreads, R::kSmemBytes + kPow2SmemBytes, j.stream>>>(
        static_cast<const uint4 *>(j.d->x), static_cast<uint4 *>(j.d->y), nvec, p, j.bp, j.d->scale);
}
template <class R, bool F32>
bool try_flat_t(const BlockJob &j, const typename R::Params &p)
{
    constexpr size_t VEC = F32 ? 4 : 8;
    const BlockDims &D = j.D;
    // blocks along the unit-stride axis only: [d0, n1] with n1 % bs == 0 (d1 = n2 = d2 = 1)
    if (D.d1 != 1 || D.n2 != 1 || D.d2 != 1 || D.bs2 != 1) return false;
    if (D.n1 % D.bs != 0 || D.bs % VEC != 0) return false;
    if ((reinterpret_cast<uintptr_t>(j.d->x) | reinterpret_cast<uintptr_t>(j.d->y)) & 15u) return false;
    const size_t lanes = D.bs / VEC;
    const size_t nvec = D.d0 * D.n1 / VEC;
    switch (lanes) {
    case 1: launch_flat<R, F32, 1>(j, p, nvec); return true;
    case 2: launch_flat<R, F32, 2>(j, p, nvec); return true;
    case 4: launch_flat<R, F32, 4>(j, p, nvec); return true;
    case 8: launch_flat<R, F32, 8>(j, p, nvec); return true;
    case 16: launch_flat<R, F32, 16>(j, p, nvec); return true;
    case 32: launch_flat<R, F32, 32>(j, p, nvec); return true;
    default: return false;
    }
}

}  // namespace

bool qtblk::try_flat(const BlockJob &j, const QtRound &P)
{
    bool taken = false;
    dispatch_rounder(P, j.d->lut, [&](auto tag, const auto &p) {
        using R = typename decltype(tag)::type;
        taken = j.f32 ? try_flat_t<R, true>(j, p) : try_flat_t<R, false>(j, p);
    });
    return taken;
}
