// qt_attn.cu -- quantized attention core in one kernel: scores never leave the SM.
//
//   ctx = fq_out( fq_p( softmax( fq_mid( fq_pre(q k^T) * alpha + mask ) ) ) v )          per (batch, head)
//
// Reference chain (modules/quantizable/modeling_bert.py:118-162, modeling_llama.py:228-263): qk_matmul -> [fq]
// -> attn_scaling -> + mask -> [fq] -> nn.Softmax -> [fq] -> av_matmul -> permute -> [fq of the output projection's
// input], each a bf16 ATen op; executed separately that is a 67 MB score tensor written and read three times per
// Llama-2-7B layer.  Here one CTA owns 128 query rows of one head and walks the key blocks TWICE:
//   pass 1  S_j = Q K_j^T (tcgen05, fp32 in TMEM) -> the op chain up to the softmax input, rounded to bf16 at every
//           point the reference materialises a tensor -> running row maximum and sum (exact two-pass softmax
//           statistics; the fake quant of the probabilities needs the FINAL row sum, so a one-pass "online" scheme
//           cannot produce the reference's values)
//   pass 2  S_j again -> p = bf16(exp(s - m) / l) -> fq -> P_j written to shared memory as the K-major A operand
//           (bf16, or fp8 codes) -> O += P_j V_j (tcgen05) -> epilogue: bf16 -> [fq] -> TMA store into [B, S, H*D]
// Recomputing S costs 1/2 of the attention FLOPs again (0.3 % of a Llama-2-7B forward) and removes every byte of
// score / probability traffic.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocation, warps 2-9 softmax / epilogue
// (thread = query row: the 32x32b TMEM load gives every thread its own row, so row statistics need no shuffles; the two
// warps of a lane quarter split each 128-key block in halves and merge their statistics once, after pass 1).
// TMEM: S double-buffered (2 x 128 columns), O (head_dim columns).  Shared memory: Q, 2 x K_j, 2 x V_j^T, P, the 8 KB
// rounding table, 8 x 4 KB output staging.
#include <cuda_fp8.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "qt_fq_common.cuh"
#include "qt_tc.cuh"

namespace {

constexpr int AT_M = 128;      // query rows per CTA
constexpr int AT_KEYS = 128;   // keys per block
constexpr int AT_THREADS = 320;
constexpr int AT_SOFTMAX_WARPS = 8;
constexpr int FQ_PRE = 1, FQ_MID = 2, FQ_POST = 4, FQ_OUT = 8;

struct AttnParams {
    int B, H, Sq, Sk, D;
    int q_blocks, key_blocks;
    int kq_bytes;  // bytes of one Q / K row = D * esz (128 or 256)
    float alpha;
    int has_alpha;
    const __nv_bfloat16 *mask;  // [mask_batches, mask_rows, Sk] additive, or null
    int mask_rows, mask_batches;
    int causal;  // the mask is the standard causal one: key blocks above the diagonal are skipped
    int flags;   // FQ_*
    int out_codes;  // 0 bf16 | 1 e4m3 | 2 e5m2 codes for the context
    int p_codes;    // probabilities (and V) as fp8 codes: P V on the FP8 tensor cores
    uint32_t idesc_s, idesc_o;
    TableParams table;  // rounding table of the (single) activation format; kind below
    int fq_kind;        // 0 none, 1 table, 2 integer direct
    QtRound rp;
};

struct Rounder {  // run-time rounding engine over the staged table (fp / posit) or the integer logic
    const unsigned char *tab;
    uint32_t clamp_bits, mx_band;
    int kind;
    const QtRound *rp;
    __device__ __forceinline__ float operator()(float f) const
    {
        const uint32_t u = __float_as_uint(f);
        if (kind == 1) {
            const uint32_t a = u & 0x7FFFFFFFu;
            uint32_t q = qt_lut_round_smem<false, 1>(tab, 0u, u >> 16, a, min(a, clamp_bits));
            if (mx_band && a >= 0x7F580000u && a != 0x7F800000u) q = QT_NAN_BITS;
            return __uint_as_float(q);
        }
        if (kind == 2) return __uint_as_float(qt_round<QTR_INT>(*rp, u));
        return f;
    }
};

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void round_pair(float &a, float &b)
{
    const uint32_t p = bf16x2_rne(a, b);
    a = __uint_as_float(p << 16);
    b = __uint_as_float(p & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t fp8x2_codes(float lo, float hi, int kind)
{
    uint32_t c = kind == 2 ? (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E5M2)
                           : (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3);
    const uint32_t inf_code = kind == 2 ? 0x7Cu : 0x7Fu;
    const uint32_t bl = __float_as_uint(lo), bh = __float_as_uint(hi);
    if ((bl & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0xFF00u) | ((bl >> 24) & 0x80u) | inf_code;
    if ((bh & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0x00FFu) | ((((bh >> 24) & 0x80u) | inf_code) << 8);
    return c;
}
__device__ __forceinline__ void named_barrier_softmax()
{
    asm volatile("bar.sync 1, %0;" ::"r"(AT_SOFTMAX_WARPS * 32) : "memory");
}

// 64 accumulator columns of one score block -> the softmax INPUT values of the reference chain (bf16 grid), in place
__device__ __forceinline__ void softmax_input(float (&f)[64], const AttnParams &p, const Rounder &round,
                                              const __nv_bfloat16 *mrow, int key0, bool row_live)
{
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        float *x = f + g * 8;
#pragma unroll
        for (int j = 0; j < 8; j += 2) round_pair(x[j], x[j + 1]);  // the bf16 output of qk_matmul
        if (p.flags & FQ_PRE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = round(x[j]);
        }
        if (p.has_alpha) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] *= p.alpha;
#pragma unroll
            for (int j = 0; j < 8; j += 2) round_pair(x[j], x[j + 1]);
        }
        const int k = key0 + g * 8;
        if (mrow && row_live && k < p.Sk) {  // Sk % 8 == 0: groups are entirely in or out
            const uint4 mm = __ldg(reinterpret_cast<const uint4 *>(mrow + k));
            const uint32_t w[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x[2 * j] += __uint_as_float(w[j] << 16);
                x[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) round_pair(x[j], x[j + 1]);
        }
        if (p.flags & FQ_MID) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = round(x[j]);
        }
        if (k >= p.Sk) {  // keys past the end (zero-filled by TMA) do not exist
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = -INFINITY;
        }
    }
}

template <bool FP8>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_vt, const __grid_constant__ CUtensorMap map_o,
               const __grid_constant__ AttnParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // layout (bytes): Q | K[2] | VT[2] | P | table 8 KB | barriers, statistics.  The 8 x 4 KB output staging of the
    // epilogue reuses the K ring (every MMA has completed by then).
    const uint32_t q_bytes = AT_M * p.kq_bytes;                     // 16 / 32 KB
    const uint32_t k_bytes = AT_KEYS * p.kq_bytes;                  // per stage
    const uint32_t esz = FP8 ? 1u : 2u;
    const uint32_t v_bytes = (uint32_t)p.D * AT_KEYS * (p.p_codes ? 1u : 2u);  // per stage: D rows x 128 keys
    const uint32_t p_bytes = AT_M * AT_KEYS * (p.p_codes ? 1u : 2u);
    const uint32_t sQ = base, sK = sQ + q_bytes, sV = sK + 2 * k_bytes, sP = sV + 2 * v_bytes;
    const uint32_t sStage = sK, sTab = sP + p_bytes, bars = sTab + QT_LUT_BYTES;
    (void)esz;
    auto bar = [&](int i) { return bars + 8u * i; };
    enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 3, B_VFULL = 5, B_VEMPTY = 7, B_SFULL = 9, B_SEMPTY = 11, B_PFULL = 13,
           B_PEMPTY = 14, B_OFULL = 15, B_COUNT = 16 };
    const uint32_t tmem_slot = bar(B_COUNT);
    const uint32_t stats = tmem_slot + 16;  // float2[128] partial row statistics of the second column half (1 KB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA: query block of one head; longest (most key blocks under a causal mask) first
    const int qb = p.q_blocks - 1 - (int)(blockIdx.x % p.q_blocks);
    const int bh = blockIdx.x / p.q_blocks, h = bh % p.H, b = bh / p.H;
    const int q0 = qb * AT_M;
    int nblocks = p.key_blocks;
    if (p.causal) nblocks = min(nblocks, (q0 + AT_M + AT_KEYS - 1) / AT_KEYS);
    const int kq_kblocks = p.kq_bytes / ROW_BYTES;                       // 128-byte k-blocks of a Q / K row (1 or 2)
    const int pv_kblocks = (AT_KEYS * (p.p_codes ? 1 : 2)) / ROW_BYTES;  // of a P / V^T row (1 or 2)

    if (warp == 0 && lane == 0) {
        mbar_init(bar(B_QFULL), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(B_KFULL + s), 1);
            mbar_init(bar(B_KEMPTY + s), 1);
            mbar_init(bar(B_VFULL + s), 1);
            mbar_init(bar(B_VEMPTY + s), 1);
            mbar_init(bar(B_SFULL + s), 1);
            mbar_init(bar(B_SEMPTY + s), AT_SOFTMAX_WARPS);
        }
        mbar_init(bar(B_PFULL), AT_SOFTMAX_WARPS);
        mbar_init(bar(B_PEMPTY), 1);
        mbar_init(bar(B_OFULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_vt)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && p.fq_kind == 1) {  // stage the 8 KB rounding table (softmax warps only use it)
        const float4 *src = reinterpret_cast<const float4 *>(p.table.table);
        for (int i = threadIdx.x - 64; i < QT_LUT_BYTES / 16; i += AT_SOFTMAX_WARPS * 32) {
            const float4 e = src[i];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sTab + 16u * i), "f"(e.x), "f"(e.y), "f"(e.z),
                         "f"(e.w)
                         : "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tmem_o = tmem_base + 256;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_arrive_expect_tx(bar(B_QFULL), q_bytes);
            for (int kb = 0; kb < kq_kblocks; ++kb)
                tma_load_4d(sQ + kb * (AT_M * ROW_BYTES), &map_q, bar(B_QFULL), kb * (ROW_BYTES / (FP8 ? 1 : 2)), q0, h, b);
            uint32_t kc = 0, vc = 0;  // running use counts of the K and V rings (2 stages each)
            for (int pass = 0; pass < 2; ++pass) {
                for (int j = 0; j < nblocks; ++j) {
                    const int ks = kc & 1;
                    mbar_wait(bar(B_KEMPTY + ks), ((kc >> 1) & 1) ^ 1u);
                    mbar_arrive_expect_tx(bar(B_KFULL + ks), k_bytes);
                    for (int kb = 0; kb < kq_kblocks; ++kb)
                        tma_load_4d(sK + ks * k_bytes + kb * (AT_KEYS * ROW_BYTES), &map_k, bar(B_KFULL + ks),
                                    kb * (ROW_BYTES / (FP8 ? 1 : 2)), j * AT_KEYS, h, b);
                    ++kc;
                    if (pass == 1) {
                        const int vs = vc & 1;
                        mbar_wait(bar(B_VEMPTY + vs), ((vc >> 1) & 1) ^ 1u);
                        mbar_arrive_expect_tx(bar(B_VFULL + vs), v_bytes);
                        for (int kb = 0; kb < pv_kblocks; ++kb)
                            tma_load_4d(sV + vs * v_bytes + kb * (p.D * ROW_BYTES), &map_vt, bar(B_VFULL + vs),
                                        j * AT_KEYS + kb * (ROW_BYTES / (p.p_codes ? 1 : 2)), 0, h, b);
                        ++vc;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t kc = 0, vc = 0, sc = 0, pc = 0;
            auto issue_s = [&](void) {  // S[sc & 1] = Q K_j^T for the next key block in ring order
                const int ks = kc & 1, sb = sc & 1;
                mbar_wait(bar(B_KFULL + ks), (kc >> 1) & 1);
                mbar_wait(bar(B_SEMPTY + sb), ((sc >> 1) & 1) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + sb * AT_KEYS;
                int n = 0;
                for (int kb = 0; kb < kq_kblocks; ++kb) {
                    const uint64_t da = make_smem_desc(sQ + kb * (AT_M * ROW_BYTES));
                    const uint64_t db = make_smem_desc(sK + ks * k_bytes + kb * (AT_KEYS * ROW_BYTES));
#pragma unroll
                    for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k, ++n)
                        tcgen05_mma<FP8>(d, da + ((k * MMA_K_BYTES) >> 4), db + ((k * MMA_K_BYTES) >> 4), p.idesc_s, n != 0);
                }
                tcgen05_commit(bar(B_KEMPTY + ks));
                tcgen05_commit(bar(B_SFULL + sb));
                ++kc;
                ++sc;
            };
            mbar_wait(bar(B_QFULL), 0);
            tcgen05_fence_after();
            for (int j = 0; j < nblocks; ++j) issue_s();  // pass 1: scores only
            // pass 2: S_{j+1} is issued before P_j is awaited, so the softmax of block j+1 overlaps P_j V_j
            issue_s();
            for (int j = 0; j < nblocks; ++j) {
                if (j + 1 < nblocks) issue_s();
                const int vs = vc & 1;
                mbar_wait(bar(B_VFULL + vs), (vc >> 1) & 1);
                mbar_wait(bar(B_PFULL), pc & 1);
                tcgen05_fence_after();
                int n = 0;
                for (int kb = 0; kb < pv_kblocks; ++kb) {
                    const uint64_t da = make_smem_desc(sP + kb * (AT_M * ROW_BYTES));
                    const uint64_t db = make_smem_desc(sV + vs * v_bytes + kb * (p.D * ROW_BYTES));
#pragma unroll
                    for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k, ++n) {
                        const uint64_t koff = (uint64_t)((k * MMA_K_BYTES) >> 4);
                        if (p.p_codes)
                            tcgen05_mma<true>(tmem_o, da + koff, db + koff, p.idesc_o, (j | n) != 0);
                        else
                            tcgen05_mma<false>(tmem_o, da + koff, db + koff, p.idesc_o, (j | n) != 0);
                    }
                }
                tcgen05_commit(bar(B_VEMPTY + vs));
                tcgen05_commit(bar(B_PEMPTY));
                ++vc;
                ++pc;
            }
            tcgen05_commit(bar(B_OFULL));
        }
    } else {
        // ===== softmax / epilogue warps 2..9 =====
        const int e = warp - 2;
        const int quarter = warp & 3, half = e >> 2;  // TMEM lanes of this warp; which 64 keys of a block
        const int r = quarter * 32 + lane;            // row of the tile owned by this thread
        const int q = q0 + r;
        const bool row_live = q < p.Sq;
        Rounder round;
        round.tab = reinterpret_cast<const unsigned char *>(__cvta_shared_to_generic((size_t)sTab));
        round.clamp_bits = p.table.cfg.clamp_bits;
        round.mx_band = p.table.cfg.mx_band;
        round.kind = p.fq_kind;
        round.rp = &p.rp;
        const __nv_bfloat16 *mrow = nullptr;
        if (p.mask)
            mrow = p.mask + ((size_t)(p.mask_batches > 1 ? b : 0) * p.mask_rows + (p.mask_rows > 1 ? min(q, p.Sq - 1) : 0)) *
                                (size_t)p.Sk;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t sc = 0;
        const float LOG2E = 1.4426950408889634f;

        // ---- pass 1: row maximum and sum over this warp's column halves
        float m = -INFINITY, l = 0.0f;
        for (int j = 0; j < nblocks; ++j, ++sc) {
            const int sb = sc & 1;
            mbar_wait(bar(B_SFULL + sb), (sc >> 1) & 1);
            tcgen05_fence_after();
            uint32_t v[64];
            tmem_ld_32x32_nowait(lane_addr + sb * AT_KEYS + half * 64, v);
            tmem_ld_32x32_nowait(lane_addr + sb * AT_KEYS + half * 64 + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_SEMPTY + sb));
            float f[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) f[i] = __uint_as_float(v[i]);
            softmax_input(f, p, round, mrow, j * AT_KEYS + half * 64, row_live);
            float cm = f[0];
#pragma unroll
            for (int i = 1; i < 64; ++i) cm = fmaxf(cm, f[i]);
            const float mn = fmaxf(m, cm);
            if (mn != -INFINITY) {
                float s = 0.0f;
#pragma unroll
                for (int i = 0; i < 64; ++i) s += ex2_approx((f[i] - mn) * LOG2E);
                l = l * ex2_approx((m - mn) * LOG2E) + s;
                m = mn;
            }
        }
        // ---- merge the two column halves of every row
        if (half == 1) {
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stats + 8u * r), "f"(m), "f"(l) : "memory");
        }
        named_barrier_softmax();
        if (half == 0) {
            float m1, l1;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m1), "=f"(l1) : "r"(stats + 8u * r));
            const float mn = fmaxf(m, m1);
            if (mn != -INFINITY) l = l * ex2_approx((m - mn) * LOG2E) + l1 * ex2_approx((m1 - mn) * LOG2E);
            m = mn;
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stats + 8u * r), "f"(m), "f"(l) : "memory");
        }
        named_barrier_softmax();
        if (half == 1) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m), "=f"(l) : "r"(stats + 8u * r));
        const float inv = __frcp_rn(l);  // all-masked rows: l = 0, m = -inf -> NaN probabilities, as torch.softmax gives

        // ---- pass 2: probabilities -> fq -> P (A operand of P V) in shared memory
        uint32_t pc = 0;
        for (int j = 0; j < nblocks; ++j, ++sc, ++pc) {
            const int sb = sc & 1;
            mbar_wait(bar(B_SFULL + sb), (sc >> 1) & 1);
            tcgen05_fence_after();
            uint32_t v[64];
            tmem_ld_32x32_nowait(lane_addr + sb * AT_KEYS + half * 64, v);
            tmem_ld_32x32_nowait(lane_addr + sb * AT_KEYS + half * 64 + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_SEMPTY + sb));
            float f[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) f[i] = __uint_as_float(v[i]);
            softmax_input(f, p, round, mrow, j * AT_KEYS + half * 64, row_live);
            uint32_t packed[32];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float *x = f + g * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ex2_approx((x[i] - m) * LOG2E) * inv;
#pragma unroll
                for (int i = 0; i < 8; i += 2) round_pair(x[i], x[i + 1]);
                if (p.flags & FQ_POST) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = round(x[i]);
                }
                if (p.p_codes) {
                    packed[g * 2] = fp8x2_codes(x[0], x[1], p.p_codes) | (fp8x2_codes(x[2], x[3], p.p_codes) << 16);
                    packed[g * 2 + 1] = fp8x2_codes(x[4], x[5], p.p_codes) | (fp8x2_codes(x[6], x[7], p.p_codes) << 16);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        packed[g * 4 + i] = __byte_perm(__float_as_uint(x[2 * i]), __float_as_uint(x[2 * i + 1]), 0x7632);
                }
            }
            mbar_wait(bar(B_PEMPTY), (pc & 1) ^ 1u);  // P_{j-1} V_{j-1} has consumed the buffer
            // K-major, 128-byte swizzle: 16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4)
            if (p.p_codes) {  // one k-block of 128 keys = 128 bytes per row; this warp's 64 keys are chunks 4 half..
                const uint32_t rowp = sP + (uint32_t)r * 128u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t dst = rowp + (uint32_t)(((half * 4 + c) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[c * 4]),
                                 "r"(packed[c * 4 + 1]), "r"(packed[c * 4 + 2]), "r"(packed[c * 4 + 3])
                                 : "memory");
                }
            } else {  // two k-blocks of 64 keys; this warp's 64 keys are k-block `half`
                const uint32_t rowp = sP + (uint32_t)half * (AT_M * ROW_BYTES) + (uint32_t)r * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t dst = rowp + (uint32_t)((c ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[c * 4]),
                                 "r"(packed[c * 4 + 1]), "r"(packed[c * 4 + 2]), "r"(packed[c * 4 + 3])
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> the MMA's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_PFULL));
        }

        // ---- epilogue: O -> bf16 -> [fq] -> [B, S, H*D]; this warp stores columns [64 half, 64 half + 64)
        mbar_wait(bar(B_OFULL), 0);
        tcgen05_fence_after();
        if (half * 64 < p.D) {
            uint32_t v[64];
            tmem_ld_32x32_nowait(lane_addr + 256 + half * 64, v);
            tmem_ld_32x32_nowait(lane_addr + 256 + half * 64 + 32, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t packed[32];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g * 8 + i]);
#pragma unroll
                for (int i = 0; i < 8; i += 2) round_pair(x[i], x[i + 1]);  // the bf16 output of av_matmul
                if (p.flags & FQ_OUT) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = round(x[i]);
                }
                if (p.out_codes) {
                    packed[g * 2] = fp8x2_codes(x[0], x[1], p.out_codes) | (fp8x2_codes(x[2], x[3], p.out_codes) << 16);
                    packed[g * 2 + 1] = fp8x2_codes(x[4], x[5], p.out_codes) | (fp8x2_codes(x[6], x[7], p.out_codes) << 16);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        packed[g * 4 + i] = __byte_perm(__float_as_uint(x[2 * i]), __float_as_uint(x[2 * i + 1]), 0x7632);
                }
            }
            const uint32_t buf = sStage + (uint32_t)e * 4096u;
            if (p.out_codes) {
                const uint32_t rowbuf = buf + (uint32_t)lane * 64u;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t dst = rowbuf + (uint32_t)((g ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[g * 4]),
                                 "r"(packed[g * 4 + 1]), "r"(packed[g * 4 + 2]), "r"(packed[g * 4 + 3])
                                 : "memory");
                }
            } else {
                const uint32_t rowbuf = buf + (uint32_t)lane * 128u;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint32_t dst = rowbuf + (uint32_t)((g ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[g * 4]),
                                 "r"(packed[g * 4 + 1]), "r"(packed[g * 4 + 2]), "r"(packed[g * 4 + 3])
                                 : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && q0 + quarter * 32 < p.Sq) {
                asm volatile(
                    "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                        reinterpret_cast<uint64_t>(&map_o)),
                    "r"(buf), "r"(half * 64), "r"(q0 + quarter * 32), "r"(h), "r"(b)
                    : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

uint32_t attn_idesc(int a_fmt, int b_fmt, int n)
{
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(AT_M >> 4) << 24);
}

}  // namespace

extern "C" int qt_attention_fq(const qt_attn_desc_t *d, void *stream)
{
    if (!d || !d->q || !d->k || !d->vt || !d->out) {
        qt_set_error("qt_attention_fq: NULL descriptor or pointer");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const bool fp8 = d->qk_type != QT_GEMM_BF16;
    const bool p_fp8 = d->pv_type != QT_GEMM_BF16;
    const int esz = fp8 ? 1 : 2;
    const int D = (int)d->head_dim;
    if (d->batch < 1 || d->heads < 1 || d->seq_q < 1 || d->seq_k < 1 || (D != 64 && D != 128) || (D * esz) % 128 ||
        d->seq_k % 16 || d->qk_type < QT_GEMM_BF16 || d->qk_type > QT_GEMM_E5M2_E4M3 || d->pv_type < QT_GEMM_BF16 ||
        d->pv_type > QT_GEMM_E5M2_E4M3 || d->out_type < 0 || d->out_type > 2) {
        qt_set_error("qt_attention_fq: unsupported shape (head_dim 64 / 128 with 128-byte-multiple rows, seq_k %% 16 == 0)");
        return QT_ERR_INVALID_ARGUMENT;
    }
    auto mis = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) != 0; };
    if (mis(d->q) || mis(d->k) || mis(d->vt) || mis(d->out) || (d->mask && mis(d->mask)) || (d->lut && mis(d->lut))) {
        qt_set_error("qt_attention_fq: pointers must be 16-byte aligned");
        return QT_ERR_UNALIGNED;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms == 0) {
        qt_set_error("qt_b200: no usable CUDA device (there is no CPU fallback)");
        return QT_ERR_CUDA;
    }
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.B = (int)d->batch;
    p.H = (int)d->heads;
    p.Sq = (int)d->seq_q;
    p.Sk = (int)d->seq_k;
    p.D = D;
    p.q_blocks = (p.Sq + AT_M - 1) / AT_M;
    p.key_blocks = (p.Sk + AT_KEYS - 1) / AT_KEYS;
    p.kq_bytes = D * esz;
    p.alpha = d->alpha;
    p.has_alpha = d->alpha != 1.0f;
    p.mask = static_cast<const __nv_bfloat16 *>(d->mask);
    p.mask_rows = d->mask ? (int)d->mask_rows : 1;
    p.mask_batches = d->mask ? (int)d->mask_batches : 1;
    p.causal = d->causal && d->mask && d->seq_q == d->seq_k;
    p.flags = d->fq_points;
    p.out_codes = d->out_type;
    p.p_codes = p_fp8 ? ((d->pv_type == QT_GEMM_E5M2 || d->pv_type == QT_GEMM_E5M2_E4M3) ? 2 : 1) : 0;
    if (d->fq_points) {
        if (!d->fmt) {
            qt_set_error("qt_attention_fq: fq_points set but fmt is NULL");
            return QT_ERR_INVALID_ARGUMENT;
        }
        int rc = qt_make_round(d->fmt, &p.rp);
        if (rc != QT_OK) return rc;
        if (p.rp.kind == QTR_INT) {
            p.fq_kind = 2;
        } else if (p.rp.kind != QTR_IDENTITY) {
            if (!d->lut || qt_lut_config(p.rp, &p.table.cfg) != QT_OK) {
                qt_set_error("qt_attention_fq: fp / posit formats need the device table from qt_lut_build_host(fmt)");
                return QT_ERR_INVALID_ARGUMENT;
            }
            p.table.table = static_cast<const QtLutEntry *>(d->lut);
            p.fq_kind = 1;
        }
    }
    if ((p.p_codes || p.out_codes) && p.fq_kind != 1) {
        qt_set_error("qt_attention_fq: fp8 codes need the e4m3 / e5m2 fake-quant step that produces them");
        return QT_ERR_INVALID_ARGUMENT;
    }
    // formats of the two products (kind::f16: bf16 = 1; kind::f8f6f4: e4m3 = 0, e5m2 = 1)
    auto fmt_a = [](int t) { return t == QT_GEMM_BF16 ? 1 : (t == QT_GEMM_E5M2 || t == QT_GEMM_E5M2_E4M3) ? 1 : 0; };
    auto fmt_b = [](int t) { return t == QT_GEMM_BF16 ? 1 : (t == QT_GEMM_E5M2 || t == QT_GEMM_E4M3_E5M2) ? 1 : 0; };
    p.idesc_s = attn_idesc(fmt_a(d->qk_type), fmt_b(d->qk_type), AT_KEYS);
    p.idesc_o = attn_idesc(fmt_a(d->pv_type), fmt_b(d->pv_type), D);

    CUtensorMap map_q, map_k, map_vt, map_o;
    int rc = make_map(&map_q, d->q, fp8, D, p.Sq, p.H, p.B, d->ld_q, d->stride_q_head, d->stride_q_batch, AT_M);
    if (rc != QT_OK) return rc;
    rc = make_map(&map_k, d->k, fp8, D, p.Sk, p.H, p.B, d->ld_k, d->stride_k_head, d->stride_k_batch, AT_KEYS);
    if (rc != QT_OK) return rc;
    rc = make_map(&map_vt, d->vt, p_fp8, p.Sk, D, p.H, p.B, p.Sk, (int64_t)D * p.Sk, (int64_t)p.H * D * p.Sk, D);
    if (rc != QT_OK) return rc;
    rc = make_map(&map_o, d->out, d->out_type != 0, D, p.Sq, p.H, p.B, d->ld_out, d->stride_out_head,
                  d->stride_out_batch, 32, d->out_type ? 64 : ROW_BYTES);
    if (rc != QT_OK) return rc;

    const size_t smem = (size_t)AT_M * p.kq_bytes + 2 * (size_t)AT_KEYS * p.kq_bytes +
                        2 * (size_t)D * AT_KEYS * (p_fp8 ? 1 : 2) + (size_t)AT_M * AT_KEYS * (p_fp8 ? 1 : 2) +
                        QT_LUT_BYTES + 8 * 16 + 16 + 1024 /* stats */ + 1024 /* alignment */;
    const unsigned grid = (unsigned)(p.B * p.H * p.q_blocks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fp8) {
        static bool done[64] = {};
        if (dev >= 64 || !done[dev]) {
            cudaFuncSetAttribute(attn_fq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (dev < 64) done[dev] = true;
        }
        attn_fq_kernel<true><<<grid, AT_THREADS, smem, st>>>(map_q, map_k, map_vt, map_o, p);
    } else {
        static bool done[64] = {};
        if (dev >= 64 || !done[dev]) {
            cudaFuncSetAttribute(attn_fq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (dev < 64) done[dev] = true;
        }
        attn_fq_kernel<false><<<grid, AT_THREADS, smem, st>>>(map_q, map_k, map_vt, map_o, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        qt_set_error("qt_attention_fq launch: %s", cudaGetErrorString(e));
        return QT_ERR_CUDA;
    }
    return QT_OK;
}
