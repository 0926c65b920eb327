// qt_gemm.cu -- quantized GEMM / batched GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM,
// fp32 accumulation, fused epilogue.
//
//   C[b, m, n] = epilogue( alpha * sum_k A[b, m, k] * B[b, n, k] )        ("NT": both operands K-major)
//
// which is F.linear(x, W) (reference modules/qat/linear.py:40-41: A = fake-quantized activations [M, K],
// B = fake-quantized weight [N, K]) and the attention score product q k^T (functional_modules.py:22-27).
// Operands hold values of the low-precision format exactly:
//   * bf16 storage (every <= 8-bit format of this library is exactly representable in bf16) -> kind::f16 MMA
//   * e4m3 / e5m2 one-byte codes -> kind::f8f6f4 MMA at twice the rate
// Epilogue (the paper's fusion levels, README table / SURVEY App. B): * alpha (attention scaling), + bias,
// activation (ReLU / GELU-erf / SiLU), + residual, then one rounding to bf16.
//
// Kernel shape (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer: 128 x 128-byte A tile + 256 x 128-byte B tile per k-block, 128-byte swizzle
//   warp 1      allocates TMEM (512 columns = two 128 x 256 fp32 accumulators); one lane issues tcgen05.mma
//               (M = 128, N = 256, K = 32 bytes per instruction, 4 per k-block) and commits to mbarriers
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns at a time -> registers -> math -> 64-byte stores,
//               overlapped with the MMAs of the next tile through the second accumulator
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "qt_internal.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int ROW_BYTES = 128;  // one k-block of a row: 64 bf16 or 128 fp8, = the swizzle span
constexpr int MMA_K_BYTES = 32;  // K extent of one tcgen05.mma in bytes (16 bf16 / 32 fp8)
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * ROW_BYTES;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * ROW_BYTES;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3 };

struct GemmParams {
    int64_t batch, M, N, K;  // K in elements
    int k_blocks;            // ceil(K * elem_bytes / 128)
    int64_t m_tiles, n_tiles, num_tiles;
    __nv_bfloat16 *C;
    int64_t ldc, strideC;
    const __nv_bfloat16 *bias;      // [N] or null
    const __nv_bfloat16 *residual;  // same layout as C, or null
    int64_t ldr, strideR;
    float alpha;
    int act;
    uint32_t idesc;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();  // try_wait itself blocks for a while; this is many seconds
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool FP8>
__device__ __forceinline__ void tcgen05_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    if (FP8)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// 32 TMEM lanes (one per thread of the warp) x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte rows, 128-byte swizzle (what TMA wrote):
// canonical UMMA layout ((8, n), 2) : ((8 x 16 B, SBO), 16 B) with SBO = 8 rows x 128 B = 1024 B.
// Fields (cute/arch/mma_sm100_desc.hpp): start address >> 4 [0,14), LBO >> 4 [16,30) (ignored for swizzled
// K-major; 1 like CUTLASS), SBO >> 4 [32,46), version = 1 [46,48), layout type SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float apply_act(float v, int act)
{
    switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_SILU: return v / (1.0f + __expf(-v));
    default: return v;
    }
}

// ----------------------------------------------------------------------------- kernel
template <bool FP8>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qt_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ GemmParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
    // barrier slots (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int64_t mt = tile % p.m_tiles, rest = tile / p.m_tiles;
                const int64_t nt = rest % p.n_tiles, b = rest / p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                    const int kcoord = kb * (FP8 ? ROW_BYTES : ROW_BYTES / 2);
                    tma_load_3d(sa, &map_a, full_bar(stage), kcoord, (int)(mt * BLOCK_M), (int)b);
                    tma_load_3d(sb, &map_b, full_bar(stage), kcoord, (int)(nt * BLOCK_N), (int)b);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * BLOCK_N;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k) {
                        // advancing K inside the swizzle atom = advancing the start address (16-byte units)
                        const uint64_t koff = (uint64_t)((k * MMA_K_BYTES) >> 4);
                        tcgen05_mma<FP8>(tmem_d, da + koff, db + koff, p.idesc, (kb | k) != 0);
                    }
                    tcgen05_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tcgen05_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int64_t mt = tile % p.m_tiles, rest = tile / p.m_tiles;
            const int64_t nt = rest % p.n_tiles, b = rest / p.n_tiles;
            const int64_t row = mt * BLOCK_M + quarter * 32 + lane;
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * BLOCK_N;
            __nv_bfloat16 *crow = p.C + b * p.strideC + row * p.ldc;
            const __nv_bfloat16 *rrow = p.residual ? p.residual + b * p.strideR + row * p.ldr : nullptr;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);  // warp-collective: every lane participates
                const int64_t n0 = nt * BLOCK_N + c * 32;
                if (row < p.M && n0 < p.N) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {  // four 16-byte groups of 8 columns
                        const int64_t n = n0 + g * 8;
                        if (n >= p.N) break;  // N % 8 == 0: groups are entirely in or out
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) * p.alpha;
                        if (p.bias) {
                            const uint4 bb = __ldg(reinterpret_cast<const uint4 *>(p.bias + n));
                            const uint32_t w[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[2 * j] += __uint_as_float(w[j] << 16);
                                f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                            }
                        }
                        if (p.act != ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], p.act);
                        }
                        if (rrow) {
                            const uint4 rr = __ldg(reinterpret_cast<const uint4 *>(rrow + n));
                            const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[2 * j] += __uint_as_float(w[j] << 16);
                                f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                            }
                        }
                        uint4 o;
                        uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            ow[j] = *reinterpret_cast<const uint32_t *>(&pk);
                        }
                        *reinterpret_cast<uint4 *>(crow + n) = o;
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(tmem_empty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// [batch, rows, K] K-major operand: dims {K, rows, batch}, box {128 bytes of K, box_rows, 1}, 128-byte swizzle.
// Out-of-bounds elements (K tail, row tail) are filled with zeros by TMA.
int make_operand_map(CUtensorMap *map, const void *ptr, bool fp8, int64_t K, int64_t rows, int64_t batch, int64_t ld,
                     int64_t stride, int box_rows)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        qt_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return QT_ERR_CUDA;
    }
    const int esz = fp8 ? 1 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)(batch > 1 ? stride : rows * ld) * esz};
    cuuint32_t box[3] = {(cuuint32_t)(ROW_BYTES / esz), (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                    const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qt_set_error("cuTensorMapEncodeTiled failed with CUresult %d (K=%lld rows=%lld batch=%lld ld=%lld)", (int)r,
                     (long long)K, (long long)rows, (long long)batch, (long long)ld);
        return QT_ERR_INVALID_ARGUMENT;
    }
    return QT_OK;
}

// instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D = F32 [4,6) = 1;
// A/B format [7,10) / [10,13): kind::f16 BF16 = 1, kind::f8f6f4 E4M3 = 0 / E5M2 = 1; both K-major;
// N >> 3 at [17,23); M >> 4 at [24,29).
uint32_t make_idesc(int a_fmt, int b_fmt)
{
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
           ((uint32_t)(BLOCK_M >> 4) << 24);
}

}  // namespace

extern "C" int qt_gemm_nt(const void *A, const void *B, void *C, int operand_type, int64_t batch, int64_t M, int64_t N,
                          int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                          int64_t strideC, float alpha, const void *bias, int activation, const void *residual,
                          int64_t ldr, int64_t strideR, void *stream)
{
    const bool fp8 = operand_type != QT_GEMM_BF16;
    if (operand_type < QT_GEMM_BF16 || operand_type > QT_GEMM_E5M2_E4M3) {
        qt_set_error("qt_gemm_nt: unknown operand_type %d", operand_type);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (batch < 1 || M < 1 || N < 1 || K < 1 || !A || !B || !C) {
        qt_set_error("qt_gemm_nt: empty problem or NULL pointer (batch=%lld M=%lld N=%lld K=%lld)", (long long)batch,
                     (long long)M, (long long)N, (long long)K);
        return QT_ERR_INVALID_ARGUMENT;
    }
    const int esz = fp8 ? 1 : 2;
    const int64_t k_align = 16 / esz;
    auto misaligned = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) != 0; };
    if (lda % k_align || ldb % k_align || strideA % k_align || strideB % k_align || N % 8 || ldc % 8 || strideC % 8 ||
        misaligned(A) || misaligned(B) || misaligned(C) || (bias && misaligned(bias)) ||
        (residual && (misaligned(residual) || ldr % 8 || strideR % 8))) {
        qt_set_error("qt_gemm_nt: operands need 16-byte aligned bases and leading dimensions, N %% 8 == 0");
        return QT_ERR_UNALIGNED;
    }
    if (activation < ACT_NONE || activation > ACT_SILU) {
        qt_set_error("qt_gemm_nt: unknown activation %d", activation);
        return QT_ERR_INVALID_ARGUMENT;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms == 0) {
        qt_set_error("qt_b200: no usable CUDA device (there is no CPU fallback)");
        return QT_ERR_CUDA;
    }
    CUtensorMap map_a, map_b;
    int rc = make_operand_map(&map_a, A, fp8, K, M, batch, lda, strideA, BLOCK_M);
    if (rc != QT_OK) return rc;
    rc = make_operand_map(&map_b, B, fp8, K, N, batch, ldb, strideB, BLOCK_N);
    if (rc != QT_OK) return rc;

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.batch = batch;
    p.M = M;
    p.N = N;
    p.K = K;
    p.k_blocks = (int)((K * esz + ROW_BYTES - 1) / ROW_BYTES);
    p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    p.num_tiles = p.m_tiles * p.n_tiles * batch;
    p.C = static_cast<__nv_bfloat16 *>(C);
    p.ldc = ldc;
    p.strideC = strideC;
    p.bias = static_cast<const __nv_bfloat16 *>(bias);
    p.residual = static_cast<const __nv_bfloat16 *>(residual);
    p.ldr = ldr;
    p.strideR = strideR;
    p.alpha = alpha;
    p.act = activation;
    switch (operand_type) {
    case QT_GEMM_BF16: p.idesc = make_idesc(1, 1); break;
    case QT_GEMM_E4M3: p.idesc = make_idesc(0, 0); break;
    case QT_GEMM_E5M2: p.idesc = make_idesc(1, 1); break;
    case QT_GEMM_E4M3_E5M2: p.idesc = make_idesc(0, 1); break;
    default: p.idesc = make_idesc(1, 0); break;  // QT_GEMM_E5M2_E4M3
    }
    const unsigned grid = (unsigned)(p.num_tiles < sms ? p.num_tiles : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (fp8) {
        static bool done[64] = {};
        if (dev >= 64 || !done[dev]) {
            cudaFuncSetAttribute(qt_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
            if (dev < 64) done[dev] = true;
        }
        qt_gemm_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_a, map_b, p);
    } else {
        static bool done[64] = {};
        if (dev >= 64 || !done[dev]) {
            cudaFuncSetAttribute(qt_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
            if (dev < 64) done[dev] = true;
        }
        qt_gemm_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(map_a, map_b, p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        qt_set_error("qt_gemm_nt launch: %s", cudaGetErrorString(e));
        return QT_ERR_CUDA;
    }
    return QT_OK;
}
