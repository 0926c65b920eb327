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
// activation (ReLU / GELU-erf / SiLU), + residual, evaluated in fp32 registers but ROUNDED to bf16 wherever the
// reference's op chain materialises a bf16 tensor (after the Linear, after the activation, after the add).
//
// Kernel shape (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 0      TMA producer: 128 x 128-byte A tile + block_n x 128-byte B tile per k-block, 128-byte swizzle
//   warp 1      allocates TMEM (512 columns = two 128 x 256 fp32 accumulators); one lane issues tcgen05.mma
//               (M = 128, N = block_n, K = 32 bytes per instruction, 4 per k-block) and commits to mbarriers
//   warps 2-9   epilogue: tcgen05.ld (32 lanes x 64 columns) -> registers -> fp32 math -> bf16 -> swizzled smem
//               -> TMA store of 32 x 64 boxes (full 128-byte rows), overlapped with the MMAs of the next tile
//               through the second accumulator
// Long problems (>= 5000 k-block units of 256 x 256 x 128 B) run as CTA PAIRS instead (template PAIR): clusters of two CTAs on
// the two SMs of a TPC, one tcgen05.mma.cta_group::2 of 256 rows per instruction, each CTA holding its 128 rows of A and
// HALF of the B tile (six 32 KB stages).  +4 ... +8 % on the Llama projections (profiles/gemm_pair_r02.log).
// Tiles are walked in bands of 2048 rows so that the tiles in flight share operands in L2 (8192^3: +12 %).
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), bulk-store groups.
// block_n (a multiple of 16 up to 256) is chosen per problem on the host so that short-K batched products (attention
// scores), small-N layers and the M = 1024 Llama projections fill the 148 SMs in the fewest waves.
//
// Tried and removed in round 2 (profiles/gemm_bench_r02_c*.log): clusters of two CTAs sharing the B tile through TMA
// multicast (leader fetches, both receive; empty barriers count both CTAs' MMA commits).  Correct on every test, but
// 3 - 15 % SLOWER than independent CTAs on all shapes (1024 x 4096 x 4096: 31.3 vs 27.0 us): the pair runs in lockstep
// and one TMA stream feeds two SMs, while L2 -> SM operand traffic was not the limiter it was assumed to be.
// Also tried and removed: a ring whose depth grows for narrow tiles (8 stages of 24 KB for 64 columns out of the same
// 192 KB).  No shape got faster and the short-K ones got much slower (BERT qkv 8.8 -> 11.1 us, QK^T 14.2 -> 22.7 us,
// profiles/gemm_bench_r02_d*.log): a producer that runs many tiles ahead competes with the epilogue's stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "qt_internal.h"
#include "qt_lut.h"
#include "qt_tc.cuh"
#include "qt_launch.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int MAX_BLOCK_N = 256;  // the tile width is a launch parameter: 64, 128 or 256 columns (see pick_block_n)
constexpr int BASE_STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * ROW_BYTES;  // 16 KB
constexpr int B_STAGE_BYTES = MAX_BLOCK_N * ROW_BYTES;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
// CTA pairs hold half of the B tile each: 32 KB stages, six of them in the same 192 KB
constexpr int PAIR_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES / 2;
constexpr int PAIR_STAGES = 6;
// Block-scaled (MX) variant: tiles of at most 192 columns, so that the scale-factor columns fit into TMEM between the two
// accumulators (columns 192 ... 239: per ring stage 4 columns for A's 128 rows and 8 for B's <= 256).  Stage layout:
// A 16 KB | B 24 KB | scale factors of A 512 B | of B 1 KB | pad.
constexpr int MX_BLOCK_N = 192;
constexpr int MX_SF_OFFSET = A_STAGE_BYTES + MX_BLOCK_N * ROW_BYTES;  // 40 KB
constexpr int MX_STAGE_BYTES = MX_SF_OFFSET + 2048;                   // 42 KB (1024-byte multiple)
constexpr int MX_SF_TMEM_COL = MX_BLOCK_N;
constexpr int MX_SF_COLS_PER_STAGE = 12;
// MX = 2: the CTA computes TWO row tiles (256 x block_n) against one B tile -- 30 % fewer operand bytes per flop, which is
// what bounds this variant.  Both TMEM accumulators belong to the tile (no epilogue overlap: ~3 % of a K = 4096 tile);
// three stages of A 32 KB | B 24 KB | scale boxes 2 KB; 16 scale columns per stage (A 8, B 8).
constexpr int MX2_STAGES = 3;
constexpr int MX2_SF_OFFSET = 2 * A_STAGE_BYTES + MX_BLOCK_N * ROW_BYTES;  // 56 KB
constexpr int MX2_STAGE_BYTES = MX2_SF_OFFSET + 2048;                      // 58 KB
constexpr int MX2_SF_COLS_PER_STAGE = 16;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;  // two per TMEM lane quarter, splitting the tile's 64-column chunks between them
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
// QT_GEMM_CODE8 variant: operands stored as one-byte codes are decoded to bf16 on their way into shared memory by
// eight more warps.  Warp roles are aligned to warpgroups (setmaxnreg works per warpgroup): 0-3 producer / MMA / idle,
// 4-11 epilogue, 12-19 decode; three ring stages make room for the 32 KB conflict-free decode table.
constexpr int DEC_WARPS = 8;
constexpr int CODE_NUM_THREADS = 128 + 32 * EPI_WARPS + 32 * DEC_WARPS;
constexpr int CODE_STAGES = 3;
constexpr int CODE_PREFETCH = 8;  // k-blocks of codes the producer keeps ahead of the decode warps in L2
constexpr int CODE_LUT_BYTES = 256 * 32 * 4;  // [code][lane] 32-bit entries: lane l always reads bank l
constexpr int EPI_CHUNK_COLS = 64;                      // one TMA store box: 32 rows x 64 bf16 (128-byte rows)
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK_COLS * 2;  // 4 KB per epilogue warp
constexpr size_t SMEM_BYTES =
    (size_t)BASE_STAGES * STAGE_BYTES + (size_t)EPI_WARPS * EPI_BUF_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr size_t CODE_SMEM_BYTES = (size_t)CODE_STAGES * STAGE_BYTES + (size_t)EPI_WARPS * EPI_BUF_BYTES +
                                   CODE_LUT_BYTES + 1024 + 256;

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3 };

struct GemmParams {
    int64_t M, N, K;         // K in elements
    uint32_t batch_inner;    // batch index b = outer * batch_inner + inner (e.g. outer = sequence, inner = head)
    int k_blocks;            // ceil(K * elem_bytes / 128)
    uint32_t m_tiles, n_tiles, num_tiles;  // < 2^31 (checked by the launcher); PAIR: m_tiles counts 256-row pairs
    uint32_t group_m;        // row tiles per band of the tile walk (see tile_coord)
    int block_n;             // tile width: 64 ... 256; multiples of 16 for the plain bf16 epilogue, else of 64
    __nv_bfloat16 *c_ptr;    // C as a plain pointer (+ strides): the partial last 64-column chunk of a tile whose width
    int64_t ldc, strideC_inner, strideC_outer;  // is not a multiple of 64 is stored with ordinary 16-byte stores
    const int32_t *causal_flag;  // device flag gating `causal` (null: unconditional)
    int causal;              // 0 off; 1: tiles strictly above the diagonal are skipped (scores of a causal attention);
                             // 2: A is lower triangular (its probabilities): the K loop of row tile mt stops at its diagonal
    int debug;               // QT_GEMM_DEBUG bit mask (timing experiments only; results are wrong when set)
    int sf_a_batched, sf_b_batched;  // MX: the operand's scale factors have one set per batch entry (else one shared set)
    int a_mn, b_mn;          // operand is MN-major: stored [K, rows] with a unit-stride row axis (backward products)
    // QT_GEMM_CODE8*: operands held as one-byte codes (K-major, strides in bytes), decoded through code_lut
    int a_code, b_code;
    const uint8_t *a_codes, *b_codes;
    int64_t lda_c, strideA_inner_c, strideA_outer_c, ldb_c, strideB_inner_c, strideB_outer_c;
    const uint16_t *code_lut;  // 256 bf16 bit patterns (qt_code_table_host), device memory
    const __nv_bfloat16 *bias;      // [N] or null
    const __nv_bfloat16 *residual;  // same layout as C, or null
    int64_t ldr, strideR_inner, strideR_outer;
    float alpha;
    int act;
    uint32_t idesc;
    // output re-quantization in the epilogue (OUT_FQ / OUT_GLU kernels)
    int fq_kind;              // 0 none, 1 binade table (fp / posit formats), 2 direct integer rounding
    int out_codes;            // 0: bf16 values; 1 / 2: e4m3 / e5m2 one-byte codes (C is uint8)
    const QtLutEntry *lut;    // QT_LUT_BYTES in global memory (read through L1)
    QtLutCfg lut_cfg;
    QtRound rp;
};

enum { OUT_PLAIN = 0, OUT_FQ = 1, OUT_GLU = 2 };

// Causal schedules (p.causal): the three roles of the kernel walk the same tile sequence, so they must agree on
// which tiles exist and how many K blocks each has.
__device__ __forceinline__ bool tile_skipped(int causal, uint32_t mt, uint32_t nt, int block_n)
{
    return causal == 1 && (int64_t)nt * block_n > (int64_t)mt * 128 + 127;  // first column past the last row
}
template <bool FP8>
__device__ __forceinline__ int tile_k_blocks(const GemmParams &p, int causal, uint32_t mt)
{
    if (causal != 2) return p.k_blocks;
    // rows of tile mt are < (mt + 1) * 128 and see k < (mt + 1) * 128: 128-byte K blocks hold 128 fp8 / 64 bf16 values
    const int64_t kb = (int64_t)(mt + 1) * (FP8 ? 1 : 2);
    return kb < p.k_blocks ? (int)kb : p.k_blocks;
}

__device__ __forceinline__ float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

// Compile-time activation: a run-time switch inside the 64-way unrolled epilogue made ~77 KB of SASS whose
// skipped blocks still thrashed the instruction cache (measured: 2400 cycles per 64-column chunk for a plain store).
template <int ACT>
__device__ __forceinline__ float apply_act(float v)
{
    if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
    if (ACT == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (ACT == ACT_SILU) return __fdividef(v, 1.0f + __expf(-v));
    return v;
}

// fake quant of a value already on the bf16 grid (bare spec: scale 1), table read through L1
__device__ __forceinline__ float epi_fq(const GemmParams &p, float f)
{
    const uint32_t u = __float_as_uint(f);
    if (p.fq_kind == 1) return __uint_as_float(qt_lut_round_dyn(p.lut, p.lut_cfg, u));
    if (p.fq_kind == 2) return __uint_as_float(qt_round<QTR_INT>(p.rp, u));
    return f;
}
__device__ __forceinline__ uint32_t epi_fp8x2(float lo, float hi, int out_codes)
{
    uint32_t c = out_codes == 2 ? (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E5M2)
                                : (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, __NV_E4M3);
    const uint32_t inf_code = out_codes == 2 ? 0x7Cu : 0x7Fu;  // +-Inf survives only fpN_eXmY; keep it Inf / NaN
    const uint32_t bl = __float_as_uint(lo), bh = __float_as_uint(hi);
    if ((bl & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0xFF00u) | ((bl >> 24) & 0x80u) | inf_code;
    if ((bh & 0x7FFFFFFFu) == 0x7F800000u) c = (c & 0x00FFu) | ((((bh >> 24) & 0x80u) | inf_code) << 8);
    return c;
}
__device__ __forceinline__ void add_bf16x8(float (&f)[8], const __nv_bfloat16 *src)
{
    const uint4 bb = __ldg(reinterpret_cast<const uint4 *>(src));
    const uint32_t w[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f[2 * j] += __uint_as_float(w[j] << 16);
        f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
    }
}

// Timeline of CTA 0 (QT_GEMM_DEBUG & 32): [role][event] -> clock64.  Roles: 0 producer (after the empty wait of
// each k-block), 1 MMA (after the tmem_empty wait, after each full wait), 2 epilogue warp 2 (after the tmem_full wait,
// after the chunk's store was issued).
// Compiled only with -DQT_GEMM_TRACE (scripts/bmm_trace.py); the product build has no trace code.
#ifdef QT_GEMM_TRACE
__device__ long long qt_gemm_trace[3][256];
__device__ __forceinline__ void trace(const GemmParams &p, int role, int &slot)
{
    if ((p.debug & 32) && blockIdx.x == 0 && slot < 256) qt_gemm_trace[role][slot++] = clock64();
}
#else
__device__ __forceinline__ void trace(const GemmParams &, int, int &) {}
#endif

// ----------------------------------------------------------------------------- kernel
// ACT: activation; AUX: the problem has a bias and / or a residual (pointers checked at run time);
// OUT: OUT_PLAIN bf16 result | OUT_FQ result fake-quantized (bf16 values or fp8 codes) | OUT_GLU act(gate) * up of a
// column-interleaved gate|up projection (64 gate columns, then the 64 up columns of the same features), fake-quantized.
// PAIR: clusters of two CTAs (the two SMs of a TPC) run one 256 x block_n tile with cta_group::2 MMAs.  Each CTA loads
// its own 128 rows of A and HALF of the B tile, so the tensor core of each SM reads half the B bytes from shared
// memory per instruction (at 128 x 256 x 16 a single-CTA MMA reads 12 KB per 128 cycles, close to the 128 B / clk port).
// Rank 0 issues the MMAs; TMA bytes of both CTAs are counted on its full barriers; its commits are multicast to both
// CTAs' empty / tmem_full barriers; both CTAs' epilogue warps hand accumulators back on its tmem_empty barriers.
// MX: fp8 operands with one UE8M0 scale per 32 K elements (kind::mxf8f6f4.block_scale).  The producer also loads the
// k-block's scale-factor boxes; the MMA thread copies them into TMEM (tcgen05.cp) in front of the four MMAs that use
// them, each selecting its byte of the scale columns through the descriptor's sf ids.
template <bool FP8, int ACT, bool AUX, int OUT, bool CODE = false, bool PAIR = false, int MX = 0>
__global__ void __launch_bounds__(CODE ? CODE_NUM_THREADS : NUM_THREADS, 1)
qt_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ GemmParams p,
               const __grid_constant__ CUtensorMap map_sfa, const __grid_constant__ CUtensorMap map_sfb)
{
    extern __shared__ unsigned char smem_raw[];
    static_assert(!(CODE && PAIR), "the decode variant runs single CTAs");
    static_assert(!MX || (FP8 && !CODE && !PAIR), "block scaling: fp8 operands, single CTAs");
    constexpr int STAGES = CODE ? CODE_STAGES : PAIR ? PAIR_STAGES : MX == 2 ? MX2_STAGES : BASE_STAGES;
    constexpr int STAGE_BYTES = MX == 2 ? MX2_STAGE_BYTES : MX ? MX_STAGE_BYTES : PAIR ? PAIR_STAGE_BYTES
                                                                                   : (A_STAGE_BYTES + B_STAGE_BYTES);
    constexpr int A_BYTES = MX == 2 ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;          // A part of a stage
    constexpr int SF_OFFSET = MX == 2 ? MX2_SF_OFFSET : MX_SF_OFFSET;             // MX: scale boxes of a stage
    constexpr int SF_A_BYTES = MX == 2 ? 1024 : 512;
    constexpr int SF_COLS = MX == 2 ? MX2_SF_COLS_PER_STAGE : MX_SF_COLS_PER_STAGE;
    constexpr int EPI_WARP0 = CODE ? 4 : 2;                              // first epilogue warp
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;         // EPI_WARPS x 4 KB store staging
    const uint32_t lut_base = epi_base + EPI_WARPS * EPI_BUF_BYTES;     // CODE: [256][32] decode table
    const uint32_t bars = lut_base + (CODE ? CODE_LUT_BYTES : 0);
    // barrier slots (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int block_n = p.block_n;
    // PAIR: the cluster walks tiles of 256 rows (p.m_tiles counts row PAIRS); this CTA owns row tile 2 * pair + rank and
    // rows [rank * block_n / 2, (rank + 1) * block_n / 2) of the B tile
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t tile0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, tile_step = PAIR ? gridDim.x >> 1 : gridDim.x;
    // tile index -> (row tile of THIS CTA, column tile, batch).  Inside a batch the tiles are walked in bands of
    // p.group_m row tiles (row tile fastest, then the column tile, then the next band): the ~148 tiles in flight then
    // cover a near-square patch of C and re-read few distinct A rows / B rows from L2 instead of all of A per wave.
    auto tile_coord = [&](uint32_t tile, uint32_t &mt, uint32_t &nt, uint32_t &b) {
        const uint32_t per_batch = p.m_tiles * p.n_tiles;
        b = tile / per_batch;
        const uint32_t t = tile - b * per_batch, band = p.group_m * p.n_tiles;
        const uint32_t g = t / band, r = t - g * band;
        const uint32_t rows = min(p.group_m, p.m_tiles - g * p.group_m);
        const uint32_t m = g * p.group_m + r % rows;
        nt = r / rows;
        mt = PAIR ? m * 2u + cta_rank : MX == 2 ? m * 2u : m;   // MX == 2: the first of the tile's two row tiles
    };
    const int b_rows = PAIR ? block_n >> 1 : block_n;   // B rows in this CTA's shared memory
    // MN-major operand tiles: one TMA box = 128 bytes of rows (64 bf16 / 128 fp8) x the K lines of a k-block
    constexpr int MN_BOX_ROWS = FP8 ? 128 : 64;
    constexpr int MN_BOX_BYTES = (FP8 ? 128 : 64) * ROW_BYTES;        // K lines per k-block x 128 bytes
    constexpr int MN_K_STEP_BYTES = (FP8 ? 32 : 16) * ROW_BYTES;      // K lines per tcgen05.mma x 128 bytes

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), CODE ? 1 + DEC_WARPS : 1);  // CODE: + one arrival per decode warp
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), PAIR ? 2 * EPI_WARPS : EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {  // the same warp of both CTAs: the columns are allocated in both SMs' TMEM
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if constexpr (PAIR) {
        if (cluster_nctarank() != 2u) __trap();  // launched without the cluster attribute: the protocol below would hang
        cluster_sync_all();                      // both CTAs' barriers exist before either signals the other
    } else {
        __syncthreads();
    }
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // PDL: everything above touched only this CTA's shared memory / TMEM and the kernel parameters; the operands, the
    // causal flag and the output buffers belong to the predecessor until it has completed
    griddep_wait();
    griddep_launch_dependents();
    // every thread reads the same (already final) flag: the three roles agree on the schedule
    const int causal = (p.causal != 0 && (p.causal_flag == nullptr || *p.causal_flag != 0)) ? p.causal : 0;

    // CODE8 register budget per warpgroup (setmaxnreg, first statement of each role's branch so that ptxas allocates
    // the role's code against it).  Registers can only move between the warpgroups of this CTA: the pool is what the
    // launch allocated, 640 x 96 = 61440, and 128 x 40 + 256 x 144 + 256 x 72 = 60416 fits (asking for more
    // than the others release blocks forever).
    if (CODE && warp >= 4 + EPI_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===== decode warps (CODE8): codes (global / L2) -> registers -> table -> bf16 -> swizzled operand tile =====
        // A thread owns 16-byte code vectors (16 elements of one row): vector v of the B tile is row v / 4, 16-byte
        // chunk v % 4 of the row's 64-byte k-block; decoded it is two 16-byte chunks (2c, 2c + 1) of the row's 128-byte
        // line, stored at chunk ^ (row & 7) -- the 128-byte swizzle TMA would have applied.  Quarter-warps write 8
        // distinct chunks of two rows: conflict-free.  The vectors of k-block i + 1 are in flight while k-block i is
        // decoded.
        const int dt = threadIdx.x - 32 * (4 + EPI_WARPS);                    // 0 .. 255
        const uint32_t my_lut = lut_base + ((uint32_t)lane << 2);
        constexpr int NT = 32 * DEC_WARPS;
        constexpr int MAXV = (BLOCK_M + MAX_BLOCK_N) * 4 / NT;                // 6 vectors per thread per k-block
        const int va = p.a_code ? BLOCK_M * 4 / NT : 0, vb = p.b_code ? block_n * 4 / NT : 0;
        struct It {
            uint32_t tile;
            int kb;
        };
        auto valid = [&](const It &it) { return it.tile < p.num_tiles; };
        auto advance = [&](It &it) {
            if (++it.kb == p.k_blocks) {
                it.kb = 0;
                it.tile += gridDim.x;
            }
        };
        auto load = [&](const It &it, uint4 (&v)[MAXV]) {
            uint32_t mt, nt, b;
            tile_coord(it.tile, mt, nt, b);
            const int64_t bi = b % p.batch_inner, bo = b / p.batch_inner;
            const int64_t k0 = (int64_t)it.kb * 64;
#pragma unroll
            for (int j = 0; j < MAXV; ++j) {
                v[j] = make_uint4(0u, 0u, 0u, 0u);
                const bool is_a = j < va;
                if (!is_a && j - va >= vb) continue;
                const int idx = (is_a ? j : j - va) * NT + dt;
                const int64_t row = (is_a ? (int64_t)mt * BLOCK_M : (int64_t)nt * block_n) + (idx >> 2);
                const int64_t k = k0 + (idx & 3) * 16;
                if (k >= p.K || row >= (is_a ? p.M : p.N)) continue;          // zero codes must decode to zero: see below
                if (p.debug & 128) continue;                                  // timing experiment: no global loads
                const uint8_t *src = is_a ? p.a_codes + bo * p.strideA_outer_c + bi * p.strideA_inner_c + row * p.lda_c + k
                                          : p.b_codes + bo * p.strideB_outer_c + bi * p.strideB_inner_c + row * p.ldb_c + k;
                v[j] = __ldg(reinterpret_cast<const uint4 *>(src));
            }
        };
        auto lookup2 = [&](uint32_t w, int sel) -> uint32_t {   // two codes (bytes sel, sel + 1 of w) -> packed bf16 x 2
            uint32_t lo, hi;
            const uint32_t c0 = __byte_perm(w, 0u, 0x4440 + sel), c1 = __byte_perm(w, 0u, 0x4441 + sel);
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lo) : "r"(my_lut + (c0 << 7)));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hi) : "r"(my_lut + (c1 << 7)));
            return __byte_perm(lo, hi, 0x5410);
        };
        auto store = [&](const uint4 (&v)[MAXV], int stage) {
            const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < MAXV; ++j) {
                const bool is_a = j < va;
                if (!is_a && j - va >= vb) continue;
                const int idx = (is_a ? j : j - va) * NT + dt;
                const uint32_t r = (uint32_t)(idx >> 2), c = (uint32_t)(idx & 3);
                const uint32_t line = (is_a ? sa : sb) + r * ROW_BYTES;
                const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t d0, d1, d2, d3;
                    if (p.debug & 64) {                                       // timing experiment: no table lookups
                        d0 = w[2 * h], d1 = w[2 * h] >> 8, d2 = w[2 * h + 1], d3 = w[2 * h + 1] >> 8;
                    } else {
                        d0 = lookup2(w[2 * h], 0), d1 = lookup2(w[2 * h], 2);
                        d2 = lookup2(w[2 * h + 1], 0), d3 = lookup2(w[2 * h + 1], 2);
                    }
                    const uint32_t dst = line + (((2 * c + h) ^ (r & 7u)) << 4);
                    if (p.debug & 256) continue;                              // timing experiment: no stores
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(d0), "r"(d1), "r"(d2), "r"(d3)
                                 : "memory");
                }
            }
        };
        // the table: [code][lane] 32-bit entries (bf16 in the low half), replicated so that lane l only touches bank l.
        // One global load per thread (thread t holds entry t); each warp then broadcasts its 32 entries by shuffle and
        // writes them as conflict-free rows.  (A loop of dependent global loads here cost 30+ us per CTA.)
        {
            const uint32_t mine = p.code_lut ? (uint32_t)__ldg(p.code_lut + dt) : 0u;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, mine, i);
                const uint32_t code = (uint32_t)(dt & ~31) + (uint32_t)i;
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(my_lut + (code << 7)), "r"(v) : "memory");
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");                 // decode warps only
        It cur = {blockIdx.x, 0};
        uint4 va_regs[MAXV], vb_regs[MAXV];
        if (valid(cur)) load(cur, va_regs);
        int stage = 0;
        uint32_t phase = 0;
        bool flip = false;
        while (valid(cur)) {
            It nxt = cur;
            advance(nxt);
            if (valid(nxt)) {
                if (flip) load(nxt, va_regs); else load(nxt, vb_regs);
            }
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (flip) store(vb_regs, stage); else store(va_regs, stage);
            if (!(p.debug & 512))
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(stage));
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
            flip = !flip;
            cur = nxt;
        }
    } else if (warp < EPI_WARP0) {
      if constexpr (CODE) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, tslot = 0;
            uint32_t phase = 0;
            const uint32_t stage_tx = CODE ? (uint32_t)((p.a_code ? 0 : A_STAGE_BYTES) + (p.b_code ? 0 : block_n * ROW_BYTES))
                                           : (uint32_t)(A_STAGE_BYTES + block_n * ROW_BYTES) +
                                                 (MX ? (uint32_t)SF_A_BYTES + 512u * (uint32_t)((block_n + 127) >> 7) : 0u) +
                                                 (MX == 2 ? (uint32_t)A_STAGE_BYTES : 0u);
            for (uint32_t tile = tile0; tile < p.num_tiles; tile += tile_step) {
                uint32_t mt, nt, b;
                tile_coord(tile, mt, nt, b);
                const int bi = (int)(b % p.batch_inner), bo = (int)(b / p.batch_inner);
                if (tile_skipped(causal, mt, nt, block_n)) continue;
                const int kbn = tile_k_blocks<FP8>(p, causal, mt);
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    trace(p, 0, tslot);
                    if constexpr (PAIR) {
                        // 2 x (A tile + half B tile) land on the leader's barrier; the peer's bytes may arrive before the
                        // leader's expect_tx (the pending arrival keeps the phase open)
                        // timing experiments (results are wrong): & 1 no B loads, & 2 no loads at all
                        const bool no_a = p.debug & 2, no_b = p.debug & 3;
                        if (cta_rank == 0)
                            mbar_arrive_expect_tx(full_bar(stage), (no_a ? 0u : 2u * A_STAGE_BYTES) +
                                                                       (no_b ? 0u : (uint32_t)block_n * ROW_BYTES));
                        const uint32_t lbar = mapa_u32(full_bar(stage), 0u);
                        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
                        const int kcoord = kb * (FP8 ? ROW_BYTES : ROW_BYTES / 2);
                        const int brow0 = (int)(nt * block_n) + (int)cta_rank * b_rows;
                        if (no_a) {
                        } else if (!p.a_mn) {
                            tma_load_4d_pair(sa, &map_a, lbar, kcoord, (int)(mt * BLOCK_M), bi, bo);
                        } else {
                            for (int j = 0; j < BLOCK_M / MN_BOX_ROWS; ++j)
                                tma_load_4d_pair(sa + j * MN_BOX_BYTES, &map_a, lbar, (int)(mt * BLOCK_M) + j * MN_BOX_ROWS,
                                                 kcoord, bi, bo);
                        }
                        if (no_b) {
                        } else if (!p.b_mn) {
                            tma_load_4d_pair(sb, &map_b, lbar, kcoord, brow0, bi, bo);
                        } else {
                            for (int j = 0; j < b_rows / MN_BOX_ROWS; ++j)
                                tma_load_4d_pair(sb + j * MN_BOX_BYTES, &map_b, lbar, brow0 + j * MN_BOX_ROWS, kcoord, bi, bo);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const int kcoord = kb * (FP8 ? ROW_BYTES : ROW_BYTES / 2);
                    if constexpr (MX == 2)  // the tile's second row tile (K-major A: checked by the launcher)
                        tma_load_4d(sa + A_STAGE_BYTES, &map_a, full_bar(stage), kcoord, (int)((mt + 1) * BLOCK_M), bi, bo);
                    // K-major operand: one box, rows x 128 bytes of K.  MN-major operand: boxes of 128 bytes of rows
                    // x one k-block of K lines (64 bf16 / 128 fp8), side by side: the canonical MN-major layout
                    if (CODE && p.a_code) {
                        // fetched and decoded into place by the decode warps; this warp only pulls the code tile of a
                        // later k-block into L2 (one bulk tensor prefetch), so that their loads see L2 latency, not HBM's
                        if (kb + CODE_PREFETCH < kbn && !(p.debug & 1024))
                            tma_prefetch_4d(&map_a, (kb + CODE_PREFETCH) * 64, (int)(mt * BLOCK_M), bi, bo);
                    } else if (!p.a_mn) {
                        tma_load_4d(sa, &map_a, full_bar(stage), kcoord, (int)(mt * BLOCK_M), bi, bo);
                    } else {
                        for (int j = 0; j < BLOCK_M / MN_BOX_ROWS; ++j)
                            tma_load_4d(sa + j * MN_BOX_BYTES, &map_a, full_bar(stage),
                                        (int)(mt * BLOCK_M) + j * MN_BOX_ROWS, kcoord, bi, bo);
                    }
                    if constexpr (MX) {
                        // scale factors of this k-block: 32 x 16-byte boxes = four 32-row groups each
                        for (int j = 0; j < (MX == 2 ? 2 : 1); ++j)
                            tma_load_5d(sa + SF_OFFSET + 512 * j, &map_sfa, full_bar(stage), (int)((mt + j) * 16), 0, kb,
                                        p.sf_a_batched ? bi : 0, p.sf_a_batched ? bo : 0);
                        // B's first 32-row group is nt * block_n / 32; a box must start on a multiple of four groups (16
                        // bytes), so the load starts up to two groups early and the MMA skips those columns
                        const uint32_t g_al = (nt * (uint32_t)(block_n >> 5)) & ~3u;
                        for (int j = 0; j < (block_n + 127) >> 7; ++j)
                            tma_load_5d(sa + SF_OFFSET + SF_A_BYTES + 512 * j, &map_sfb, full_bar(stage), (int)((g_al + 4u * j) * 4u),
                                        0, kb, p.sf_b_batched ? bi : 0, p.sf_b_batched ? bo : 0);
                    }
                    if (CODE && p.b_code) {
                        if (kb + CODE_PREFETCH < kbn && !(p.debug & 1024))
                            tma_prefetch_4d(&map_b, (kb + CODE_PREFETCH) * 64, (int)(nt * block_n), bi, bo);
                    } else if (!p.b_mn) {
                        tma_load_4d(sb, &map_b, full_bar(stage), kcoord, (int)(nt * block_n), bi, bo);
                    } else {
                        for (int j = 0; j < block_n / MN_BOX_ROWS; ++j)
                            tma_load_4d(sb + j * MN_BOX_BYTES, &map_b, full_bar(stage),
                                        (int)(nt * block_n) + j * MN_BOX_ROWS, kcoord, bi, bo);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // the whole warp runs this loop in converged code; the tcgen05 wrappers elect the lane that issues (qt_tc.cuh)
        if (cta_rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0, tslot = 0;
            uint32_t acc_phase = 0;
            for (uint32_t tile = tile0; tile < p.num_tiles; tile += tile_step) {
                uint32_t mt, nt, b_unused;
                tile_coord(tile, mt, nt, b_unused);
                if (tile_skipped(causal, mt, nt, block_n)) continue;
                const int kbn = tile_k_blocks<FP8>(p, causal, mt);
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
                trace(p, 1, tslot);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BLOCK_N;   // MX == 2: acc stays 0, both halves used
                for (int kb = 0; kb < kbn; ++kb) {
                    mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
                    trace(p, 1, tslot);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t da = p.a_mn ? make_smem_desc_mn(sa, MN_BOX_BYTES) : make_smem_desc(sa);
                    const uint64_t db = p.b_mn ? make_smem_desc_mn(sb, MN_BOX_BYTES) : make_smem_desc(sb);
                    // advancing K = advancing the start address (16-byte units): 32 bytes inside the swizzle atom of
                    // a K-major tile; 16 (bf16) / 32 (fp8) K lines of 128 bytes in an MN-major tile
                    const uint64_t ka = p.a_mn ? (uint64_t)(MN_K_STEP_BYTES >> 4) : (uint64_t)(MMA_K_BYTES >> 4);
                    const uint64_t kbs = p.b_mn ? (uint64_t)(MN_K_STEP_BYTES >> 4) : (uint64_t)(MMA_K_BYTES >> 4);
                    if constexpr (MX != 0) {
                        // scale columns of this stage: A's (4 per row tile), then B's (4 per 128 rows)
                        const uint32_t sfa_t = tmem_base + MX_SF_TMEM_COL + (uint32_t)stage * SF_COLS;
                        const uint32_t sfb_t = sfa_t + (MX == 2 ? 8u : 4u);
                        const uint32_t sfb_mma = sfb_t + ((nt * (uint32_t)(block_n >> 5)) & 3u);  // see the producer
                        const uint32_t sf_smem = sa + SF_OFFSET;
                        // (copying the NEXT k-block's scale factors in front of this k-block's MMAs was tried: slower)
                        if (!((p.debug & 16384) && kb != 0)) {  // debug: timing without the per-k-block copies
                            for (int j = 0; j < (MX == 2 ? 2 : 1); ++j) tcgen05_cp_32x128b_warpx4(sfa_t + 4u * j, sf_smem + 512 * j);
                            for (int j = 0; j < (block_n + 127) >> 7; ++j)
                                tcgen05_cp_32x128b_warpx4(sfb_t + 4u * j, sf_smem + SF_A_BYTES + 512 * j);
                        }
#pragma unroll
                        for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k) {  // sf ids: byte k of the scale columns
                            const uint32_t idk = p.idesc | ((uint32_t)k << 29) | ((uint32_t)k << 4);
                            tcgen05_mma_mx(tmem_d, da + k * ka, db + k * kbs, idk, sfa_t, sfb_mma, (kb | k) != 0);
                            if constexpr (MX == 2)   // second row tile: A 16 KB further, accumulator 256 columns further
                                tcgen05_mma_mx(tmem_d + MAX_BLOCK_N, da + (A_STAGE_BYTES >> 4) + k * ka, db + k * kbs, idk,
                                               sfa_t + 4u, sfb_mma, (kb | k) != 0);
                        }
                    }
#pragma unroll
                    for (int k = 0; !MX && k < ROW_BYTES / MMA_K_BYTES; ++k) {
                        if constexpr (PAIR)
                            tcgen05_mma_pair<FP8>(tmem_d, da + k * ka, db + k * kbs, p.idesc, (kb | k) != 0);
                        else
                            tcgen05_mma<FP8>(tmem_d, da + k * ka, db + k * kbs, p.idesc, (kb | k) != 0);
                    }
                    // frees the smem slot once these MMAs have read it (PAIR: in both CTAs)
                    if constexpr (PAIR) tcgen05_commit_pair(empty_bar(stage), (uint16_t)3);
                    else tcgen05_commit(empty_bar(stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                // accumulator complete -> epilogue (PAIR: each CTA's warps read their own 128 lanes)
                if constexpr (PAIR) tcgen05_commit_pair(tmem_full_bar(acc), (uint16_t)3);
                else tcgen05_commit(tmem_full_bar(acc));
                if (MX == 2 || ++acc == 2) {   // MX == 2: one accumulator set, its barriers alternate phase every tile
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
      }
    } else {
        if constexpr (CODE) asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        // ===== epilogue warps (2..9; 4..11 in the CODE8 variant) =====
        // A warp may touch only the TMEM lanes of its quarter (warp id % 4); the two warps of a quarter take
        // alternate 64-column chunks.  Per chunk: 2 x tcgen05.ld (thread = row, registers = columns) -> fp32 math
        // -> bf16 -> the warp's 4 KB staging buffer in the 128-byte-swizzle layout -> one TMA store of the
        // 32 x 64 box (full 128-byte rows in HBM; rows / columns outside C are clipped by the TMA unit).
        const int e = warp - EPI_WARP0;
        const int quarter = warp & 3, half = e >> 2;
        const uint32_t buf = epi_base + (uint32_t)e * EPI_BUF_BYTES;
        const int chunks = (block_n + EPI_CHUNK_COLS - 1) / EPI_CHUNK_COLS;
        const int tail_cols = block_n % EPI_CHUNK_COLS;   // != 0: the last chunk is partial (OUT_PLAIN only)
        int acc = 0, tslot = 0;
        uint32_t acc_phase = 0;
        // hand-back of an accumulator: PAIR -> the leader's barrier (it counts both CTAs' epilogue warps)
        auto release_acc = [&](int a) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(a), 0u));
            else mbar_arrive(tmem_empty_bar(a));
        };
        for (uint32_t tile = tile0; tile < p.num_tiles; tile += tile_step) {
            uint32_t mt, nt, b;
            tile_coord(tile, mt, nt, b);
            const int bi = (int)(b % p.batch_inner), bo = (int)(b / p.batch_inner);
            if (tile_skipped(causal, mt, nt, block_n)) continue;
            // MX == 2: the two warps of a lane quarter take one row tile (= one accumulator half) each, all of its chunks
            const int64_t row0 = (int64_t)(mt + (MX == 2 ? half : 0)) * BLOCK_M + quarter * 32;
            const int64_t row = row0 + lane;
            // the longest wait of the kernel (a whole K loop): parked, not spinning (QT_GEMM_DEBUG & 2048: spin, for A/B)
            if (p.debug & 2048)
                mbar_wait(tmem_full_bar(acc), acc_phase);
            else
                mbar_wait_parked(tmem_full_bar(acc), acc_phase, 20000u);
            if (warp == 2 && lane == 0) trace(p, 2, tslot);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                   (uint32_t)(MX == 2 ? half : acc) * MAX_BLOCK_N;
            const __nv_bfloat16 *rrow =
                (AUX && p.residual && row < p.M)
                    ? p.residual + (int64_t)bo * p.strideR_outer + (int64_t)bi * p.strideR_inner + row * p.ldr
                    : nullptr;
            // output chunks of 64 columns; a GLU output chunk consumes two accumulator chunks (gate, up)
            const int out_chunks = OUT == OUT_GLU ? chunks >> 1 : chunks;
            const int out_block_n = OUT == OUT_GLU ? block_n >> 1 : block_n;
            const int64_t n_out_total = OUT == OUT_GLU ? p.N >> 1 : p.N;
            if (MX != 2 && half >= out_chunks) {  // narrow tiles: the second warp of the quarter has no chunk, only the hand-back
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            if (p.debug & 8192) {  // timing experiment: no epilogue work, only the hand-back
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0 && (MX == 2 || half < out_chunks)) release_acc(acc);
                if (MX == 2 || ++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
                continue;
            }
            constexpr int OC_STEP = MX == 2 ? 1 : 2;
            for (int oc = MX == 2 ? 0 : half; oc < out_chunks; oc += OC_STEP) {
                const int c = OUT == OUT_GLU ? 2 * oc : oc;  // first accumulator chunk
                const bool last = oc + OC_STEP >= out_chunks;
                uint32_t v[64];
                tmem_ld_32x32_nowait(taddr + c * EPI_CHUNK_COLS, v);
                tmem_ld_32x32_nowait(taddr + c * EPI_CHUNK_COLS + 32, v + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (warp == 2 && lane == 0) trace(p, 2, tslot);  // T1: accumulator chunk in registers
                if (OUT != OUT_GLU && last) {  // last TMEM read of this warp for this tile: hand the accumulator back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                }
                const int64_t n_in0 = (int64_t)nt * block_n + c * EPI_CHUNK_COLS;     // accumulator / bias column
                const int64_t n0 = (int64_t)nt * out_block_n + oc * EPI_CHUNK_COLS;   // output column
                uint32_t gate[OUT == OUT_GLU ? 32 : 1];
                if constexpr (OUT == OUT_GLU) {
                    // gate half: bf16(act(bf16(acc + bias))), kept packed while the up half is fetched
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) * p.alpha;
                        if (AUX && p.bias && n_in0 + g * 8 < p.N) add_bf16x8(f, p.bias + n_in0 + g * 8);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(apply_act<ACT>(bf16_round(f[2 * j])),
                                                                            apply_act<ACT>(bf16_round(f[2 * j + 1])));
                            gate[g * 4 + j] = *reinterpret_cast<const uint32_t *>(&pk);
                        }
                    }
                    tmem_ld_32x32_nowait(taddr + (c + 1) * EPI_CHUNK_COLS, v);
                    tmem_ld_32x32_nowait(taddr + (c + 1) * EPI_CHUNK_COLS + 32, v + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (last) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(acc);
                    }
                }
                uint32_t packed[32];  // bf16: 64 values; codes: the first 16 words
#pragma unroll
                for (int g = 0; g < 8; ++g) {  // eight groups of 8 columns
                    const int64_t n = n0 + g * 8;
                    const bool in_n = n < n_out_total;  // N % 8 == 0: groups are entirely in or out
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) * p.alpha;
                    if constexpr (OUT == OUT_GLU) {
                        if (AUX && p.bias && n_in0 + EPI_CHUNK_COLS + g * 8 < p.N)
                            add_bf16x8(f, p.bias + n_in0 + EPI_CHUNK_COLS + g * 8);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // bf16(gate) * bf16(up), rounded: the product op of the MLP
                            f[2 * j] = bf16_round(f[2 * j]) * __uint_as_float(gate[g * 4 + j] << 16);
                            f[2 * j + 1] = bf16_round(f[2 * j + 1]) * __uint_as_float(gate[g * 4 + j] & 0xFFFF0000u);
                        }
                    } else {
                        if (AUX && p.bias && in_n) add_bf16x8(f, p.bias + n);
                        // The reference materialises a bf16 tensor after the Linear, after the activation and after
                        // the residual add (three ATen ops); rounding at the same points keeps the fused epilogue on
                        // the reference's values instead of merely near them (a 1-ulp bf16 difference flips ~3 % of
                        // the codes of an 8-bit format downstream).
                        if (ACT != ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = apply_act<ACT>(bf16_round(f[j]));
                        }
                        if (AUX && rrow && in_n) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j]);
                            add_bf16x8(f, rrow + n);
                        }
                    }
                    if constexpr (OUT == OUT_PLAIN) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            packed[g * 4 + j] = *reinterpret_cast<const uint32_t *>(&pk);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = epi_fq(p, bf16_round(f[j]));  // the consumer's input hook
                        if (p.out_codes) {
                            packed[g * 2] = epi_fp8x2(f[0], f[1], p.out_codes) | (epi_fp8x2(f[2], f[3], p.out_codes) << 16);
                            packed[g * 2 + 1] =
                                epi_fp8x2(f[4], f[5], p.out_codes) | (epi_fp8x2(f[6], f[7], p.out_codes) << 16);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                packed[g * 4 + j] = __byte_perm(__float_as_uint(f[2 * j]), __float_as_uint(f[2 * j + 1]), 0x7632);
                        }
                    }
                }
                if (warp == 2 && lane == 0) trace(p, 2, tslot);  // T2: math done
                if constexpr (OUT == OUT_PLAIN) {
                    if (tail_cols && oc == chunks - 1) {
                        // Partial chunk of a tile whose width was chosen to fill the SMs in fewer waves (e.g. 240 columns
                        // for N = 4096 on 148 SMs): a 64-column TMA box would spill into the next tile, so this thread
                        // writes its row's valid 16-byte groups itself.
                        if (row < p.M) {
                            __nv_bfloat16 *crow = p.c_ptr + (int64_t)bo * p.strideC_outer + (int64_t)bi * p.strideC_inner +
                                                  row * p.ldc + n0;
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                if (g * 8 < tail_cols && n0 + g * 8 < p.N)
                                    *reinterpret_cast<uint4 *>(crow + g * 8) =
                                        make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
                        }
                        continue;
                    }
                }
                // the previous store from this buffer must have finished READING it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                if (warp == 2 && lane == 0) trace(p, 2, tslot);  // T3: staging buffer free
                if (OUT != OUT_PLAIN && p.out_codes) {
                    // 64-byte rows, CU_TENSOR_MAP_SWIZZLE_64B: 16-byte chunk c of row r sits at c ^ ((r >> 1) & 3)
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
                if (warp == 2 && lane == 0) trace(p, 2, tslot);  // T4: staged
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy
                __syncwarp();
                if (warp == 2 && lane == 0) trace(p, 2, tslot);  // T5: fenced
                if (lane == 0 && n0 < n_out_total && row0 < p.M) {
                    asm volatile(
                        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                            reinterpret_cast<uint64_t>(&map_c)),
                        "r"(buf), "r"((int)n0), "r"((int)row0), "r"(bi), "r"(bo)
                        : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (warp == 2 && lane == 0) trace(p, 2, tslot);
            }
            if (MX == 2 || ++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        // the staging buffers must outlive the stores' READS; the writes themselves complete with the grid (what a
        // dependent kernel's griddepcontrol.wait / the stream order waits for), so the tail does not wait for them
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    tcgen05_fence_before();
    if constexpr (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and signal its barriers
    else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                         : "memory");
    }
}

// ----------------------------------------------------------------------------- host side
// instruction descriptor (cute/arch/mma_sm100_desc.hpp, InstrDescriptor): D = F32 [4,6) = 1;
// A/B format [7,10) / [10,13): kind::f16 BF16 = 1, kind::f8f6f4 E4M3 = 0 / E5M2 = 1; both K-major;
// A / B major-ness at bit 15 / 16 (0 K-major, 1 MN-major); N >> 3 at [17,23); M >> 4 at [24,29).
uint32_t make_idesc(int a_fmt, int b_fmt, int block_n, int a_mn = 0, int b_mn = 0, bool pair = false)
{
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(a_mn != 0) << 15) |
           ((uint32_t)(b_mn != 0) << 16) | ((uint32_t)(block_n >> 3) << 17) |
           ((uint32_t)((pair ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);   // cta_group::2: M = 256 across the pair
}

// Tile width.  The persistent grid runs ceil(tiles / SMs) rounds; a round costs the larger of the tile's MMA time
// and its epilogue time, plus a per-tile hand-over.  Constants measured on B200 (scripts/bmm_probe.py):
// an MMA instruction takes N/2 cycles at N = 256 but never less than ~96 (narrow tiles are bound by the
// shared-memory reads of the A operand), the epilogue ~11.3 cycles per output column of a 128-row tile.
int pick_block_n(int64_t batch, int64_t M, int64_t N, int k_blocks, int sms, int min_bn = 64, int step = 64,
                 int tile_m = BLOCK_M)
{
    const int64_t m_tiles = (M + tile_m - 1) / tile_m;
    int best = MAX_BLOCK_N;
    double best_cost = 0.0;
    // step 64: the 256 / 128 / 64 ladder (halving).  step 16: every multiple of 16 -- the width that puts the tiles on
    // the 148 SMs in the fewest, fullest waves (N = 4096 at M = 1024: 240 columns = 144 tiles in one wave instead of
    // 128 tiles of 256; N = 12288: 224 columns = 440 tiles = 3 waves of 224 instead of 3 of 256).
    for (int bn = MAX_BLOCK_N; bn >= min_bn; bn = (step == 64 ? bn >> 1 : bn - step)) {
        const int64_t tiles = m_tiles * ((N + bn - 1) / bn) * batch;
        const double rounds = (double)((tiles + sms - 1) / sms);
        const double mma = (double)k_blocks * 4.0 * (bn / 2.0 > 96.0 ? bn / 2.0 : 96.0);
        const double epi = 11.3 * bn;
        const double cost = rounds * ((mma > epi ? mma : epi) + 500.0) + epi;
        if (bn == MAX_BLOCK_N || cost < best_cost * 0.97) {
            best = bn;
            best_cost = cost;
        }
    }
    return best;
}

thread_local bool g_pair = false;  // set by qt_gemm_nt_ex (same thread) before it dispatches: launch the CTA-pair kernel
thread_local const CUtensorMap *g_map_sfa = nullptr, *g_map_sfb = nullptr;  // block-scaled launches: scale-factor maps

template <bool FP8, int ACT, bool AUX, int OUT, bool CODE, bool PAIR, int MX = 0>
void launch_kernel(int dev, unsigned grid, cudaStream_t st, const CUtensorMap &map_a, const CUtensorMap &map_b,
                   const CUtensorMap &map_c, const GemmParams &p)
{
    static bool done[64] = {};
    constexpr size_t smem = CODE ? CODE_SMEM_BYTES : SMEM_BYTES;
    static_assert((size_t)PAIR_STAGES * PAIR_STAGE_BYTES <= (size_t)BASE_STAGES * STAGE_BYTES, "pair ring fits the same smem");
    if (dev >= 64 || !done[dev]) {
        cudaFuncSetAttribute(qt_gemm_kernel<FP8, ACT, AUX, OUT, CODE, PAIR, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
        if (dev < 64) done[dev] = true;
    }
    static_assert((size_t)BASE_STAGES * MX_STAGE_BYTES <= (size_t)BASE_STAGES * STAGE_BYTES &&
                  (size_t)MX2_STAGES * MX2_STAGE_BYTES <= (size_t)BASE_STAGES * STAGE_BYTES, "MX rings fit the same smem");
    // the scale-factor maps exist only for block-scaled launches; the other variants never touch the parameters
    const CUtensorMap &sfa = MX ? *g_map_sfa : map_c, &sfb = MX ? *g_map_sfb : map_c;
    if (PAIR)
        qt_launch_cluster(qt_gemm_kernel<FP8, ACT, AUX, OUT, CODE, PAIR, MX>, dim3(grid), dim3(NUM_THREADS), smem, st, 2u,
                          map_a, map_b, map_c, p, sfa, sfb);
    else
        qt_launch(qt_gemm_kernel<FP8, ACT, AUX, OUT, CODE, PAIR, MX>, dim3(grid),
                  dim3(CODE ? CODE_NUM_THREADS : NUM_THREADS), smem, st, map_a, map_b, map_c, p, sfa, sfb);
}

template <bool FP8, int ACT, bool AUX, int OUT = OUT_PLAIN, bool CODE = false>
void launch_variant(int dev, unsigned grid, cudaStream_t st, const CUtensorMap &map_a, const CUtensorMap &map_b,
                    const CUtensorMap &map_c, const GemmParams &p)
{
    if constexpr (!CODE) {
        if (g_pair) return launch_kernel<FP8, ACT, AUX, OUT, false, true>(dev, grid, st, map_a, map_b, map_c, p);
    }
    launch_kernel<FP8, ACT, AUX, OUT, CODE, false>(dev, grid, st, map_a, map_b, map_c, p);
}

}  // namespace

#ifdef QT_GEMM_TRACE
extern "C" int qt_gemm_debug_trace(long long *host_out)
{
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_out, qt_gemm_trace, sizeof(long long) * 3 * 256) == cudaSuccess ? QT_OK : QT_ERR_CUDA;
}
#endif

extern "C" int qt_gemm_nt_ex(const qt_gemm_desc_t *d, void *stream)
{
    if (!d) {
        qt_set_error("qt_gemm_nt_ex: NULL descriptor");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const int operand_type = d->operand_type;
    const bool b_code = operand_type == QT_GEMM_CODE8_B || operand_type == QT_GEMM_CODE8_AB;
    const bool a_code = operand_type == QT_GEMM_CODE8_AB;
    const bool fp8 = operand_type != QT_GEMM_BF16 && !b_code;
    if (operand_type < QT_GEMM_BF16 || operand_type > QT_GEMM_CODE8_AB) {
        qt_set_error("qt_gemm_nt: unknown operand_type %d", operand_type);
        return QT_ERR_INVALID_ARGUMENT;
    }
    const int64_t M = d->M, N = d->N, K = d->K, inner = d->batch_inner, outer = d->batch_outer;
    if (inner < 1 || outer < 1 || M < 1 || N < 1 || K < 1 || !d->A || !d->B || !d->C) {
        qt_set_error("qt_gemm_nt: empty problem or NULL pointer (batch=%lldx%lld M=%lld N=%lld K=%lld)",
                     (long long)outer, (long long)inner, (long long)M, (long long)N, (long long)K);
        return QT_ERR_INVALID_ARGUMENT;
    }
    const int esz = fp8 ? 1 : 2;
    const int64_t k_align = 16 / esz;
    auto misaligned = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) != 0; };
    const int a_mn = d->a_major != 0, b_mn = d->b_major != 0;
    if (b_code) {
        const bool bad_a = a_code && (d->lda % 16 || (inner > 1 && d->strideA_inner % 16) || (outer > 1 && d->strideA_outer % 16));
        if (!d->code_lut || K % 16 || d->ldb % 16 || (inner > 1 && d->strideB_inner % 16) || bad_a ||
            (outer > 1 && d->strideB_outer % 16) || a_mn || b_mn || d->causal || d->fq_fmt || d->glu ||
            d->activation != ACT_NONE) {
            qt_set_error("qt_gemm_nt: QT_GEMM_CODE8* needs code_lut (qt_code_table_host on the device), K-major code "
                         "operands with K and every stride a multiple of 16, and the plain epilogue (alpha, bias, residual)");
            return QT_ERR_INVALID_ARGUMENT;
        }
    }
    const bool mx = d->sf_a != nullptr || d->sf_b != nullptr;
    if (mx) {
        if (!d->sf_a || !d->sf_b || !fp8 || a_mn || d->causal || d->fq_fmt || d->glu ||
            d->activation != ACT_NONE || d->sf_rows_a % 128 || d->sf_rows_b % 128 || d->sf_rows_a < M || d->sf_rows_b < N ||
            (reinterpret_cast<uintptr_t>(d->sf_a) | reinterpret_cast<uintptr_t>(d->sf_b)) & 15u) {
            qt_set_error("qt_gemm_nt: block-scaled products take fp8 operands (A K-major), both scale arrays "
                         "(qt_mx_pack_scales, row counts padded to 128) and the plain epilogue (alpha, bias, residual)");
            return QT_ERR_INVALID_ARGUMENT;
        }
    }
    if ((d->a_major & ~1) || (d->b_major & ~1) || ((a_mn || b_mn) && d->causal)) {
        qt_set_error("qt_gemm_nt: a_major / b_major are QT_MAJOR_K or QT_MAJOR_MN; the causal schedules take K-major "
                     "operands");
        return QT_ERR_INVALID_ARGUMENT;
    }
    const bool bad_ab = d->lda % k_align || d->ldb % k_align || (inner > 1 && (d->strideA_inner % k_align ||
                        d->strideB_inner % k_align)) || (outer > 1 && (d->strideA_outer % k_align ||
                        d->strideB_outer % k_align));
    const int c_align = d->out_type ? 16 : 8;
    const bool bad_c = N % 8 || d->ldc % c_align || (inner > 1 && d->strideC_inner % c_align) ||
                       (outer > 1 && d->strideC_outer % c_align) || (d->out_type && (d->glu ? N / 2 : N) % 16);
    const bool bad_r = d->residual && (misaligned(d->residual) || d->ldr % 8 || (inner > 1 && d->strideR_inner % 8) ||
                                       (outer > 1 && d->strideR_outer % 8));
    if (bad_ab || bad_c || bad_r || misaligned(d->A) || misaligned(d->B) || misaligned(d->C) ||
        (d->bias && misaligned(d->bias))) {
        qt_set_error("qt_gemm_nt: operands need 16-byte aligned bases, leading dimensions and batch strides, N %% 8 == 0");
        return QT_ERR_UNALIGNED;
    }
    if (d->activation < ACT_NONE || d->activation > ACT_SILU) {
        qt_set_error("qt_gemm_nt: unknown activation %d", d->activation);
        return QT_ERR_INVALID_ARGUMENT;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms == 0) {
        qt_set_error("qt_b200: no usable CUDA device (there is no CPU fallback)");
        return QT_ERR_CUDA;
    }
    const int64_t batch = inner * outer;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    // output re-quantization / gated product
    const bool glu = d->glu != 0;
    const bool requant = d->fq_fmt != nullptr;
    const int64_t n_out = glu ? N / 2 : N;
    if (d->out_type < 0 || d->out_type > 2 || (d->out_type != 0 && !requant) || (glu && (N % 128 || d->residual)) ||
        (glu && d->activation == ACT_NONE) || (requant && !glu && d->activation != ACT_NONE)) {
        qt_set_error("qt_gemm_nt: output options: fp8 codes need fq_fmt; glu needs N %% 128 == 0, an activation and no "
                     "residual; fq_fmt without glu takes no activation");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (requant) {
        int rc2 = qt_make_round(d->fq_fmt, &p.rp);
        if (rc2 != QT_OK) return rc2;
        const bool e4m3 = d->fq_fmt->kind == QT_KIND_FP && d->fq_fmt->ebits == 4 && d->fq_fmt->mbits == 3 && !d->fq_fmt->is_unsigned;
        const bool e5m2 = d->fq_fmt->kind == QT_KIND_FP && d->fq_fmt->ebits == 5 && d->fq_fmt->mbits == 2 && !d->fq_fmt->is_unsigned;
        if ((d->out_type == 1 && !e4m3) || (d->out_type == 2 && !e5m2)) {
            qt_set_error("qt_gemm_nt: fp8 code output needs an e4m3 / e5m2 fq_fmt of the same kind");
            return QT_ERR_INVALID_ARGUMENT;
        }
        if (p.rp.kind == QTR_INT) {
            p.fq_kind = 2;
        } else if (p.rp.kind == QTR_IDENTITY) {
            p.fq_kind = 0;
        } else {
            if (!d->fq_lut || (reinterpret_cast<uintptr_t>(d->fq_lut) & 15u) || qt_lut_config(p.rp, &p.lut_cfg) != QT_OK) {
                qt_set_error("qt_gemm_nt: fq_fmt of an fp / posit format needs the 16-byte aligned device table (fq_lut)");
                return QT_ERR_INVALID_ARGUMENT;
            }
            p.fq_kind = 1;
            p.lut = static_cast<const QtLutEntry *>(d->fq_lut);
        }
        p.out_codes = d->out_type;
    }
    const int c_esz = p.out_codes ? 1 : 2;
    p.M = M;
    p.N = N;
    p.K = K;
    p.batch_inner = (uint32_t)inner;
    p.k_blocks = (int)((K * esz + ROW_BYTES - 1) / ROW_BYTES);
    p.a_mn = a_mn;
    p.b_mn = b_mn;
    p.a_code = a_code;
    p.b_code = b_code;
    p.a_codes = static_cast<const uint8_t *>(d->A);
    p.b_codes = static_cast<const uint8_t *>(d->B);
    p.lda_c = d->lda;
    p.strideA_inner_c = inner > 1 ? d->strideA_inner : 0;
    p.strideA_outer_c = outer > 1 ? d->strideA_outer : 0;
    p.ldb_c = d->ldb;
    p.strideB_inner_c = inner > 1 ? d->strideB_inner : 0;
    p.strideB_outer_c = outer > 1 ? d->strideB_outer : 0;
    p.code_lut = static_cast<const uint16_t *>(d->code_lut);
    // an MN-major B tile is made of boxes of 128 bytes of rows: 64 bf16 rows, 128 fp8 rows
    // any multiple of 16 for the plain bf16 epilogue with a K-major B; the other variants keep whole 64-column chunks
    const bool fine_bn = !glu && !requant && !b_mn && !b_code && !(getenv("QT_GEMM_COARSE_TILES") != nullptr);
    // CTA pairs (cta_group::2): 256-row tiles on the two SMs of a TPC.  Measured on B200 (profiles/gemm_pair_r02.log):
    // +4 ... +8 % on the long problems (Llama qkv / down / lm_head), -2 ... -5 % on problems that last under ~30 us (cluster
    // launch and the two cluster-wide syncs are a fixed cost), so the choice is by work: 256 x 256 x 64 k-block units.
    // The causal schedules and the decode variant stay on single CTAs.  QT_GEMM_PAIR=0 / 1 forces the choice (A/B, tests).
    const bool pair_ok = !b_code && !d->causal && sms % 2 == 0 && !mx;
    const int64_t pairs256 = ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((N + MAX_BLOCK_N - 1) / MAX_BLOCK_N) * batch;
    bool pair = pair_ok && M > BLOCK_M && pairs256 * p.k_blocks >= 5000;  // measured: 4096 units (o-proj) lose 3-5 %, 5504 (fp8 down-proj) gain 6 %
    if (const char *pe = getenv("QT_GEMM_PAIR")) pair = pair_ok && pe[0] == '1';
    g_pair = pair;
    if (pair) {
        // each CTA holds block_n / 2 rows of B: whole 128-byte row boxes for an MN-major B (64 bf16 / 128 fp8 rows)
        const int min_bn = b_mn ? (fp8 ? 256 : 128) : glu ? 128 : 64;
        p.block_n = pick_block_n(batch, M, N, p.k_blocks, sms / 2, min_bn, fine_bn ? 16 : 64, 2 * BLOCK_M);
    } else
    p.block_n = pick_block_n(batch, M, N, p.k_blocks, sms, (glu || (b_mn && fp8)) ? 128 : 64, fine_bn ? 16 : 64);
    // two row tiles per CTA (MX = 2) whenever the problem has them: fewer operand bytes per flop (QT_GEMM_MX1=1: old form)
    const bool mx2 = mx && M > BLOCK_M && getenv("QT_GEMM_MX1") == nullptr;
    if (mx) {  // 192 / 128 / 64 columns: the same cost model on the widths the scale-factor columns leave room for
        const int64_t m_t = (M + (mx2 ? 2 : 1) * BLOCK_M - 1) / ((mx2 ? 2 : 1) * BLOCK_M);
        double best_cost = 0.0;
        for (int bn = b_mn ? 128 : MX_BLOCK_N; bn >= (b_mn ? 128 : 64); bn -= 64) {  // MN-major fp8 B: 128-row boxes
            const double rounds = (double)((m_t * ((N + bn - 1) / bn) * batch + sms - 1) / sms);
            const double mma = (mx2 ? 2.0 : 1.0) * (double)p.k_blocks * 4.0 * (bn / 2.0 > 96.0 ? bn / 2.0 : 96.0);
            const double epi = 11.3 * bn;
            const double cost = mx2 ? rounds * (mma + epi + 500.0) : rounds * ((mma > epi ? mma : epi) + 500.0) + epi;
            if (best_cost == 0.0 || cost < best_cost * 0.97) {
                p.block_n = bn;
                best_cost = cost;
            }
        }
    }
    p.c_ptr = static_cast<__nv_bfloat16 *>(d->C);
    p.ldc = d->ldc;
    p.strideC_inner = inner > 1 ? d->strideC_inner : 0;
    p.strideC_outer = outer > 1 ? d->strideC_outer : 0;
    p.causal = d->causal;
    p.causal_flag = d->causal_flag;
    if (p.causal < 0 || p.causal > 2 || (p.causal == 1 && M != N) || (p.causal == 2 && M != K)) {
        qt_set_error("qt_gemm_nt: causal = 1 needs a square output (M == N), causal = 2 a square A (M == K); got "
                     "causal=%d M=%lld N=%lld K=%lld", p.causal, (long long)M, (long long)N, (long long)K);
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (const char *bn = getenv("QT_GEMM_BN")) {  // tests: force a tile width (multiple of 16) where the epilogue allows it
        const int v = atoi(bn);
        if (fine_bn && !mx && v >= 16 && v <= MAX_BLOCK_N && v % 16 == 0) p.block_n = v;
        if (mx && !b_mn && (v == 64 || v == 128 || v == 192)) p.block_n = v;
    }
    if (const char *dbg = getenv("QT_GEMM_DEBUG")) {  // timing experiments: 4 / 8 / 16 force the tile width
        p.debug = atoi(dbg);
        if ((p.debug & 4) && !(pair && b_mn && fp8)) p.block_n = 128;
        if ((p.debug & 8) && !glu && !(b_mn && (fp8 || pair))) p.block_n = 64;
        if ((p.debug & 16) && !mx) p.block_n = 256;
    }
    const int tile_m = (pair || mx2) ? 2 * BLOCK_M : BLOCK_M;   // p.m_tiles counts 256-row tiles in pair / MX = 2 mode
    const int64_t m_tiles = (M + tile_m - 1) / tile_m, n_tiles = (N + p.block_n - 1) / p.block_n;
    if (m_tiles * n_tiles * batch >= (int64_t)1 << 31 || M >= (int64_t)1 << 31 || N >= (int64_t)1 << 31 ||
        inner >= (int64_t)1 << 31 || outer >= (int64_t)1 << 31) {
        qt_set_error("qt_gemm_nt: problem too large (%lld tiles)", (long long)(m_tiles * n_tiles * batch));
        return QT_ERR_INVALID_ARGUMENT;
    }
    p.m_tiles = (uint32_t)m_tiles;
    p.n_tiles = (uint32_t)n_tiles;
    p.num_tiles = (uint32_t)(m_tiles * n_tiles * batch);
    p.group_m = (uint32_t)(2048 / tile_m);   // bands of 2048 rows; QT_GEMM_GROUP_M overrides (tile rows; 0 = one band)
    if (const char *gm = getenv("QT_GEMM_GROUP_M")) p.group_m = (uint32_t)atoi(gm);
    if (p.group_m == 0 || p.group_m > p.m_tiles) p.group_m = p.m_tiles;

    CUtensorMap map_a, map_b, map_c;
    // K-major: [rows, K] boxes of rows x 128 bytes of K.  MN-major: [K, rows] boxes of one k-block of K lines x 128
    // bytes of rows; K lines / rows past the end read as zero either way.
    const int k_lines = fp8 ? 128 : 64;
    int rc = make_map(&map_c, d->C, c_esz == 1, n_out, M, inner, outer, d->ldc, d->strideC_inner, d->strideC_outer, 32,
                      c_esz == 1 ? 64 : ROW_BYTES);
    if (rc != QT_OK) return rc;
    map_a = map_b = map_c;
    // code operands: the decode warps fetch them with plain loads; the maps serve the producer's L2 prefetches
    if (a_code) {
        rc = make_map(&map_a, d->A, true, K, M, inner, outer, d->lda, d->strideA_inner, d->strideA_outer, BLOCK_M, 64);
        if (rc != QT_OK) return rc;
    }
    if (b_code) {
        rc = make_map(&map_b, d->B, true, K, N, inner, outer, d->ldb, d->strideB_inner, d->strideB_outer, p.block_n, 64);
        if (rc != QT_OK) return rc;
    }
    if (!a_code) {
        rc = a_mn ? make_map(&map_a, d->A, fp8, M, K, inner, outer, d->lda, d->strideA_inner, d->strideA_outer, k_lines)
                  : make_map(&map_a, d->A, fp8, K, M, inner, outer, d->lda, d->strideA_inner, d->strideA_outer, BLOCK_M);
        if (rc != QT_OK) return rc;
    }
    if (!b_code) {
        rc = b_mn ? make_map(&map_b, d->B, fp8, N, K, inner, outer, d->ldb, d->strideB_inner, d->strideB_outer, k_lines)
                  : make_map(&map_b, d->B, fp8, K, N, inner, outer, d->ldb, d->strideB_inner, d->strideB_outer,
                             pair ? p.block_n / 2 : p.block_n);
        if (rc != QT_OK) return rc;
    }

    CUtensorMap map_sfa, map_sfb;
    if (mx) {
        const int64_t k128 = (K + 127) / 128;
        // sf_batched_*: one scale set per batch entry (activations / both operands of a batched matmul) or one shared
        p.sf_a_batched = d->sf_a_batched != 0 && batch > 1;
        p.sf_b_batched = d->sf_b_batched != 0 && batch > 1;
        rc = make_sf_map(&map_sfa, d->sf_a, d->sf_rows_a, k128, inner, outer, p.sf_a_batched != 0);
        if (rc == QT_OK) rc = make_sf_map(&map_sfb, d->sf_b, d->sf_rows_b, k128, inner, outer, p.sf_b_batched != 0);
        if (rc != QT_OK) return rc;
        g_map_sfa = &map_sfa;
        g_map_sfb = &map_sfb;
    }
    p.bias = static_cast<const __nv_bfloat16 *>(d->bias);
    p.residual = static_cast<const __nv_bfloat16 *>(d->residual);
    p.ldr = d->ldr;
    p.strideR_inner = d->strideR_inner;
    p.strideR_outer = d->strideR_outer;
    p.alpha = d->alpha;
    p.act = d->activation;
    switch (operand_type) {
    case QT_GEMM_BF16:
    case QT_GEMM_CODE8_B:
    case QT_GEMM_CODE8_AB: p.idesc = make_idesc(1, 1, p.block_n, a_mn, b_mn, pair); break;
    case QT_GEMM_E4M3: p.idesc = make_idesc(0, 0, p.block_n, a_mn, b_mn, pair); break;
    case QT_GEMM_E5M2: p.idesc = make_idesc(1, 1, p.block_n, a_mn, b_mn, pair); break;
    case QT_GEMM_E4M3_E5M2: p.idesc = make_idesc(0, 1, p.block_n, a_mn, b_mn, pair); break;
    default: p.idesc = make_idesc(1, 0, p.block_n, a_mn, b_mn, pair); break;  // QT_GEMM_E5M2_E4M3
    }
    if (mx)  // block-scaled descriptor: no accumulator-format field (bits 4-5 are B's sf id), scale format E8M0 at bit 23
        p.idesc = (p.idesc & ~(3u << 4)) | (1u << 23);
    const unsigned grid = pair ? 2u * (p.num_tiles < (uint32_t)(sms / 2) ? p.num_tiles : (unsigned)(sms / 2))
                               : (p.num_tiles < (uint32_t)sms ? p.num_tiles : (unsigned)sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aux = d->bias != nullptr || d->residual != nullptr;
    if (mx2) {
        aux ? launch_kernel<true, ACT_NONE, true, OUT_PLAIN, false, false, 2>(dev, grid, st, map_a, map_b, map_c, p)
            : launch_kernel<true, ACT_NONE, false, OUT_PLAIN, false, false, 2>(dev, grid, st, map_a, map_b, map_c, p);
    } else if (mx) {
        aux ? launch_kernel<true, ACT_NONE, true, OUT_PLAIN, false, false, 1>(dev, grid, st, map_a, map_b, map_c, p)
            : launch_kernel<true, ACT_NONE, false, OUT_PLAIN, false, false, 1>(dev, grid, st, map_a, map_b, map_c, p);
    } else if (b_code || ((p.debug & 4096) && operand_type == QT_GEMM_BF16 && !glu && !requant && d->activation == ACT_NONE)) {
        aux ? launch_variant<false, ACT_NONE, true, OUT_PLAIN, true>(dev, grid, st, map_a, map_b, map_c, p)
            : launch_variant<false, ACT_NONE, false, OUT_PLAIN, true>(dev, grid, st, map_a, map_b, map_c, p);
    } else if (glu) {
        if (d->activation != ACT_SILU) {
            qt_set_error("qt_gemm_nt: the gated epilogue is built for SiLU (Llama-style MLP)");
            return QT_ERR_INVALID_ARGUMENT;
        }
        if (fp8)
            aux ? launch_variant<true, ACT_SILU, true, OUT_GLU>(dev, grid, st, map_a, map_b, map_c, p)
                : launch_variant<true, ACT_SILU, false, OUT_GLU>(dev, grid, st, map_a, map_b, map_c, p);
        else
            aux ? launch_variant<false, ACT_SILU, true, OUT_GLU>(dev, grid, st, map_a, map_b, map_c, p)
                : launch_variant<false, ACT_SILU, false, OUT_GLU>(dev, grid, st, map_a, map_b, map_c, p);
    } else if (requant) {
        if (fp8)
            aux ? launch_variant<true, ACT_NONE, true, OUT_FQ>(dev, grid, st, map_a, map_b, map_c, p)
                : launch_variant<true, ACT_NONE, false, OUT_FQ>(dev, grid, st, map_a, map_b, map_c, p);
        else
            aux ? launch_variant<false, ACT_NONE, true, OUT_FQ>(dev, grid, st, map_a, map_b, map_c, p)
                : launch_variant<false, ACT_NONE, false, OUT_FQ>(dev, grid, st, map_a, map_b, map_c, p);
    } else {
        const int variant = (fp8 ? 8 : 0) + d->activation * 2 + (aux ? 1 : 0);
        switch (variant) {
#define QT_GEMM_CASE(F, A, X)                                           \
    case (F ? 8 : 0) + A * 2 + (X ? 1 : 0):                             \
        launch_variant<F, A, X>(dev, grid, st, map_a, map_b, map_c, p); \
        break;
            QT_GEMM_CASE(false, ACT_NONE, false) QT_GEMM_CASE(false, ACT_NONE, true)
            QT_GEMM_CASE(false, ACT_RELU, false) QT_GEMM_CASE(false, ACT_RELU, true)
            QT_GEMM_CASE(false, ACT_GELU, false) QT_GEMM_CASE(false, ACT_GELU, true)
            QT_GEMM_CASE(false, ACT_SILU, false) QT_GEMM_CASE(false, ACT_SILU, true)
            QT_GEMM_CASE(true, ACT_NONE, false) QT_GEMM_CASE(true, ACT_NONE, true)
            QT_GEMM_CASE(true, ACT_RELU, false) QT_GEMM_CASE(true, ACT_RELU, true)
            QT_GEMM_CASE(true, ACT_GELU, false) QT_GEMM_CASE(true, ACT_GELU, true)
            QT_GEMM_CASE(true, ACT_SILU, false) QT_GEMM_CASE(true, ACT_SILU, true)
#undef QT_GEMM_CASE
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        qt_set_error("qt_gemm_nt launch: %s", cudaGetErrorString(e));
        return QT_ERR_CUDA;
    }
    return QT_OK;
}

extern "C" int qt_gemm_nt(const void *A, const void *B, void *C, int operand_type, int64_t batch, int64_t M, int64_t N,
                          int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                          int64_t strideC, float alpha, const void *bias, int activation, const void *residual,
                          int64_t ldr, int64_t strideR, void *stream)
{
    qt_gemm_desc_t d;
    memset(&d, 0, sizeof(d));
    d.A = A;
    d.B = B;
    d.C = C;
    d.operand_type = operand_type;
    d.M = M;
    d.N = N;
    d.K = K;
    d.batch_inner = batch;
    d.batch_outer = 1;
    d.lda = lda;
    d.ldb = ldb;
    d.ldc = ldc;
    d.strideA_inner = strideA;
    d.strideB_inner = strideB;
    d.strideC_inner = strideC;
    d.alpha = alpha;
    d.bias = bias;
    d.activation = activation;
    d.residual = residual;
    d.ldr = ldr;
    d.strideR_inner = strideR;
    return qt_gemm_nt_ex(&d, stream);
}
