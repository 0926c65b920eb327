// qt_tc.cuh -- tcgen05 / TMEM / TMA / mbarrier plumbing shared by the GEMM (qt_gemm.cu) and the attention kernel
// (qt_attn.cu): PTX wrappers, the shared-memory matrix descriptor of a K-major 128-byte-swizzled tile, and the host
// helper that builds 4-D tensor maps.  Anonymous namespace: one copy per translation unit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "qt_internal.h"

namespace {

constexpr int ROW_BYTES = 128;   // one k-block of a row: 64 bf16 or 128 fp8, = the swizzle span
constexpr int MMA_K_BYTES = 32;  // K extent of one tcgen05.mma in bytes (16 bf16 / 32 fp8)

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
// The same wait with a suspend-time hint: the hardware may park the thread for up to `hint_ns` (it is woken when the
// phase completes).  Without the hint try_wait returns almost at once and the loop spins at full issue rate -- ncu showed
// 37 % of all warp samples of the CODE8 GEMM in these loops (profiles/ncu_r02_code8.txt).  For LONG waits only
// (an epilogue warp waiting for a whole K loop).
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t hint_ns)
{
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(hint_ns)
            : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}
// non-blocking: has the phase of this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// pull a box into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar)
{
    // Executed by the WHOLE MMA warp in converged code; one elected lane issues.  (A loop body under `if (lane == 0)`
    // made every operand a per-thread register that has to be broadcast into uniform registers inside a divergence loop
    // -- ELECT / R2UR.BROADCAST / BRA.U.ANY, ~25 instructions per MMA on the single thread that paces the tensor pipe.)
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
        : "memory");
}
// ----- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run ONE 256-row MMA.  Each holds its own
// 128 rows of A, half of the B tile and its 128 lanes of the accumulator; the leader (cluster rank 0) issues.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared::cta offset of this kernel's layout) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA's share of a pair's stage: data into THIS CTA's shared memory, bytes counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1,
                                                 int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// arrives on the barrier at the same offset in every CTA of `mask` once the pair's MMAs issued so far have completed
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar, uint16_t mask)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n" ::"r"(bar),
        "h"(mask)
        : "memory");
}
template <bool FP8>
__device__ __forceinline__ void tcgen05_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate)
{
    if (FP8)
        asm volatile(
            "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// ----- block-scaled fp8 (kind::mxf8f6f4.block_scale): one UE8M0 scale per 32 K elements of a row, held in TMEM.
// Scale factors of a 128-element k-block: a 32-bit TMEM column per group of 32 rows (lane = row % 32, replicated over
// the four lane quarters), byte j = the j-th 32-element block; the instruction descriptor's sf ids pick the byte.
template <int DUMMY = 0>
__device__ __forceinline__ void tcgen05_mma_mx(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// 32 rows x 16 bytes of shared memory (rows 16 bytes apart) -> 4 TMEM columns, broadcast to the four lane quarters
__device__ __forceinline__ void tcgen05_cp_32x128b_warpx4(uint32_t tmem_dst, uint32_t smem_addr)
{
    // no-swizzle K-major matrix descriptor: core matrices of 8 rows x 16 bytes, 128 bytes apart (SBO); version 1
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;\n\t}\n" ::"r"(tmem_dst),
                 "l"(d)
                 : "memory");
}
// 5-D tiled load without swizzle (scale-factor boxes: bytes, row in group, k-block, batch inner, batch outer)
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
template <bool FP8>
__device__ __forceinline__ void tcgen05_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    if (FP8)
        asm volatile(
            "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// 32 TMEM lanes (one per thread of the warp) x 32 consecutive 32-bit columns
// (asynchronous: the registers are valid after tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t *v)
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

// MN-major operand tile (the operand is stored [K, rows] with a unit-stride row axis): TMA boxes of 128 bytes of rows
// x `k lines`, 128-byte swizzle, placed side by side `box_bytes` apart.  Canonical UMMA layout (units of 16 bytes)
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)): 8 K lines of 128 bytes form a swizzle atom (SBO = 1024 B between atoms
// along K), the next 128 bytes of rows start LBO = box_bytes further (cute/atom/mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t box_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((box_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ----------------------------------------------------------------------------- host: tensor maps
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

// QT_TMA_L2PROMO = 0 / 1 / 2 / 3: none / 64 B / 128 B / 256 B (default) L2 promotion of the maps (A/B runs)
inline CUtensorMapL2promotion l2_promotion()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QT_TMA_L2PROMO");
        v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
    }
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
         : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// [outer, inner, rows, cols] tensor with a unit-stride last axis and arbitrary (16-byte aligned) strides on the other
// three: dims {cols, rows, inner, outer}, box {box_cols, box_rows, 1, 1}, 128-byte swizzle (box_cols * esz == 128).
// Loads: out-of-bounds elements (K tail, row tail) are filled with zeros.  Stores: they are not written.
int make_map(CUtensorMap *map, const void *ptr, bool one_byte, int64_t cols, int64_t rows, int64_t inner, int64_t outer,
             int64_t ld, int64_t stride_inner, int64_t stride_outer, int box_rows, int box_bytes = ROW_BYTES)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        qt_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return QT_ERR_CUDA;
    }
    // cuTensorMapEncodeTiled is a DRIVER entry point: it needs a context current on the calling thread.  Runtime
    // calls bind the primary context lazily, and a thread that has only ever inherited its device from a guard --
    // autograd's backward worker, where dgrad / wgrad are launched -- may not have one yet (CUDA_ERROR_INVALID_CONTEXT).
    // cudaSetDevice (CUDA 12) initialises and binds the primary context of the device; once per thread and device.
    {
        static thread_local int bound_device = -1;
        int dev = -1;
        if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_device) {
            cudaSetDevice(dev);
            cudaFree(nullptr);
            bound_device = dev;
        }
    }
    const int esz = one_byte ? 1 : 2;
    // size-1 axes never move: give them a harmless, valid stride
    if (inner <= 1) stride_inner = rows * ld;
    if (outer <= 1) stride_outer = (inner <= 1 ? rows * ld : inner * stride_inner);
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)stride_inner * esz, (cuuint64_t)stride_outer * esz};
    cuuint32_t box[4] = {(cuuint32_t)(box_bytes / esz), (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, one_byte ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                    const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qt_set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%lld rows=%lld batch=%lldx%lld ld=%lld "
                     "strides=%lld,%lld)", (int)r, (long long)cols, (long long)rows, (long long)outer, (long long)inner,
                     (long long)ld, (long long)stride_inner, (long long)stride_outer);
        return QT_ERR_INVALID_ARGUMENT;
    }
    return QT_OK;
}

// Packed scale factors of one operand (qt_mx_pack_scales): bytes [batch][K128][32][groups * 4]; a box is 32 rows x 16
// bytes (four 32-row groups) of one 128-element k-block -- exactly the 32x128b source of tcgen05.cp.  The batch is
// inner-major like the operands (entry = outer * inner_count + inner); an operand shared by the whole batch (the weight
// of a Linear) has batched = false and ignores the batch coordinates.
int make_sf_map(CUtensorMap *map, const void *ptr, int64_t rows_pad, int64_t k128, int64_t inner_count, int64_t outer_count,
                bool batched)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        qt_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return QT_ERR_CUDA;
    }
    const cuuint64_t inner = (cuuint64_t)(rows_pad / 32) * 4;
    const cuuint64_t per_entry = inner * 32 * (cuuint64_t)k128;
    cuuint64_t dims[5] = {inner, 32, (cuuint64_t)k128, (cuuint64_t)inner_count, (cuuint64_t)outer_count};
    // a shared operand has batch extents 1 (the kernel then passes batch coordinates 0); strides stay valid multiples of 16
    cuuint64_t strides[4] = {inner, inner * 32, batched ? per_entry : inner * 32,
                             batched ? per_entry * (cuuint64_t)inner_count : inner * 32};
    if (!batched) dims[3] = dims[4] = 1;
    cuuint32_t box[5] = {16, 32, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qt_set_error("cuTensorMapEncodeTiled (scale factors) failed with CUresult %d (rows_pad=%lld k128=%lld)", (int)r,
                     (long long)rows_pad, (long long)k128);
        return QT_ERR_INVALID_ARGUMENT;
    }
    return QT_OK;
}

}  // namespace
