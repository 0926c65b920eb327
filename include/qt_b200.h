/*
 * qt_b200.h -- C ABI of the B200-native fake-quant hot path (libqt_b200.so).
 *
 * Drop-in boundary for jeffreyyu0602/quantized-training.  The reference has no
 * native code; its hot path is the torch op sequence below, and these entry
 * points are what a binding for that path calls instead (reference file:line
 * cited per function; paths relative to src/quantized_training/).
 *
 * Conventions: plain pointers and sizes, no torch types.  All data pointers are
 * DEVICE pointers unless the name ends in _host.  Kernels are asynchronous on
 * `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 * Every function returns 0 (QT_OK) or a QT_ERR_* code; the message for the last
 * error on the calling thread is available from qt_last_error().  The caller
 * owns all memory.  No global mutable state: calls on distinct streams are
 * independent.  There is NO CPU fallback: compute entry points need a CUDA
 * device and fail with QT_ERR_CUDA otherwise.
 */
#ifndef QT_B200_H
#define QT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QT_OK 0
#define QT_ERR_UNSUPPORTED_DTYPE 1 /* reference: ValueError("Unsupported dtype: ...") fake_quantize.py:95 */
#define QT_ERR_INVALID_ARGUMENT 2
#define QT_ERR_CUDA 3
#define QT_ERR_UNALIGNED 4
#define QT_NO_LUT 5 /* not an error: the format has no binade-constant table and runs on the direct path */

/* size of the per-format constant table used by the fast path of qt_fq_forward (see qt_lut_build_host) */
#define QT_LUT_BYTES 8192

/* element type of x / y */
#define QT_BF16 0
#define QT_F32 1

/* format families */
#define QT_KIND_IDENTITY 0 /* "float32", "bfloat16": table is the identity */
#define QT_KIND_INT 1      /* intN / uintN            fake_quantize.py:43-52 */
#define QT_KIND_FP 2       /* e4m3/e5m2 and fpN_eXmY  fake_quantize.py:55-80, fp8.py */
#define QT_KIND_POSIT 3    /* positN_ES               fake_quantize.py:84-86, posit.py */

/* QT_KIND_FP flavours */
#define QT_FP_CUSTOM 0 /* "e4m3", "e5m2", "fp8.e4m3": fp8.py:10-67  (non-finite -> NaN, zero results are +0) */
#define QT_FP_MX 1     /* "fpN_eXmY": fp8.py:147-203 evaluated in bf16 (Inf passes, NaN band, below-half-min quirk) */

/* A parsed dtype string.  Plain data; fill it with qt_format_from_string(). */
typedef struct qt_format {
    int32_t kind;        /* QT_KIND_* */
    int32_t flavour;     /* QT_FP_* for QT_KIND_FP, else 0 */
    int32_t nbits;       /* total bits (int, posit, fpN) */
    int32_t ebits;       /* fp: exponent bits; posit: es */
    int32_t mbits;       /* fp: explicit mantissa bits */
    int32_t is_unsigned; /* uintN, or fpN_eXmY with N == X + Y (scale formats) */
    float max_value;     /* largest magnitude of the format: qmax / max_norm / maxpos */
    float min_value;     /* int: qmin (negative); others: -max_value */
} qt_format_t;

/* Library identification, e.g. "qt_b200 0.1 sm_100a". */
const char *qt_version(void);
const char *qt_last_error(void);

/* Replaces the regex dispatch of get_quantization_map(dtype) (fake_quantize.py:31-95):
 * "int8", "uint4", "e4m3", "E5M2", "fp8.e4m3", "fp8_e4m3", "fp6_e3m2", "fp4_e2m1",
 * "posit8_1", "float32", "bfloat16".  "nfK" is outside this path -> QT_ERR_UNSUPPORTED_DTYPE. */
int qt_format_from_string(const char *dtype, qt_format_t *fmt);

/* get_quant_min_max(dtype) (quantizer/quantizer.py:53-94). */
int qt_format_min_max(const char *dtype, double *qmin, double *qmax);

/* HOST evaluation of the device rounding logic on all 65 536 bf16 bit patterns:
 * table_host[i] = bf16 bits of round(bf16 value with bits i).  This is the same
 * total function the reference stores in the `qmap` buffer (fake_quantize.py:304);
 * the kernels never read such a table -- this is for API parity and for tests. */
int qt_table_host(const qt_format_t *fmt, uint16_t *table_host);

/* Fast-path constants for a format: 512 x {p1, p2, d, l} floats, one entry per (sign, exponent) of a bf16
 * input, such that  round_fmt(x) = fma(saturate(fma(|x|, p1, p2)), d, l)  on every bf16 x (derived from the
 * bitwise rounding logic and verified on all 65 536 inputs before returning).  This plays the part of the
 * reference's 128 KB `qmap` buffer (fake_quantize.py:304) at 8 KB and without a per-element gather from HBM:
 * the caller copies lut_host to the device once and passes it to qt_fq_forward.
 * Returns QT_OK, or QT_NO_LUT for formats that run on the direct path (intN, uintN, float32, bfloat16). */
int qt_lut_build_host(const qt_format_t *fmt, void *lut_host);

/* The observer half of FusedAmaxObsFakeQuantFunction.forward, minus the amax of the
 * current tensor (fake_quantize.py:225-242).  Per channel c of `channels`:
 *   amax = max_i history[i][c]  (read BEFORE the insert -> delayed scaling, NaN propagates)
 *   history = roll(history, -1, 0);  history[0][c] = 0   (slot the next kernel max-accumulates into)
 *   sf = amax / quant_max, kept only if amax > 0 and finite; optional 2^ceil(log2 sf)
 *   scale[c] = sf
 * Call it before qt_fq_forward / qt_amax with amax_out = history (slot 0). */
/* HOST evaluation of the force_scale_power_of_two rounding the device applies: 2 ** ceil(log2(sf)) as the
 * reference computes it in fp32 (fake_quantize.py:240-241), restated without log2 (see qt_round.h) -- for tests. */
float qt_scale_pow2_host(float sf);

int qt_scale_update(float *history, int amax_history_len, size_t channels, float *scale,
                    float quant_max, int force_scale_power_of_two, void *stream);

/* The fused quantize-dequantize pass (fake_quantize.py:244-246 + decomposed.py:146-163):
 *   s = scale.to(x.dtype);  y = round_fmt(x / s) * s     (each op rounded to x's dtype;
 *   fp32 inputs are first truncated to bf16 with round-to-odd, as vmap does)
 * and, in the same pass, amax_out[c] = max(amax_out[c], max|x|) (fake_quantize.py:217-223).
 * x, y: contiguous, viewed as [outer, channels, inner]; per-tensor / unobserved: channels = 1.
 * scale:    `channels` floats on the device, or NULL for an exact scale of 1 (bare specs).
 * amax_out: `channels` floats on the device (non-negative; NaN sticks), or NULL.
 * lut:      QT_LUT_BYTES on the device from qt_lut_build_host(fmt), or NULL to round on the direct
 *           bitwise path (same results, more integer instructions per element).
 * x and y must not overlap. */
int qt_fq_forward(const void *x, void *y, size_t outer, size_t channels, size_t inner, int elem_type,
                  const qt_format_t *fmt, const float *scale, float *amax_out, const void *lut, void *stream);

/* Quantize to one-byte codes for the FP8 tensor-core GEMM: codes[i] = OCP fp8 encoding (e4m3fn / e5m2) of
 * q = round_fmt(x[i] / s), the value qt_fq_forward multiplies by s; decode(code) == q exactly.  Formats: e4m3, e5m2,
 * fp8_e4m3, fp8_e5m2 (others: QT_ERR_UNSUPPORTED_DTYPE).  Per tensor only (one scale or NULL), n elements,
 * `lut` is required.  amax_out as in qt_fq_forward.  This is the fused form of the reference's
 * fake-quant (fake_quantize.py:244-246) followed by the cast a true-FP8 GEMM needs: 3 bytes of traffic per
 * bf16 element instead of 4 + 3. */
int qt_quantize_codes(const void *x, void *codes, size_t n, int elem_type, const qt_format_t *fmt,
                      const float *scale, float *amax_out, const void *lut, void *stream);

/* ---- one-byte codes of ANY <= 8-bit format: the storage form of quantized GEMM operands (qt_codes.h) ---------------
 * The reference emits codes for posits only (`return_pbits`, posit.py:60-65) and keeps every other format as bf16
 * values; the layouts here follow it for posits and define the rest subject to decode(code) == qmap[idx]
 * (SURVEY.md App. A):
 *   QT_CODE_NATIVE  positN_ES: the N-bit posit word, two's complement for negative values, sign-extended to a byte
 *                   (== the reference's pbits; NaR for NaN); intN: the two's-complement integer; uintN: the integer;
 *                   fp formats: sign | biased exponent | mantissa (OCP E4M3 / E5M2 / E3M2 / E2M3 / E2M1 layouts).
 *   QT_CODE_E4M3 / QT_CODE_E5M2  the OCP fp8 encoding of the same value, legal when every value of the format is an
 *                   e4m3 (e5m2) value -- fp6_e3m2, fp6_e2m3, fp4_e2m1, int2..int5 are subsets of e4m3 -- so those
 *                   operands run on the FP8 tensor cores (QT_GEMM_E4M3...) with no decode step.
 * Exceptions to decode(encode(q)) == q, all outside what a GEMM operand can meaningfully hold: the sign of zero; NaN
 * in formats without a NaN code (-> +0); +-Inf in formats without an Inf code (-> +-max).
 * qt_code_table_host: table256_host[b] = bf16 bits of decode(byte b) -- what the GEMM's decode warps look up
 * (QT_GEMM_CODE8); fails with QT_ERR_UNSUPPORTED_DTYPE if the format has more than 8 bits or does not fit the container.
 * qt_encode_codes_host: HOST evaluation of round_fmt + encode on bf16 bit patterns (tests, API parity).
 * qt_quantize_codes8: codes[i] = encode(round_fmt(x[i] / s)), per tensor, amax_out as in qt_fq_forward. */
#define QT_CODE_NATIVE 0
#define QT_CODE_E4M3 1
#define QT_CODE_E5M2 2
int qt_code_table_host(const qt_format_t *fmt, int code_kind, uint16_t *table256_host);
int qt_encode_codes_host(const qt_format_t *fmt, int code_kind, const uint16_t *bf16_bits_host, uint8_t *codes_host,
                         size_t n);
int qt_quantize_codes8(const void *x, void *codes, size_t n, int elem_type, const qt_format_t *fmt, int code_kind,
                       const float *scale, float *amax_out, void *stream);

/* Observer only (fake quant disabled, e.g. calibration): amax_out[c] = max(amax_out[c], max|x|). */
int qt_amax(const void *x, size_t outer, size_t channels, size_t inner, int elem_type,
            float *amax_out, void *stream);

/* ---- block-scaled qschemes: microscaling and group_wise_affine (qt_block.cu) ---------------------------------
 * One call = MXFakeQuantFunction.forward (fake_quantize.py:105-129: calculate_mx_qparam decomposed.py:372-419,
 * quantize :171-210, expand :127-140, tiling of mx_utils.py:62-121) or GroupWiseAffineFakeQuantFunction.forward
 * (fake_quantize.py:138-190).  The scale is a function of the block being quantized (no history, no delay):
 *   microscaling       s = amax(|block|) / quant_max [through the scale_fmt codebook], or with
 *                      force_scale_power_of_two 2^(floor(log2 amax) - floor(log2 quant_max)); s <= 0 or NaN -> 1;
 *                      y = round_fmt(x / s) * s
 *   group_wise_affine  sf = (max - min) / (quant_max - quant_min), <= 0 or NaN -> 1; zp = -min / sf + quant_min
 *                      [both through the scale_fmt codebook]; y = (clamp(round(x / sf + zp), qmin, qmax) - zp) * sf
 * with every intermediate rounded to the tensor's dtype as the reference's separate torch ops do.
 * x, y: contiguous, seen as [d0, n1, d1, n2, d2].  Axis n1 is tiled with block_size; axis n2 is tiled too when
 * block_axis2 != 0 (two-axis blocks, e.g. ax=(-2,-1)), else it is an ordinary axis (pass n2 = d2 = 1 for the
 * usual [outer, n, inner] view of a single tiled axis).  Edge blocks are zero padded as in the reference (this
 * matters for the min / max of the affine scheme only).
 * scale (and zero_point for the affine scheme): OUTPUT, one float per block, row-major over the block grid
 * [d0, ceil(n1/bs), d1, ceil(n2/bs) or n2, d2] -- what the reference leaves in the module's `scale` buffer.
 * fmt / lut: element format as in qt_fq_forward (ignored by the affine scheme, which rounds to integers).
 * scale_fmt: codebook of the parameters (`scale_dtype`, e.g. fp8_e5m3), or NULL.
 * pow2_table: QT_POW2_TABLE_WORDS uint32 on the device from qt_block_pow2_table_host(elem_type, ...); needed
 * when force_scale_power_of_two != 0. */
#define QT_BLOCK_MX 0
#define QT_BLOCK_AFFINE 1
#define QT_POW2_TABLE_WORDS 288
typedef struct qt_block_desc {
    const void *x;
    void *y;
    int32_t elem_type; /* QT_BF16 / QT_F32 */
    int32_t qscheme;   /* QT_BLOCK_MX / QT_BLOCK_AFFINE */
    int64_t d0, n1, d1, n2, d2;
    int32_t block_size;
    int32_t block_axis2;
    float quant_min, quant_max;
    int32_t force_scale_power_of_two;
    int32_t reserved;
    const qt_format_t *fmt;
    const void *lut;
    const qt_format_t *scale_fmt;
    const void *pow2_table;
    float *scale;
    float *zero_point;
    /* Optional: the parameter codebook as a 65 536-entry table (uint16 bf16 patterns, device) instead of scale_fmt --
     * what torch.ops.quantized_ops.calculate_mx_qparam receives as `scale_qmap` (decomposed.py:366-419).  y == NULL:
     * compute the parameters only. */
    const void *scale_table;
} qt_block_desc_t;
int qt_fq_block(const qt_block_desc_t *desc, void *stream);

/* HOST: the function  amax -> floor(log2(amax))  as the reference evaluates it in the tensor's dtype
 * (mx_utils.py:44-48: for bf16 tensors log2() is rounded to bf16 before the floor).  table_host[e], e = 1..254:
 * mantissa threshold (23-bit units) at or above which the result is e - 126 instead of e - 127;
 * table_host[256 + k], k = 0..22: the same for subnormal amax with leading bit k (threshold on the whole pattern).
 * QT_POW2_TABLE_WORDS uint32. */
int qt_block_pow2_table_host(int elem_type, uint32_t *table_host);

/* ---- operator surface: torch.ops.quantized_ops.{vmap, quantize, dequantize} (decomposed.py:143-262) ----------
 * The reference's PT2E graphs call these ops with an explicit 65 536-entry table tensor (`qmap`), so the table may
 * be ANY codebook (e.g. NF4) -- this entry point therefore gathers from the caller's table instead of using the
 * bitwise rounders.  x, y, scale, zero_point all have the element type `elem_type` (the result dtype of the
 * reference's `input / scale` after torch's type promotion; the binding promotes).
 *   QT_TABLE_QUANTIZE    u = x / s [+ zp];  y = table_a ? table_a[idx(u)] : u            (decomposed.py:171-210)
 *   QT_TABLE_DEQUANTIZE  v = table_a ? table_a[idx(x)] : x;  d = zp ? (v - zp) * s : v * s;
 *                        y = table_b ? table_b[idx(d)] : d                                (:218-262)
 * idx(v) = the bf16 bit pattern of v; fp32 values are truncated to bf16 with round-to-odd first (vmap, :146-163).
 * Tables: uint16[65536] of bf16 bit patterns on the device, or NULL.  NaN results of the arithmetic carry the bit
 * pattern the reference's CPU run produces (bf16: 0x7FC0; fp32: first NaN operand quieted, else 0xFFC00000), so a
 * table that tells NaN patterns apart is indexed identically.  scale / zero_point: one value (scalar_params != 0) or one value per block of the grid
 * [d0, ceil(n1/bs), d1, ceil(n2/bs) or n2, d2] (`expand`, decomposed.py:127-140), tensor seen as [d0,n1,d1,n2,d2]
 * like qt_fq_block. */
#define QT_TABLE_QUANTIZE 0
#define QT_TABLE_DEQUANTIZE 1
#define QT_TABLE_LOOKUP 2 /* y = table_a[idx(x)] on the raw bits (vmap): scale / zero_point unused */
typedef struct qt_table_op_desc {
    const void *x;
    void *y;
    int32_t elem_type;
    int32_t op;
    int64_t d0, n1, d1, n2, d2;
    int32_t block_size;
    int32_t block_axis2;
    int32_t scalar_params;
    int32_t reserved;
    const void *scale;
    const void *zero_point; /* or NULL */
    const void *table_a;
    const void *table_b;
} qt_table_op_desc_t;
int qt_table_op(const qt_table_op_desc_t *desc, void *stream);

/* ---- quantized GEMM / batched GEMM (tcgen05 + TMEM + TMA) -------------------------------------------
 * C[b, m, n] = epilogue(alpha * sum_k A[b, m, k] * B[b, n, k]),  fp32 accumulation, C in bf16.
 * Replaces F.linear(x_q, W_q, bias) of the QAT Linear (modules/qat/linear.py:40-41; A = activations [M, K],
 * B = weight [N, K]) and torch.matmul(q, k^T) of MatmulFunctional (modules/quantizable/functional_modules.py:
 * 22-27; A = q [B*H, S, D], B = k [B*H, S, D]).  Operands hold values of the quantized format exactly:
 * as bf16 (any format of this library with <= 8 bits), or as one-byte e4m3 / e5m2 codes (FP8 tensor cores).
 * Fused epilogue, in this order: * alpha, + bias[n] (bf16), [round] activation, [round] + residual[b, m, n] (bf16),
 * round to bf16 -- the intermediate roundings are those of the reference's separate bf16 ops, applied in registers.
 * Leading dimensions and batch strides are in elements; bases, lda/ldb (in bytes), ldc, ldr must be 16-byte
 * aligned and N % 8 == 0.  batch == 1: strides are ignored. */
#define QT_GEMM_BF16 0
#define QT_GEMM_E4M3 1      /* A and B are e4m3 codes */
#define QT_GEMM_E5M2 2      /* A and B are e5m2 codes */
#define QT_GEMM_E4M3_E5M2 3 /* A e4m3, B e5m2 */
#define QT_GEMM_E5M2_E4M3 4 /* A e5m2, B e4m3 (gradient x weight in the backward pass) */
/* Operands stored as ONE-BYTE CODES of any <= 8-bit format (posit8, int8, ...: qt_quantize_codes8, QT_CODE_NATIVE) and
 * decoded exactly to bf16 INSIDE the kernel: eight decode warps fetch the codes, look them up in a bank-conflict-free
 * copy of `code_lut` (qt_code_table_host: 256 bf16 values) and write the 128-byte-swizzled bf16 tiles the tensor cores
 * read -- weights cost 1 byte of HBM traffic instead of 2 and the products are bit-identical to the bf16-operand path.
 * CODE8_B: A bf16 (TMA), B codes [N, K]; CODE8_AB: both codes (same format).  K-major, K % 16 == 0, strides % 16 == 0,
 * plain epilogue (alpha, bias, residual).  The decode costs shared-memory bandwidth the MMA also needs: use it where
 * the product is bound by the weight traffic (small M), not for compute-bound shapes. */
#define QT_GEMM_CODE8_B 5
#define QT_GEMM_CODE8_AB 6
#define QT_ACT_NONE 0
#define QT_ACT_RELU 1
#define QT_ACT_GELU 2 /* exact erf GELU (BERT / RoBERTa) */
#define QT_ACT_SILU 3
int qt_gemm_nt(const void *A, const void *B, void *C, int operand_type, int64_t batch, int64_t M, int64_t N,
               int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB, int64_t strideC,
               float alpha, const void *bias, int activation, const void *residual, int64_t ldr, int64_t strideR,
               void *stream);

/* The same product with a two-level batch, b = outer * batch_inner + inner, each level with its own stride per
 * operand: attention on [B, S, H, D] projections without transpose copies (outer = B, inner = H, rows = S:
 * lda = H*D, strideA_inner = D, strideA_outer = S*H*D), and a context written straight into [B, S, H*D].
 * Replaces the permute + contiguous + torch.matmul chain of the quantizable attention blocks
 * (modules/quantizable/modeling_bert.py:104-168, modeling_llama.py:201-263).  Zero-initialise the struct:
 * fields added later default to "off". */
typedef struct qt_gemm_desc {
    const void *A, *B;
    void *C;
    int32_t operand_type; /* QT_GEMM_* */
    int32_t activation;   /* QT_ACT_* */
    int64_t M, N, K;
    int64_t batch_inner, batch_outer; /* >= 1 */
    int64_t lda, ldb, ldc;            /* row strides, elements */
    int64_t strideA_inner, strideA_outer, strideB_inner, strideB_outer, strideC_inner, strideC_outer;
    float alpha;
    const void *bias;     /* bf16 [N] or NULL */
    const void *residual; /* bf16, indexed like C with its own strides, or NULL */
    int64_t ldr, strideR_inner, strideR_outer;
    /* Output re-quantization in the epilogue -- the input hook of the CONSUMING module (quantize.py:116-150) applied
     * by the producer: C = fq(bf16(result)) for a bare (scale 1) spec.  fq_fmt NULL = off.  fq_lut: device table from
     * qt_lut_build_host(fq_fmt) (fp / posit formats).  out_type QT_OUT_BF16: C holds the fake-quantized bf16 values;
     * QT_OUT_E4M3 / QT_OUT_E5M2: C is a one-byte tensor of their fp8 codes (ldc and strides in elements = bytes).
     * glu != 0: B is a gate|up projection interleaved in blocks of 64 rows (64 gate features, then the same 64
     * features of up, ...; N = 2 * features, N % 128 == 0); the epilogue computes act(bf16(gate)) * bf16(up) with the
     * bf16 roundings of HF's LlamaMLP op chain and C has N / 2 columns.  activation must be QT_ACT_SILU. */
    const qt_format_t *fq_fmt;
    const void *fq_lut;
    int32_t out_type;
    int32_t glu;
    /* Causal attention schedules (M x N tiles are 128 rows high).  QT_CAUSAL_OUT_LOWER (needs M == N): output tiles
     * that lie entirely above the diagonal are neither computed nor written -- the consumer (qt_softmax_fq with
     * QT_SOFTMAX_CAUSAL) does not read them.  QT_CAUSAL_A_LOWER (needs M == K): A[m, k] is known to be zero -- or
     * unwritten, see qt_softmax_fq -- for k >= 128 * (m / 128 + 1), so the reduction of row tile mt stops there. */
    int32_t causal;
    int32_t reserved;
    /* NULL: the schedule above applies.  Otherwise a device int32 written earlier on the same stream (e.g. by
     * qt_causal_mask_check): the schedule applies only if it is non-zero, else the full product is computed -- the
     * decision is made on the device, so a forward captured in a CUDA graph stays correct when the mask changes. */
    const int32_t *causal_flag;
    /* Operand storage order.  QT_MAJOR_K (0, default): A is [M, K], B is [N, K], unit-stride K axis, lda / ldb =
     * row strides.  QT_MAJOR_MN: the operand is stored transposed -- A as [K, M], B as [K, N], unit-stride row axis,
     * lda / ldb = the stride between K lines -- and is read by the tensor cores as it lies (MN-major shared-memory
     * descriptors), so the backward products of a Linear y = x W^T need no transpose copies
     * (autograd of F.linear, modules/qat/linear.py:40-41):
     *   dgrad  gx[M, K] = g[M, N] W[N, K]      A = g  (K-major), B = W (MN-major: contraction over its row axis)
     *   wgrad  gW[N, K] = g[M, N]^T x[M, K]    A = g  (MN-major), B = x (MN-major)
     * and torch.matmul(x, y) with y [K, N] row-major is A = x, B = y (MN-major).  With one-byte operands an MN-major
     * B needs N tiles of 128 rows (chosen automatically).  The causal schedules take K-major operands. */
    int32_t a_major, b_major;
    const void *code_lut; /* QT_GEMM_CODE8*: device uint16[256], bf16 bits of decode(byte) */
    /* Block-scaled operands (OCP microscaling: one power-of-two scale per 32 elements along K), fp8 operand types
     * only: C = sum over blocks of (a_block . b_block) * 2^(ea - 127) * 2^(eb - 127) on tcgen05.mma kind::mxf8f6f4
     * .block_scale -- the product the reference's linear_mx forms by dequantizing both operands first
     * (decomposed.py:311-331).  sf_a / sf_b: UE8M0 exponent bytes packed by qt_mx_pack_scales ([K / 128][32][rows_pad /
     * 32][4]), sf_rows_a / sf_rows_b their padded row counts (multiples of 128).  NULL: plain fp8 product.
     * A K-major; B K-major or MN-major (the second operand of torch.matmul as stored: tiles of 128 columns); plain
     * epilogue (alpha, bias, residual). */
    const void *sf_a, *sf_b;
    int64_t sf_rows_a, sf_rows_b;
    /* batched products (torch.matmul of matmul_mx, decomposed.py:341-363): non-zero = the operand's scale array holds
     * one packed set per batch entry, entry = outer * batch_inner + inner; zero = one set shared by the batch. */
    int32_t sf_a_batched, sf_b_batched;
} qt_gemm_desc_t;
#define QT_MAJOR_K 0
#define QT_MAJOR_MN 1
#define QT_CAUSAL_OUT_LOWER 1
#define QT_CAUSAL_A_LOWER 2
int qt_gemm_nt_ex(const qt_gemm_desc_t *desc, void *stream);

/* Scales of one block-scaled operand, fp32 [rows, kblocks32] (kblocks32 = ceil(K / 32), every entry a power of two
 * 2^e with -126 <= e <= 127), to the byte layout the block-scaled MMA path loads: out[kb][m0][g][j] (uint8) = biased
 * exponent of scale[g * 32 + m0][kb * 4 + j], kb < ceil(kblocks32 / 4), m0 < 32, g < rows_pad / 32, j < 4; rows_pad =
 * rows rounded up to 128; entries outside the matrix are 0.  out holds ceil(kblocks32 / 4) * rows_pad * 4 bytes.
 * *ok_out (device int32, may be NULL) is AND-ed with "every scale is such a power of two". */
int qt_mx_pack_scales(const float *scale, int64_t rows, int64_t kblocks32, void *out, int32_t *ok_out, void *stream);
/* The same for `batch` matrices laid end to end (out: batch packed sets, each as above); transposed != 0: each scale
 * matrix is stored [kblocks32, rows] (the second operand of matmul_mx, scaled along its first matrix axis). */
int qt_mx_pack_scales_ex(const float *scale, int64_t batch, int64_t rows, int64_t kblocks32, int transposed, void *out,
                         int32_t *ok_out, void *stream);

/* ---- ops between the GEMMs, fused with the fake-quant steps around them (qt_fused.cu) -------------------------
 * All tensors bf16 on the device, 16-byte aligned, column counts multiples of 8.  `fmt` / `lut` as in qt_fq_forward.
 * fq_points is a bit mask of which fake-quant steps of the chain are present (= which hooks quantize() installed):
 * QT_FQ_PRE (input of the op), QT_FQ_MID (softmax only: input of nn.Softmax), QT_FQ_POST (input of the consuming GEMM).
 * Each step has its own per-tensor scale pointer (device float, NULL = bare spec, scale 1).  Every intermediate the
 * reference materialises as a bf16 tensor is rounded to bf16 at the same point. */
#define QT_FQ_PRE 1
#define QT_FQ_MID 2
#define QT_FQ_POST 4
/* qt_softmax_fq only, OR-ed into fq_points: the mask is the standard causal one of a square attention
 * (cols == mask_rows; mask[r, c] = finfo(bf16).min for c > r, 0 otherwise) and no QT_FQ_MID step exists.  Scores at
 * c > r are then not read (they may be unwritten: QT_CAUSAL_OUT_LOWER), their probabilities are exactly 0, and
 * probabilities at c >= 128 * (r / 128 + 1) are not written (consume them with QT_CAUSAL_A_LOWER); the mask tensor
 * itself is not read either.  (A NaN / +Inf score at a masked position, which would poison the row in the reference,
 * is therefore ignored: the schedule is for finite scores.)  With a non-NULL
 * causal_flag (device int32) the bit takes effect only if *causal_flag != 0. */
#define QT_SOFTMAX_CAUSAL 16
/* out_type: what the op stores.  QT_OUT_BF16: the fake-quantized values.  QT_OUT_E4M3 / QT_OUT_E5M2: their one-byte
 * fp8 codes (operands of the QT_GEMM_E4M3.. products); allowed only when the output step (QT_FQ_POST) is an unscaled
 * e4m3 / e5m2 fake quant, so that decode(code) is exactly the value the bf16 form would hold. */
#define QT_OUT_BF16 0
#define QT_OUT_E4M3 1
#define QT_OUT_E5M2 2

/* probs = fq_post(softmax(fq_mid(fq_pre(scores) * alpha + mask)))  over the last axis.
 * Replaces attn_scaling -> (+ mask) -> nn.Softmax -> av_matmul input hook (modules/quantizable/modeling_bert.py:
 * 142-158, modeling_llama.py:228-246).  scores/probs: [rows, cols] contiguous, cols <= 4096.  mask: NULL, or bf16
 * [mask_batches, mask_rows, cols] added to row r as mask[(r / rows_per_batch) % .., r % mask_rows]
 * (rows_per_batch = heads * mask_rows; mask_batches == 1 broadcasts over the batch). */
int qt_softmax_fq(const void *scores, void *probs, size_t rows, size_t cols, float alpha, const void *mask,
                  size_t rows_per_batch, size_t mask_rows, size_t mask_batches, int fq_points, int out_type,
                  const qt_format_t *fmt, const float *scale_pre, const float *scale_mid, const float *scale_post,
                  const void *lut, const int32_t *causal_flag, void *stream);

/* *flag_out = 1 if the additive bf16 mask [batches, rows, rows] is the standard causal mask (finfo(bf16).min strictly
 * above the diagonal, zero elsewhere) for every batch entry, else 0.  Asynchronous on `stream`; feed the flag to
 * qt_gemm_nt_ex (causal_flag) and qt_softmax_fq (QT_SOFTMAX_CAUSAL + causal_flag).  Replaces nothing in the
 * reference: it is what lets the kernels exploit the structure of the mask HF hands to the attention blocks
 * (modeling_llama.py:228-246) without a host round trip. */
int qt_causal_mask_check(const void *mask, size_t batches, size_t rows, int32_t *flag_out, void *stream);

/* 0 when `stream` is not being captured into a CUDA graph, otherwise the id of the capture in progress: host-side
 * caches of per-forward launches (the flag above) key on it so that every captured graph contains its own check. */
unsigned long long qt_stream_capture_id(void *stream);

/* y = fq_post(norm(fq_pre(x))), rows of `cols` <= 8192; y_raw (optional, bf16): norm(fq_pre(x)) before the output step,
 * for consumers that read the un-quantized tensor (BERT: the residual input of the next add).  kind 0: LlamaRMSNorm (x * rsqrt(mean x^2 + eps) rounded to
 * bf16, then * weight); kind 1: nn.LayerNorm (weight, bias may be NULL for no bias); kind 2: MobileBERT NoNorm
 * (x * weight + bias, two bf16 ops, no row statistics; transformers modeling_mobilebert.NoNorm).  Replaces the norm module plus the
 * input hooks of the Linear layers that read it (same tensor quantized once instead of once per consumer). */
int qt_norm_fq(const void *x, void *y, void *y_raw, size_t rows, size_t cols, int kind, const void *weight,
               const void *bias,
               float eps, int fq_points, int out_type, const qt_format_t *fmt, const float *scale_pre,
               const float *scale_post, const void *lut, void *stream);

/* The same with the residual add in front: x' = bf16(fq_a(x) + fq_b(res)), then y = fq_post(norm(fq_pre(x'))).
 * QT_FQ_RES_A / QT_FQ_RES_B: the hooks `prepare` puts on the two inputs of the block's AddFunctional (op group
 * `residual`), bare specs of `fmt`.  kind 3: no norm (the hooked add alone).  One pass replaces dense-output hook,
 * residual hook, add, norm-input hook, norm and the consumer's input hook (modeling_bert.py:187-191,
 * modeling_mobilebert.py:118-124): six launches of the reference's structure. */
#define QT_FQ_RES_A 32
#define QT_FQ_RES_B 64
int qt_add_norm_fq(const void *x, const void *res, void *y, void *y_raw, size_t rows, size_t cols, int kind,
                   const void *weight, const void *bias, float eps, int fq_points, int out_type, const qt_format_t *fmt,
                   const float *scale_pre, const float *scale_post, const void *lut, void *stream);

/* out = fq_post(act(gate) * up)  (up == NULL: fq_post(act(gate))); rows with independent strides ld_* (elements), so
 * gate and up can be the two halves of one fused projection.  activation: QT_ACT_*.  QT_FQ_PRE: the activation module's
 * own input hook (op group `activation`, bare spec) applied to gate first.  Replaces act_fn, the product and down_proj's
 * input hook of the HF MLP blocks. */
int qt_act_mul_fq(const void *gate, const void *up, void *out, size_t rows, size_t cols, size_t ld_gate, size_t ld_up,
                  size_t ld_out, int activation, int fq_points, int out_type, const qt_format_t *fmt,
                  const float *scale_post, const void *lut, void *stream);

/* Merged LoRA weight of the QAT LoRA Linear (modules/qat/lora.py:44-52), one pass over W:
 *   out = fq_post( bf16( W + bf16( bf16(Bq @ Aq) * scaling ) ) ),  Aq = fq(A), Bq = fq(B) when QT_FQ_PRE is set.
 * W / out: bf16 [N, K] contiguous (out may alias neither input); A: bf16 [r, K] (lora_A.weight); B: bf16 [N, r]
 * (lora_B.weight); K % 8 == 0.  QT_FQ_PRE fake-quantizes A and B with the BARE format (scale 1: use it only for
 * un-observed specs -- with a live observer the reference updates the amax history on each of its three calls, so the
 * caller quantizes A and B through the module and passes them in already quantized); QT_FQ_POST quantizes the merged
 * weight, scale_post NULL = bare or a frozen per-tensor scale.  Replaces clone + 3 fake-quants + matmul + mul + add. */
int qt_lora_merge_fq(const void *W, const void *A, const void *B, void *out, size_t N, size_t K, int r, float scaling,
                     int fq_points, const qt_format_t *fmt, const float *scale_post, const void *lut, void *stream);

/* Rotary position embedding + the qk_matmul input hooks, q and k (k may be NULL) in one launch:
 * out = fq(x * cos + rotate_half(x) * sin).  x: [tokens, heads, head_dim] with token stride ld (elements);
 * cos/sin: [cos_rows, head_dim], token t reads row t % cos_rows.  Replaces HF apply_rotary_pos_emb + two hooks. */
int qt_rope_fq(const void *q, void *q_out, size_t ld_q, size_t ld_q_out, int q_heads, const void *k, void *k_out,
               size_t ld_k, size_t ld_k_out, int k_heads, size_t tokens, int head_dim, const void *cos_table,
               const void *sin_table, size_t cos_rows, int fq_points, int out_type, const qt_format_t *fmt,
               const float *scale_q, const float *scale_k, const void *lut, void *stream);

/* out[b, h, d, s] = fq(v[b, s, h, d]): the values as the K-major operand of probabilities x values.
 * v: token stride ld_tok, batch stride batch_stride (elements); out contiguous [batch, heads, head_dim, seq]. */
int qt_fq_transpose(const void *v, void *out, int batch, int seq, int heads, int head_dim, size_t ld_tok,
                    size_t batch_stride, int fq_points, int out_type, const qt_format_t *fmt,
                    const float *scale_post, const void *lut, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QT_B200_H */
