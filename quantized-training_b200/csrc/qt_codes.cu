// qt_codes.cu -- quantize to one-byte codes of ANY <= 8-bit format (qt_quantize_codes8) and the 256-entry decode table
// the GEMM's decode warps read (qt_code_table_host).  Layouts: qt_codes.h.  The pass is the fake-quant pass
// (fake_quantize.py:244-246) stopped at q = round_fmt(x / s), followed by encode(q): 2 + 1 bytes per bf16 element.
// The rounding here is the direct bitwise logic of qt_round.h (no table): this pass runs once per cached weight.
#include <math.h>

#include "qt_codes.h"
#include "qt_fq_common.cuh"
#include "qt_launch.cuh"

namespace {

struct CodeParams {
    QtRound rp;
    QtCode code;
    int container;  // QT_CODE_*
};

__device__ __forceinline__ uint32_t encode_one(const CodeParams &p, uint32_t q)
{
    if (p.container == QT_CODE_E4M3) return (uint32_t)qt_encode_fp(4, 3, false, q) & 0xFFu;
    if (p.container == QT_CODE_E5M2) return (uint32_t)qt_encode_fp(5, 2, false, q) & 0xFFu;
    return (uint32_t)qt_encode_native(p.code, q) & 0xFFu;
}

template <bool F32, bool AMAX>
__global__ void __launch_bounds__(256)
codes8_kernel(const void *__restrict__ xv, uint8_t *__restrict__ y, size_t n, const __grid_constant__ CodeParams p,
              const float *__restrict__ scale, float *__restrict__ amax_out)
{
    griddep_wait();
    griddep_launch_dependents();
    ScaleBf16 sc = {1.0f, 1.0f};
    if (scale) sc = load_scale<F32>(scale, 0);
    const bool unit = sc.s == 1.0f;
    uint32_t amax = 0u;
    // 8 elements per thread per step: a 16-byte vector of bf16 (two 16-byte vectors of fp32), one 8-byte store
    const size_t nvec = n / 8;
    const bool aligned = (reinterpret_cast<uintptr_t>(xv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 7u) == 0;
    const size_t vec_elems = aligned ? nvec * 8 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i * 8 < vec_elems; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t q[8];
        if (F32) {
            const uint4 a = __ldcs(static_cast<const uint4 *>(xv) + 2 * i), b = __ldcs(static_cast<const uint4 *>(xv) + 2 * i + 1);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (AMAX) amax = max(amax, w[k] & 0x7FFFFFFFu);
                const uint32_t u = unit ? w[k] : __float_as_uint(__fdiv_rn(__uint_as_float(w[k]), sc.s));
                q[k] = qt_round_dyn(p.rp, f32_to_bf16_rto_hi(u));
            }
        } else {
            const uint4 v = __ldcs(static_cast<const uint4 *>(xv) + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t lo = w[k] << 16, hi = w[k] & 0xFFFF0000u;
                if (AMAX) amax = max(amax, max(lo & 0x7FFFFFFFu, hi & 0x7FFFFFFFu));
                q[2 * k] = qt_round_dyn(p.rp, unit ? lo : bf16_rne_hi(bf16_quotient<DIV_EXACT>(lo, sc)));
                q[2 * k + 1] = qt_round_dyn(p.rp, unit ? hi : bf16_rne_hi(bf16_quotient<DIV_EXACT>(hi, sc)));
            }
        }
        uint2 o;
        o.x = encode_one(p, q[0]) | (encode_one(p, q[1]) << 8) | (encode_one(p, q[2]) << 16) | (encode_one(p, q[3]) << 24);
        o.y = encode_one(p, q[4]) | (encode_one(p, q[5]) << 8) | (encode_one(p, q[6]) << 16) | (encode_one(p, q[7]) << 24);
        reinterpret_cast<uint2 *>(y)[i] = o;
    }
    // tail / misaligned views, element by element
    for (size_t i = vec_elems + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t u;
        if (F32) {
            const uint32_t b = static_cast<const uint32_t *>(xv)[i];
            if (AMAX) amax = max(amax, b & 0x7FFFFFFFu);
            u = f32_to_bf16_rto_hi(unit ? b : __float_as_uint(__fdiv_rn(__uint_as_float(b), sc.s)));
        } else {
            const uint32_t b = (uint32_t) static_cast<const uint16_t *>(xv)[i] << 16;
            if (AMAX) amax = max(amax, b & 0x7FFFFFFFu);
            u = unit ? b : bf16_rne_hi(bf16_quotient<DIV_EXACT>(b, sc));
        }
        y[i] = (uint8_t)encode_one(p, qt_round_dyn(p.rp, u));
    }
    if (AMAX) block_amax_commit(amax, amax_out);
}

int make_code(const qt_format_t *fmt, int code_kind, CodeParams *p)
{
    int rc = qt_make_round(fmt, &p->rp);
    if (rc != QT_OK) return rc;
    if (fmt->kind == QT_KIND_IDENTITY || fmt->nbits < 1 || fmt->nbits > 8) {
        qt_set_error("one-byte codes exist for formats of at most 8 bits (got %d bits)", fmt->nbits);
        return QT_ERR_UNSUPPORTED_DTYPE;
    }
    if (code_kind < QT_CODE_NATIVE || code_kind > QT_CODE_E5M2) {
        qt_set_error("unknown code kind %d", code_kind);
        return QT_ERR_INVALID_ARGUMENT;
    }
    p->code.kind = p->rp.kind;
    p->code.nbits = fmt->nbits;
    p->code.ebits = fmt->ebits;
    p->code.mbits = fmt->mbits;
    p->code.is_unsigned = fmt->is_unsigned;
    p->container = code_kind;
    return QT_OK;
}

uint16_t bf16_bits_of(double v)
{
    const float f = (float)v;  // exact for every <= 8-bit format of this library
    uint32_t u;
    memcpy(&u, &f, 4);
    if (v != v) return 0x7FC0u;
    return (uint16_t)(u >> 16);
}

}  // namespace

extern "C" int qt_code_table_host(const qt_format_t *fmt, int code_kind, uint16_t *table256_host)
{
    CodeParams p;
    if (!fmt || !table256_host) {
        qt_set_error("qt_code_table_host: NULL argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    int rc = make_code(fmt, code_kind, &p);
    if (rc != QT_OK) return rc;
    for (int b = 0; b < 256; ++b) {
        double v = 0.0;
        const int32_t s8 = (int8_t)b;
        if (code_kind == QT_CODE_E4M3) {
            v = qt_decode_fp(4, 3, false, b);
        } else if (code_kind == QT_CODE_E5M2) {
            v = qt_decode_fp(5, 2, false, b);
        } else if (p.code.kind == QTR_INT) {
            v = fmt->is_unsigned ? (double)b : (double)s8;
        } else if (p.code.kind == QTR_POSIT) {
            const int32_t lim = 1 << (fmt->nbits - 1);
            v = (s8 >= -lim && s8 < lim) ? qt_decode_posit(fmt->nbits, fmt->ebits, s8) : 0.0;
        } else {
            v = (b < (1 << fmt->nbits)) ? qt_decode_fp(fmt->ebits, fmt->mbits, fmt->is_unsigned != 0, b) : 0.0;
        }
        table256_host[b] = bf16_bits_of(v);
    }
    // every member of the format must survive encode -> decode (this is what makes a container code legal)
    uint16_t *full = (uint16_t *)malloc(65536 * sizeof(uint16_t));
    if (!full) return QT_ERR_INVALID_ARGUMENT;
    rc = qt_table_host(fmt, full);
    for (uint32_t i = 0; rc == QT_OK && i < 65536u; ++i) {
        const uint32_t q = (uint32_t)full[i] << 16;
        const uint32_t a = q & 0x7FFFFFFFu;
        if (a >= 0x7F800000u) continue;  // non-finite results: see qt_codes.h
        uint32_t c;
        if (code_kind == QT_CODE_E4M3) c = (uint32_t)qt_encode_fp(4, 3, false, q) & 0xFFu;
        else if (code_kind == QT_CODE_E5M2) c = (uint32_t)qt_encode_fp(5, 2, false, q) & 0xFFu;
        else c = (uint32_t)qt_encode_native(p.code, q) & 0xFFu;
        const uint16_t back = table256_host[c];
        const bool same = back == full[i] || ((back | full[i]) & 0x7FFFu) == 0u;  // the sign of zero is not kept
        if (!same) {
            free(full);
            qt_set_error("format is not representable in code kind %d: value with bf16 bits 0x%04x decodes to 0x%04x",
                         code_kind, full[i], back);
            return QT_ERR_UNSUPPORTED_DTYPE;
        }
    }
    free(full);
    return rc;
}

extern "C" int qt_encode_codes_host(const qt_format_t *fmt, int code_kind, const uint16_t *bf16_bits_host, uint8_t *codes_host,
                                    size_t n)
{
    CodeParams p;
    if (!fmt || (n && (!bf16_bits_host || !codes_host))) {
        qt_set_error("qt_encode_codes_host: NULL argument");
        return QT_ERR_INVALID_ARGUMENT;
    }
    int rc = make_code(fmt, code_kind, &p);
    if (rc != QT_OK) return rc;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t q = qt_round_dyn(p.rp, (uint32_t)bf16_bits_host[i] << 16);
        uint32_t c;
        if (code_kind == QT_CODE_E4M3) c = (uint32_t)qt_encode_fp(4, 3, false, q);
        else if (code_kind == QT_CODE_E5M2) c = (uint32_t)qt_encode_fp(5, 2, false, q);
        else c = (uint32_t)qt_encode_native(p.code, q);
        codes_host[i] = (uint8_t)(c & 0xFFu);
    }
    return QT_OK;
}

extern "C" int qt_quantize_codes8(const void *x, void *codes, size_t n, int elem_type, const qt_format_t *fmt,
                                  int code_kind, const float *scale, float *amax_out, void *stream)
{
    CodeParams p;
    if (!fmt || (n && (!x || !codes)) || (elem_type != QT_BF16 && elem_type != QT_F32)) {
        qt_set_error("qt_quantize_codes8: NULL argument or unknown element type");
        return QT_ERR_INVALID_ARGUMENT;
    }
    int rc = make_code(fmt, code_kind, &p);
    if (rc != QT_OK) return rc;
    if (num_sms() == 0) return no_device();
    if (n == 0) return QT_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = grid_for((n / 8 + 255) / 256 + 1, 8);
    const bool f32 = elem_type == QT_F32;
#define QT_CODES8(F, A) qt_launch(codes8_kernel<F, A>, dim3(grid), dim3(256), 0, st, x, static_cast<uint8_t *>(codes), n, p, scale, amax_out)
    if (f32)
        amax_out ? QT_CODES8(true, true) : QT_CODES8(true, false);
    else
        amax_out ? QT_CODES8(false, true) : QT_CODES8(false, false);
#undef QT_CODES8
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "quantize-to-codes8 kernel launch");
    return QT_OK;
}

// ----------------------------------------------------------------------------- block-scale packing (qt_mx_pack_scales)
namespace {
__global__ void mx_pack_scales_kernel(const float *__restrict__ scale, long long rows, long long kblocks32, int transposed,
                                      uint8_t *__restrict__ out, long long per_entry, long long total, long long groups,
                                      int32_t *__restrict__ ok_out)
{
    bool ok = true;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i / per_entry, o = i - e * per_entry;
        const long long j = o & 3, g = (o >> 2) % groups, m0 = ((o >> 2) / groups) & 31, kb = (o >> 2) / groups >> 5;
        const long long r = g * 32 + m0, c = kb * 4 + j;
        uint8_t v = 0;
        if (r < rows && c < kblocks32) {
            const float *m = scale + e * rows * kblocks32;
            const uint32_t b = __float_as_uint(transposed ? m[c * rows + r] : m[r * kblocks32 + c]);
            const uint32_t ex = b >> 23;  // sign bit included: a negative scale fails the range test
            ok = ok && (b & 0x007FFFFFu) == 0u && ex >= 1u && ex <= 254u;
            v = (uint8_t)ex;
        }
        out[i] = v;
    }
    if (ok_out && !ok) atomicAnd(ok_out, 0);
}
}  // namespace

extern "C" int qt_mx_pack_scales_ex(const float *scale, int64_t batch, int64_t rows, int64_t kblocks32, int transposed,
                                    void *out, int32_t *ok_out, void *stream)
{
    if (!scale || !out || batch < 1 || rows < 1 || kblocks32 < 1) {
        qt_set_error("qt_mx_pack_scales: NULL argument or empty scale matrix");
        return QT_ERR_INVALID_ARGUMENT;
    }
    if (num_sms() == 0) return no_device();
    const long long rows_pad = (rows + 127) / 128 * 128, k128 = (kblocks32 + 3) / 4;
    const long long per_entry = k128 * rows_pad * 4, total = per_entry * batch;
    const unsigned grid = grid_for((size_t)((total + 255) / 256), 8);
    mx_pack_scales_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        scale, rows, kblocks32, transposed, static_cast<uint8_t *>(out), per_entry, total, rows_pad / 32, ok_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mx_pack_scales kernel launch");
    return QT_OK;
}

extern "C" int qt_mx_pack_scales(const float *scale, int64_t rows, int64_t kblocks32, void *out, int32_t *ok_out,
                                 void *stream)
{
    return qt_mx_pack_scales_ex(scale, 1, rows, kblocks32, 0, out, ok_out, stream);
}
