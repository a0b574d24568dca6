// SRS on-disk format -> device-resident base sets.
// [UP] halo2_proofs/src/poly/kzg/commitment.rs `ParamsKZG::{read_custom, write_custom}` with halo2_proofs::SerdeFormat and
// [UP] halo2curves 0.3.x src/derive/curve.rs (`GroupEncoding::{to_bytes, from_bytes}`, `SerdeObject::{read_raw, write_raw}`);
// the reference loads `params/kzg_bn254_{k}.srs` at src/scaffold.rs:119,174,271 -- once per proof in prove_private.
// File layout: k (u32 LE) | g[0..2^k) | g_lagrange[0..2^k) | g2 | s_g2, every point in one of
//   Processed          : compressed, 32 bytes = canonical little-endian x with (y & 1) << 7 in byte 31; all-zero = identity
//   RawBytes           : x | y as 2 x 4 u64 Montgomery limbs (the in-memory layout), each limb vector < p and the point on the curve
//   RawBytesUnchecked  : the same bytes, no checks
// Upstream decompresses with `parallelize` (one square root = a 254-bit exponentiation per point); here one thread per
// point does the same arithmetic and the decoded points stay on the device as a registered base set (with window tables),
// so the per-proof re-read of scaffold.rs:174 costs one file read instead of 2^(k+1) square roots plus an upload.
#include "common.h"
#include "ec.cuh"

namespace h2b {

__device__ __forceinline__ bool fq_limbs_below_p(const Fq& a) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a.l[i] < FpParams<FQ>::P(i)) return true;
        if (a.l[i] > FpParams<FQ>::P(i)) return false;
    }
    return false;
}
__device__ __forceinline__ Fq fq_curve_b() {           // 3 in Montgomery form
    const Fq one = fp_one<FQ>();
    return fp_add(fp_dbl(one), one);
}
// a^((p + 1) / 4): the square root candidate for p = 3 mod 4 ([UP] halo2curves Fq::sqrt)
__device__ __noinline__ Fq fq_sqrt_candidate(const Fq& a) {
    uint32_t e[8];
    uint64_t c = 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c += FpParams<FQ>::P(i); e[i] = (uint32_t)c; c >>= 32; }        // p + 1 (no carry out: p < 2^254)
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = (e[i] >> 2) | (i < 7 ? e[i + 1] << 30 : 0);
    Fq r = fp_one<FQ>();
    bool started = false;
#pragma unroll 1
    for (int i = 253; i >= 0; --i) {
        if (started) r = fp_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) {
            r = started ? fp_mul(r, a) : a;
            started = true;
        }
    }
    return r;
}

__device__ __forceinline__ void note_invalid(unsigned long long* first_bad, size_t i) {
#ifdef H2B_EMU
    if (*first_bad > i) *first_bad = i;
#else
    atomicMin(first_bad, (unsigned long long)i);
#endif
}

// GroupEncoding::from_bytes for G1Affine
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint4* __restrict__ in, size_t n, uint4* __restrict__ out, unsigned long long* first_bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq x = fp_load<FQ>(in + 2 * i);
    const uint32_t ysign = x.l[7] >> 31;
    x.l[7] &= 0x7fffffffu;
    Affine p;
    p.x = fp_zero<FQ>(); p.y = fp_zero<FQ>();
    bool ok = fq_limbs_below_p(x);                       // Fq::from_bytes rejects non-canonical x
    if (ok && !(fp_is_zero(x) && !ysign)) {              // x = 0 without the sign bit is the identity
        const Fq xm = fp_to_mont(x);
        const Fq rhs = fp_add(fp_mul(fp_sqr(xm), xm), fq_curve_b());
        Fq y = fq_sqrt_candidate(rhs);
        ok = fp_eq(fp_sqr(y), rhs);
        const uint32_t sign = fp_from_mont(y).l[0] & 1;
        if (sign ^ ysign) y = fp_neg(y);
        p.x = xm; p.y = y;
    }
    if (!ok) { note_invalid(first_bad, i); p.x = fp_zero<FQ>(); p.y = fp_zero<FQ>(); }
    affine_store(out + 4 * i, p);
}

// SerdeObject::read_raw + is_on_curve for G1Affine (RawBytes); the bytes already are the in-memory layout
__global__ void __launch_bounds__(128) g1_check_raw_kernel(const uint4* __restrict__ in, size_t n, unsigned long long* first_bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = affine_load(in + 4 * i);
    bool ok = fq_limbs_below_p(p.x) && fq_limbs_below_p(p.y);
    if (ok && !affine_is_identity(p)) ok = fp_eq(fp_sub(fp_sqr(p.y), fp_mul(fp_sqr(p.x), p.x)), fq_curve_b());
    if (!ok) note_invalid(first_bad, i);
}

// GroupEncoding::to_bytes for G1Affine
__global__ void __launch_bounds__(128) g1_compress_kernel(const uint4* __restrict__ in, size_t n, uint4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = affine_load(in + 4 * i);
    Fq x = fp_zero<FQ>();
    if (!affine_is_identity(p)) {
        x = fp_from_mont(p.x);
        x.l[7] |= (fp_from_mont(p.y).l[0] & 1) << 31;
    }
    fp_store<FQ>(out + 2 * i, x);
}

static const unsigned long long NO_BAD = ~0ull;

int g1_decode_run(DeviceCtx& ctx, const void* d_bytes, size_t n, int format, void* d_out, uint64_t* first_invalid, cudaStream_t stream) {
    if (format < H2B_SERDE_PROCESSED || format > H2B_SERDE_RAW_BYTES_UNCHECKED) { set_error("g1_decode: unknown format %d", format); return H2B_ERR_BAD_ARGUMENT; }
    if (first_invalid) *first_invalid = NO_BAD;
    if (n == 0) return H2B_OK;
    if (!d_bytes || !d_out) { set_error("g1_decode: null pointer"); return H2B_ERR_BAD_ARGUMENT; }
    H2B_TRY(ctx.srs_status.reserve(8));
    unsigned long long* d_bad = (unsigned long long*)ctx.srs_status.p;
    H2B_CUDA(cudaMemsetAsync(d_bad, 0xff, 8, stream));
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (format == H2B_SERDE_PROCESSED) {
        H2B_LAUNCH(g1_decompress_kernel, grid, 128, 0, stream, (const uint4*)d_bytes, n, (uint4*)d_out, d_bad);
    } else {
        if (format == H2B_SERDE_RAW_BYTES) H2B_LAUNCH(g1_check_raw_kernel, grid, 128, 0, stream, (const uint4*)d_bytes, n, d_bad);
        if (d_out != d_bytes) H2B_CUDA(cudaMemcpyAsync(d_out, d_bytes, n * 64, cudaMemcpyDeviceToDevice, stream));
    }
    H2B_CUDA(cudaGetLastError());
    if (first_invalid) {
        unsigned long long bad = NO_BAD;
        H2B_CUDA(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, stream));
        H2B_CUDA(cudaStreamSynchronize(stream));
        *first_invalid = bad;
        if (bad != NO_BAD) { set_error("g1_decode: point %llu is not a valid encoding of a curve point", bad); return H2B_ERR_BAD_ARGUMENT; }
    }
    return H2B_OK;
}

int g1_encode_run(DeviceCtx& ctx, const void* d_affine, size_t n, void* d_out_bytes, cudaStream_t stream) {
    (void)ctx;
    if (n == 0) return H2B_OK;
    if (!d_affine || !d_out_bytes) { set_error("g1_encode: null pointer"); return H2B_ERR_BAD_ARGUMENT; }
    H2B_LAUNCH(g1_compress_kernel, (unsigned)((n + 127) / 128), 128, 0, stream, (const uint4*)d_affine, n, (uint4*)d_out_bytes);
    H2B_CUDA(cudaGetLastError());
    return H2B_OK;
}

}  // namespace h2b
