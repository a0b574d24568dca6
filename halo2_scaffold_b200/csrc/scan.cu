// Grand-product building blocks on device-resident Fr columns: the permutation and lookup
// arguments of the prover ([UP] halo2_proofs/src/plonk/permutation/prover.rs, plonk/lookup/prover.rs) build z(X) as the
// running product of numerator[i] / denominator[i]; the denominators are inverted with ff::BatchInvert (zeros stay
// zero) and z(omega^i) is the exclusive prefix product.  Both are one pass over the column on the CPU; here:
//   * batch inversion: Montgomery's trick per thread over K strided elements (3 multiplications per element + one Fermat
//     inversion per thread), prefix products parked in a scratch column;
//   * exclusive prefix product: chunk products of 16 consecutive elements, recursive scan of the chunk products (one CTA
//     Hillis-Steele at the bottom), then every chunk replays its elements from its scanned offset.
#include "common.h"

namespace h2b {

// ---- batch inversion -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fr_batch_invert_kernel(uint4* __restrict__ a, uint4* __restrict__ pre, size_t n, uint32_t K, uint32_t G) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G) return;
    Fr run = fp_one<FR>();
    for (uint32_t j = 0; j < K; ++j) {
        const size_t idx = (size_t)j * G + t;
        if (idx >= n) break;
        Fr x = fp_load<FR>(a + 2 * idx);
        fp_store<FR>(pre + 2 * idx, run);
        if (!fp_is_zero(x)) run = fp_mul(run, x);
    }
    Fr inv = fp_inv(run);
    for (uint32_t j = K; j-- > 0;) {
        const size_t idx = (size_t)j * G + t;
        if (idx >= n) continue;
        Fr x = fp_load<FR>(a + 2 * idx);
        if (fp_is_zero(x)) continue;
        Fr p = fp_load<FR>(pre + 2 * idx);
        fp_store<FR>(a + 2 * idx, fp_mul(inv, p));
        inv = fp_mul(inv, x);
    }
}

int fr_batch_invert_run(DeviceCtx& ctx, void* d_a, size_t n, cudaStream_t stream) {
    if (n == 0) return H2B_OK;
    if (!d_a) { set_error("batch_invert: null pointer"); return H2B_ERR_BAD_ARGUMENT; }
    H2B_TRY(ctx.scan_scratch.reserve(n * 32));
    // K elements per thread: enough threads to fill the GPU, at most 128 so that the Fermat inversion (about 380
    // multiplications) costs 3 more multiplications per element
    const size_t resident = (size_t)ctx.sm_count * 1024;
    size_t K = n / resident;
    if (K < 8) K = 8;
    if (K > 128) K = 128;
    const uint32_t G = (uint32_t)((n + K - 1) / K);
    H2B_LAUNCH(fr_batch_invert_kernel, (G + 127) / 128, 128, 0, stream, (uint4*)d_a, (uint4*)ctx.scan_scratch.p, n, (uint32_t)K, G);
    H2B_CUDA(cudaGetLastError());
    return H2B_OK;
}

// ---- exclusive prefix product ----------------------------------------------------------------------------------------
static const uint32_t SCAN_K = 16;          // consecutive elements per thread
static const uint32_t SCAN_BASE = 1024;     // the bottom of the recursion: one CTA

__global__ void __launch_bounds__(128) fr_chunk_product_kernel(const uint4* __restrict__ in, size_t n, uint4* __restrict__ prod, uint32_t chunks) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks) return;
    const size_t lo = (size_t)t * SCAN_K;
    Fr run = fp_load<FR>(in + 2 * lo);
    for (uint32_t i = 1; i < SCAN_K && lo + i < n; ++i) run = fp_mul(run, fp_load<FR>(in + 2 * (lo + i)));
    fp_store<FR>(prod + 2 * (size_t)t, run);
}

// out[i] = offset[t] * in[lo] * ... * in[i-1] for the chunk t that holds i; in-place safe (each element is read before it is written)
__global__ void __launch_bounds__(128) fr_chunk_replay_kernel(const uint4* __restrict__ in, size_t n, const uint4* __restrict__ offset, uint4* __restrict__ out,
                                                            uint32_t chunks) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks) return;
    const size_t lo = (size_t)t * SCAN_K;
    Fr run = fp_load<FR>(offset + 2 * (size_t)t);
    for (uint32_t i = 0; i < SCAN_K && lo + i < n; ++i) {
        Fr x = fp_load<FR>(in + 2 * (lo + i));
        fp_store<FR>(out + 2 * (lo + i), run);
        run = fp_mul(run, x);
    }
}

// one CTA: exclusive prefix product of count <= SCAN_BASE elements (Hillis-Steele in shared memory), in place
__global__ void __launch_bounds__(1024) fr_scan_base_kernel(uint4* __restrict__ a, uint32_t count) {
    H2B_DYN_SMEM(uint4, sh);
    const uint32_t tid = threadIdx.x;
    Fr x = tid < count ? fp_load<FR>(a + 2 * (size_t)tid) : fp_one<FR>();
    for (uint32_t d = 1; d < count; d <<= 1) {
        fp_store<FR>(sh + 2 * tid, x);
        __syncthreads();
        if (tid >= d) x = fp_mul(x, fp_load<FR>(sh + 2 * (tid - d)));
        __syncthreads();
    }
    // inclusive -> exclusive
    fp_store<FR>(sh + 2 * tid, x);
    __syncthreads();
    if (tid < count) fp_store<FR>(a + 2 * (size_t)tid, tid ? fp_load<FR>(sh + 2 * (tid - 1)) : fp_one<FR>());
}

static int prefix_product_rec(DeviceCtx& ctx, const uint4* in, uint4* out, size_t n, uint4* scratch, cudaStream_t stream) {
    if (n <= SCAN_BASE) {
        if (in != out) H2B_CUDA(cudaMemcpyAsync(out, in, n * 32, cudaMemcpyDeviceToDevice, stream));
        uint32_t threads = 32;
        while (threads < n) threads <<= 1;
        H2B_LAUNCH(fr_scan_base_kernel, 1, threads, (size_t)threads * 32, stream, out, (uint32_t)n);
        H2B_CUDA(cudaGetLastError());
        return H2B_OK;
    }
    const uint32_t chunks = (uint32_t)((n + SCAN_K - 1) / SCAN_K);
    H2B_LAUNCH(fr_chunk_product_kernel, (chunks + 127) / 128, 128, 0, stream, in, n, scratch, chunks);
    H2B_TRY(prefix_product_rec(ctx, scratch, scratch, chunks, scratch + 2 * (size_t)chunks, stream));
    H2B_LAUNCH(fr_chunk_replay_kernel, (chunks + 127) / 128, 128, 0, stream, in, n, (const uint4*)scratch, out, chunks);
    H2B_CUDA(cudaGetLastError());
    return H2B_OK;
}

int fr_prefix_product_run(DeviceCtx& ctx, const void* d_in, void* d_out, size_t n, cudaStream_t stream) {
    if (n == 0) return H2B_OK;
    if (!d_in || !d_out) { set_error("prefix_product: null pointer"); return H2B_ERR_BAD_ARGUMENT; }
    if (n > ((size_t)1 << 31)) { set_error("prefix_product: at most 2^31 elements"); return H2B_ERR_BAD_ARGUMENT; }
    // chunk products of all levels: n/16 + n/256 + ... < n/15 elements (+ slack)
    H2B_TRY(ctx.scan_scratch.reserve((n / 15 + 4 * SCAN_BASE) * 32));
    return prefix_product_rec(ctx, (const uint4*)d_in, (uint4*)d_out, n, (uint4*)ctx.scan_scratch.p, stream);
}

}  // namespace h2b
