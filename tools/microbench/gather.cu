// What is the random-access rate of HBM3e for the MSM's point gathers?
// A table of 64-byte records (13 GiB, like the 13 window tables of a 2^24-point SRS) is read at hashed positions:
//   bytes per gather 32 / 64, U independent gathers in flight per thread, CTAs per SM, and the L2 fetch granularity
//   (cudaLimitMaxL2FetchGranularity 32 / 64 / 128).  Also random 64-byte WRITES (a sort that moves points, not indices).
// Output: one JSON line per configuration with G accesses/s.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int BYTES, int U>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ table, uint32_t nrec_mask, uint32_t iters, uint32_t* __restrict__ sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint4 v[U][BYTES / 16];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = hash32(t * 2654435761u + it * U + u) & nrec_mask;
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) v[u][k] = __ldg(table + 4 * (size_t)r + k);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) acc ^= v[u][k].x ^ v[u][k].y ^ v[u][k].z ^ v[u][k].w;
    }
    sink[t] = acc;
}

template <int BYTES>
__global__ void __launch_bounds__(256) scatter_kernel(uint4* __restrict__ table, uint32_t nrec_mask, uint32_t iters, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t r = hash32(t * 2654435761u + it + seed) & nrec_mask;
        const uint4 v = make_uint4(t, it, r, seed);
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) table[4 * (size_t)r + k] = v;
    }
}

template <int BYTES, int U>
static void run_gather(const uint4* table, uint32_t mask, int sms, int bps, const char* note) {
    const int blocks = sms * bps, threads = 256;
    uint32_t* sink;
    cudaMalloc(&sink, (size_t)blocks * threads * 4);
    const uint32_t iters = 2048 / U;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<BYTES, U><<<blocks, threads>>>(table, mask, 8, sink);
    cudaEventRecord(e0);
    gather_kernel<BYTES, U><<<blocks, threads>>>(table, mask, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * threads * iters * U;
    printf("{\"op\": \"gather\", \"bytes\": %d, \"in_flight_per_thread\": %d, \"ctas_per_sm\": %d, \"note\": \"%s\", \"ms\": %.3f, \"g_per_s\": %.2f, \"useful_tb_s\": %.3f}\n", BYTES, U, bps, note, ms,
           n / ms / 1e6, n * BYTES / ms / 1e9);
    fflush(stdout);
    cudaFree(sink);
}

template <int BYTES>
static void run_scatter(uint4* table, uint32_t mask, int sms, int bps, const char* note) {
    const int blocks = sms * bps, threads = 256;
    const uint32_t iters = 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    scatter_kernel<BYTES><<<blocks, threads>>>(table, mask, 8, 1);
    cudaEventRecord(e0);
    scatter_kernel<BYTES><<<blocks, threads>>>(table, mask, iters, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * threads * iters;
    printf("{\"op\": \"scatter\", \"bytes\": %d, \"ctas_per_sm\": %d, \"note\": \"%s\", \"ms\": %.3f, \"g_per_s\": %.2f, \"useful_tb_s\": %.3f}\n", BYTES, bps, note, ms, n / ms / 1e6, n * BYTES / ms / 1e9);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const uint32_t nrec = 1u << 27;      // 2^27 x 64 B = 8 GiB
    uint4* table;
    if (cudaMalloc(&table, (size_t)nrec * 64) != cudaSuccess) { printf("{\"error\": \"cudaMalloc\"}\n"); return 1; }
    cudaMemset(table, 1, (size_t)nrec * 64);
    for (int gran : {0, 32, 64, 128}) {
        char note[64];
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            snprintf(note, sizeof(note), "L2 fetch granularity %d (set rc %d, now %zu)", gran, (int)e, got);
        } else {
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            snprintf(note, sizeof(note), "default granularity %zu", got);
        }
        run_gather<64, 1>(table, nrec - 1, sms, 8, note);
        run_gather<64, 2>(table, nrec - 1, sms, 8, note);
        run_gather<64, 4>(table, nrec - 1, sms, 8, note);
        run_gather<64, 8>(table, nrec - 1, sms, 8, note);
        run_gather<64, 4>(table, nrec - 1, sms, 2, note);
        run_gather<64, 4>(table, nrec - 1, sms, 4, note);
        run_gather<32, 4>(table, nrec - 1, sms, 8, note);
        run_gather<32, 8>(table, nrec - 1, sms, 8, note);
        run_scatter<64>(table, nrec - 1, sms, 8, note);
        run_scatter<32>(table, nrec - 1, sms, 8, note);
    }
    cudaFree(table);
    return 0;
}
