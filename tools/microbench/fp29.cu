// Prototype + micro-benchmark of the unsaturated-limb Montgomery multiplication (9 x 29-bit limbs, R' = 2^261):
// every partial product is a carry-free IMAD.WIDE.U32 into a 64-bit column accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp29 fp29.cu && ./fp29
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define NL9 9
#define MASK29 0x1fffffffu
// BN254 Fq modulus in 29-bit limbs and -p^-1 mod 2^29
__device__ __constant__ uint32_t dummy_c;
struct P29 {
    // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    static __host__ __device__ constexpr uint32_t P(int i) {
        constexpr uint32_t t[9] = {0x187cfd47, 0x10460b6, 0x1c72a34f, 0x2d522d0, 0x1585d978, 0x2db40c0, 0xa6e141, 0xe5c2634, 0x30644e};
        return t[i];
    }
    static constexpr uint32_t INV = 0x1a7ef10d ^ 0;   // filled by host check below
};

template <int MODE>
__device__ __forceinline__ void mul29(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9], uint32_t inv) {
    uint32_t m[9];
    u64 acc = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        u64 acc2 = 0;
#pragma unroll
        for (int i = 0; i <= k; ++i) {
            if (MODE == 1 && (i & 1)) acc2 += (u64)a[i] * b[k - i]; else acc += (u64)a[i] * b[k - i];
        }
#pragma unroll
        for (int i = 0; i < k; ++i) {
            if (MODE == 1 && (i & 1)) acc2 += (u64)m[i] * P29::P(k - i); else acc += (u64)m[i] * P29::P(k - i);
        }
        if (MODE == 1) acc += acc2;
        m[k] = ((uint32_t)acc * inv) & MASK29;
        acc += (u64)m[k] * P29::P(0);
        acc >>= 29;
    }
#pragma unroll
    for (int k = 9; k < 17; ++k) {
        u64 acc2 = 0;
#pragma unroll
        for (int i = k - 8; i < 9; ++i) {
            if (MODE == 1 && (i & 1)) { acc2 += (u64)a[i] * b[k - i]; acc2 += (u64)m[i] * P29::P(k - i); }
            else { acc += (u64)a[i] * b[k - i]; acc += (u64)m[i] * P29::P(k - i); }
        }
        if (MODE == 1) acc += acc2;
        r[k - 9] = (uint32_t)acc & MASK29;
        acc >>= 29;
    }
    r[8] = (uint32_t)acc;
}

// operand scanning: row i adds a_j*b_i to 9 independent column accumulators (b_i and then m_i stay in the
// operand-reuse slot for 9 instructions, the multiplicand p_j is an immediate)
__device__ __forceinline__ void mul29_os(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9], uint32_t inv) {
    u64 acc[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i + j] += (u64)a[j] * b[i];
        uint32_t m = ((uint32_t)acc[i] * inv) & MASK29;
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i + j] += (u64)m * P29::P(j);
        acc[i + 1] += acc[i] >> 29;
    }
#pragma unroll
    for (int k = 9; k < 17; ++k) {
        r[k - 9] = (uint32_t)acc[k] & MASK29;
        acc[k + 1] += acc[k] >> 29;
    }
    r[8] = (uint32_t)acc[17];
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* sink, const uint32_t* in, int iters, uint32_t inv) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[9], b[9], c[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { a[i] = in[i] ^ (t & 0xff); b[i] = in[9 + i]; }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 2) mul29_os(c, a, b, inv); else mul29<MODE>(c, a, b, inv);
#pragma unroll
        for (int i = 0; i < 9; ++i) { a[i] = b[i]; b[i] = c[i]; }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) x ^= b[i];
    sink[t] = x;
}

__global__ void one(uint32_t* out, const uint32_t* in, uint32_t inv) {
    uint32_t a[9], b[9], c[9];
    for (int i = 0; i < 9; ++i) { a[i] = in[i]; b[i] = in[9 + i]; }
    mul29<0>(c, a, b, inv);
    for (int i = 0; i < 9; ++i) out[i] = c[i];
    mul29_os(c, a, b, inv);
    for (int i = 0; i < 9; ++i) out[9 + i] = c[i];
}

int main() {
    // host-side check values are produced by tools (python) -- here we only time and print one product
    uint32_t h_in[18] = {0x1234567, 0x89abcd, 0x1fffffff, 0x1, 0x2, 0x3, 0x4, 0x5, 0x1234,   0x1fedcba, 0x765432, 0x10, 0x20, 0x30, 0x40, 0x50, 0x60, 0x4321};
    uint32_t inv = 0;
    {   // -p^-1 mod 2^29 by Newton iteration on the low limb
        uint32_t p0 = P29::P(0), x = 1;
        for (int i = 0; i < 6; ++i) x = x * (2 - p0 * x);
        inv = (0u - x) & MASK29;
    }
    uint32_t *d_in, *d_out, *sink;
    cudaMalloc(&d_in, sizeof(h_in)); cudaMalloc(&d_out, 18 * 4);
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    one<<<1, 1>>>(d_out, d_in, inv);
    uint32_t h_out[18];
    cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
    printf("{\"inv\": \"0x%x\", \"r0\": [", inv);
    for (int i = 0; i < 9; ++i) printf("%u%s", h_out[i], i < 8 ? ", " : "], \"r1\": [");
    for (int i = 0; i < 9; ++i) printf("%u%s", h_out[9 + i], i < 8 ? ", " : "]}\n");
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    for (int mode = 0; mode < 3; ++mode) {
        for (int bps : {2, 4, 8}) {
            int blocks = sms * bps, threads = 256, iters = 2048;
            cudaMalloc(&sink, (size_t)blocks * threads * 4);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            if (mode == 0) bench<0><<<blocks, threads>>>(sink, d_in, 64, inv); else if (mode == 1) bench<1><<<blocks, threads>>>(sink, d_in, 64, inv); else bench<2><<<blocks, threads>>>(sink, d_in, 64, inv);
            cudaEventRecord(e0);
            if (mode == 0) bench<0><<<blocks, threads>>>(sink, d_in, iters, inv); else if (mode == 1) bench<1><<<blocks, threads>>>(sink, d_in, iters, inv); else bench<2><<<blocks, threads>>>(sink, d_in, iters, inv);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double muls = (double)iters * blocks * threads;
            printf("{\"variant\": \"fp29_mul mode %d\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"gmul_per_s\": %.2f, \"clk_per_warp_mul_per_smsp\": %.1f}\n", mode, bps, ms,
                   muls / ms / 1e6, 32.0 / (muls / (ms * 1e-3) / (sms * 4.0) / 1.965e9));
            cudaFree(sink);
        }
    }
    return 0;
}
