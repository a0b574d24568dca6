// What limits IMAD.WIDE.U32 throughput? (dependent chains, operand variety, immediates, carry-out)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 2048
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* sink, const uint32_t* in, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[i] + t; b[i] = in[8 + i] ^ t; }
    u64 c0 = t, c1 = t + 1, c2 = t + 2, c3 = t + 3, c4 = t + 4, c5 = t + 5, c6 = t + 6, c7 = t + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (V == 0) {          // 8 independent accumulators, distinct register multiplicands
                c0 += (u64)a[0] * b[u]; c1 += (u64)a[1] * b[(u + 1) & 7]; c2 += (u64)a[2] * b[(u + 2) & 7]; c3 += (u64)a[3] * b[(u + 3) & 7];
                c4 += (u64)a[4] * b[(u + 4) & 7]; c5 += (u64)a[5] * b[(u + 5) & 7]; c6 += (u64)a[6] * b[(u + 6) & 7]; c7 += (u64)a[7] * b[(u + 7) & 7];
            } else if (V == 1) {   // ONE accumulator: fully dependent chain of 8
                c0 += (u64)a[0] * b[u]; c0 += (u64)a[1] * b[(u + 1) & 7]; c0 += (u64)a[2] * b[(u + 2) & 7]; c0 += (u64)a[3] * b[(u + 3) & 7];
                c0 += (u64)a[4] * b[(u + 4) & 7]; c0 += (u64)a[5] * b[(u + 5) & 7]; c0 += (u64)a[6] * b[(u + 6) & 7]; c0 += (u64)a[7] * b[(u + 7) & 7];
            } else if (V == 2) {   // two accumulators, alternating
                c0 += (u64)a[0] * b[u]; c1 += (u64)a[1] * b[(u + 1) & 7]; c0 += (u64)a[2] * b[(u + 2) & 7]; c1 += (u64)a[3] * b[(u + 3) & 7];
                c0 += (u64)a[4] * b[(u + 4) & 7]; c1 += (u64)a[5] * b[(u + 5) & 7]; c0 += (u64)a[6] * b[(u + 6) & 7]; c1 += (u64)a[7] * b[(u + 7) & 7];
            } else if (V == 3) {   // 8 independent accumulators, one multiplicand is an immediate
                c0 += (u64)a[0] * 0x187cfd47u; c1 += (u64)a[1] * 0x10460b6u; c2 += (u64)a[2] * 0x1c72a34fu; c3 += (u64)a[3] * 0x2d522d0u;
                c4 += (u64)a[4] * 0x1585d978u; c5 += (u64)a[5] * 0x2db40c0u; c6 += (u64)a[6] * 0xa6e141u; c7 += (u64)a[7] * 0xe5c2634u;
                a[u] += (uint32_t)c0;
            } else if (V == 4) {   // 4 accumulators
                c0 += (u64)a[0] * b[u]; c1 += (u64)a[1] * b[(u + 1) & 7]; c2 += (u64)a[2] * b[(u + 2) & 7]; c3 += (u64)a[3] * b[(u + 3) & 7];
                c0 += (u64)a[4] * b[(u + 4) & 7]; c1 += (u64)a[5] * b[(u + 5) & 7]; c2 += (u64)a[6] * b[(u + 6) & 7]; c3 += (u64)a[7] * b[(u + 7) & 7];
            } else if (V == 5) {   // 8 independent products WITHOUT accumulate (mul.wide), xor-folded on the ALU
                u64 p0 = (u64)a[0] * b[u], p1 = (u64)a[1] * b[(u + 1) & 7], p2 = (u64)a[2] * b[(u + 2) & 7], p3 = (u64)a[3] * b[(u + 3) & 7];
                u64 p4 = (u64)a[4] * b[(u + 4) & 7], p5 = (u64)a[5] * b[(u + 5) & 7], p6 = (u64)a[6] * b[(u + 6) & 7], p7 = (u64)a[7] * b[(u + 7) & 7];
                c0 ^= p0; c1 ^= p1; c2 ^= p2; c3 ^= p3; c4 ^= p4; c5 ^= p5; c6 ^= p6; c7 ^= p7;
            }
        }
    }
    u64 x = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
    sink[t] = (uint32_t)x ^ (uint32_t)(x >> 32) ^ a[0] ^ a[3];
}
template <int V> void run(const char* name, int sms, const uint32_t* d_in) {
    for (int bps : {2, 8}) {
        uint32_t* sink; int blocks = sms * bps, threads = 256;
        cudaMalloc(&sink, (size_t)blocks * threads * 4);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<V><<<blocks, threads>>>(sink, d_in, 32);
        cudaEventRecord(e0);
        k<V><<<blocks, threads>>>(sink, d_in, ITERS);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = 64.0 * ITERS * blocks * threads;
        double lanes = ops / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
        printf("{\"variant\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"clk_per_warp_instr\": %.2f}\n", name, bps, ms, 32.0 / lanes);
        cudaFree(sink);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t h[16]; for (int i = 0; i < 16; ++i) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t* d_in; cudaMalloc(&d_in, 64); cudaMemcpy(d_in, h, 64, cudaMemcpyHostToDevice);
    run<0>("8 acc, distinct regs", p.multiProcessorCount, d_in);
    run<1>("1 acc (dependent chain)", p.multiProcessorCount, d_in);
    run<2>("2 acc", p.multiProcessorCount, d_in);
    run<4>("4 acc", p.multiProcessorCount, d_in);
    run<3>("8 acc, immediate multiplicand", p.multiProcessorCount, d_in);
    run<5>("8 mul.wide no acc + xor", p.multiProcessorCount, d_in);
    return 0;
}
