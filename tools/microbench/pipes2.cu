// Round-2 pipe micro-benchmarks for sm_100a (B200): what does a 32x32->64 multiply-add really cost in the operand patterns a
// multi-limb Montgomery multiplication uses, and can the FP64 pipe (DFMA, idle in every kernel of this library) carry part
// of the products?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
// One JSON line per variant: clk per warp-instruction per SM sub-partition at the clock measured with clock64().
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define ITERS 2048

// ---- B: row-wise schoolbook, 8x8 wide MADs into 16 independent 64-bit column accumulators (no carries at all) ----
__device__ __forceinline__ void rows8x8(u64 (&acc)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(a[j]), "r"(b[i]));
}
// ---- C: column-wise (product scanning): every product of a column goes into ONE accumulator (dependent chain) ----
__device__ __forceinline__ void cols8x8(u64 (&acc)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
    for (int k = 0; k < 15; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (k - i >= 0 && k - i < 8) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[i]), "r"(b[k - i]));
}
// ---- D: mul.wide (no accumulate) into 16 results, xor-folded by the ALU ----
__device__ __forceinline__ void mulwide8x8(u64 (&acc)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u64 t;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[j]), "r"(b[i]));
            acc[i + j] ^= t;
        }
}
// ---- E: the carry-chain form of the shipped multiplier: row i = 8 products added with mad.lo.cc / madc.hi.cc (IMAD.WIDE.X) ----
__device__ __forceinline__ void rows8x8_carry(uint32_t (&r)[17], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // even products
        asm volatile("mad.lo.cc.u32 %0,%8,%12,%0; madc.hi.cc.u32 %1,%8,%12,%1; madc.lo.cc.u32 %2,%9,%12,%2; madc.hi.cc.u32 %3,%9,%12,%3;"
                     "madc.lo.cc.u32 %4,%10,%12,%4; madc.hi.cc.u32 %5,%10,%12,%5; madc.lo.cc.u32 %6,%11,%12,%6; madc.hi.cc.u32 %7,%11,%12,%7;"
                     : "+r"(r[i]), "+r"(r[i + 1]), "+r"(r[i + 2]), "+r"(r[i + 3]), "+r"(r[i + 4]), "+r"(r[i + 5]), "+r"(r[i + 6]), "+r"(r[i + 7])
                     : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b[i]));
        asm volatile("addc.u32 %0,%0,0;" : "+r"(r[i + 8]));
    }
}

// ---- F: unsaturated 9 x 29-bit Montgomery multiplication (BN254 Fq), products fused with their 64-bit accumulation by ptxas ----
#define MASK29 0x1fffffffu
__device__ __forceinline__ constexpr uint32_t PL(int i) {
    constexpr uint32_t t[9] = {0x187cfd47, 0x10460b6, 0x1c72a34f, 0x2d522d0, 0x1585d978, 0x2db40c0, 0xa6e141, 0xe5c2634, 0x30644e};
    return t[i];
}
__device__ __forceinline__ u64 mulw(uint32_t a, uint32_t b) { u64 t; asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b)); return t; }
__device__ __forceinline__ void mul29(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9], uint32_t inv) {
    u64 acc[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i + j] += mulw(a[j], b[i]);
        uint32_t m = ((uint32_t)acc[i] * inv) & MASK29;
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i + j] += mulw(m, PL(j));
        acc[i + 1] += acc[i] >> 29;
    }
#pragma unroll
    for (int k = 9; k < 17; ++k) {
        r[k - 9] = (uint32_t)acc[k] & MASK29;
        acc[k + 1] += acc[k] >> 29;
    }
    r[8] = (uint32_t)acc[17];
}
// the same with every product forced to the accumulate-free form: mad.wide with a 64-bit addend written as PTX (ptxas emits
// IMAD.WIDE Rd, Ra, Rb, RZ + IADD3 / IADD3.X on the ALU pipe for it)
__device__ __forceinline__ void mul29_split(uint32_t (&r)[9], const uint32_t (&a)[9], const uint32_t (&b)[9], uint32_t inv) {
    u64 acc[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(a[j]), "r"(b[i]));
        uint32_t m = ((uint32_t)acc[i] * inv) & MASK29;
#pragma unroll
        for (int j = 0; j < 9; ++j) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(m), "r"(PL(j)));
        acc[i + 1] += acc[i] >> 29;
    }
#pragma unroll
    for (int k = 9; k < 17; ++k) {
        r[k - 9] = (uint32_t)acc[k] & MASK29;
        acc[k + 1] += acc[k] >> 29;
    }
    r[8] = (uint32_t)acc[17];
}

// ---- G: saturated 32-bit limbs, products as pure mul.wide, accumulation by add.cc chains on the ALU pipe (4 products: 4 wide + 9 adds) ----
__device__ __forceinline__ void row_split4(uint32_t* r, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t bi) {
    asm volatile("{ .reg .u64 t0,t1,t2,t3; .reg .u32 l0,h0,l1,h1,l2,h2,l3,h3;\n"
                 "mul.wide.u32 t0,%9,%13; mul.wide.u32 t1,%10,%13; mul.wide.u32 t2,%11,%13; mul.wide.u32 t3,%12,%13;\n"
                 "mov.b64 {l0,h0},t0; mov.b64 {l1,h1},t1; mov.b64 {l2,h2},t2; mov.b64 {l3,h3},t3;\n"
                 "add.cc.u32 %0,%0,l0; addc.cc.u32 %1,%1,h0; addc.cc.u32 %2,%2,l1; addc.cc.u32 %3,%3,h1;\n"
                 "addc.cc.u32 %4,%4,l2; addc.cc.u32 %5,%5,h2; addc.cc.u32 %6,%6,l3; addc.cc.u32 %7,%7,h3; addc.u32 %8,%8,0; }"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bi));
}
__device__ __forceinline__ void row_carry4(uint32_t* r, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t bi) {
    asm volatile("mad.lo.cc.u32 %0,%9,%13,%0; madc.hi.cc.u32 %1,%9,%13,%1; madc.lo.cc.u32 %2,%10,%13,%2; madc.hi.cc.u32 %3,%10,%13,%3;"
                 "madc.lo.cc.u32 %4,%11,%13,%4; madc.hi.cc.u32 %5,%11,%13,%5; madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.cc.u32 %7,%12,%13,%7; addc.u32 %8,%8,0;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bi));
}

template <int V>
__global__ void __launch_bounds__(256) kern(uint32_t* sink, int iters, u64* clk) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = t * 2654435761u + i * 40503u + 1; b[i] = t * 40503u + i * 2654435761u + 3; }
    u64 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = t + i;
    uint32_t r[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) r[i] = t + i;
    double d[8], x = 1.0000000001 + t * 1e-12, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 1.0 + i + t;
    u64 c0 = clock64(), g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int it = 0; it < iters; ++it) {
        if (V == 0) rows8x8(acc, a, b);
        if (V == 1) cols8x8(acc, a, b);
        if (V == 2) mulwide8x8(acc, a, b);
        if (V == 3) rows8x8_carry(r, a, b);
        if (V == 4) {   // 64 DFMA, 8 independent chains, distinct multiplicands
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(x), "d"(y));
        }
        if (V == 5) {   // 64 DFMA interleaved with 64 wide MADs (no carry): do the FP64 and integer pipes overlap?
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(a[j]), "r"(b[i]));
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(x), "d"(y));
                }
        }
        if (V == 6) {   // 64 DFMA interleaved with the carry-chain rows (64 IMAD.WIDE.X-class)
            rows8x8_carry(r, a, b);
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(x), "d"(y));
        }
        if (V == 7) {   // 64 wide MADs (no carry) + 128 IADD3-class ALU ops: does ALU work hide behind the multiplier?
            rows8x8(acc, a, b);
#pragma unroll
            for (int q = 0; q < 16; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[j]) : "r"(a[(j + q) & 7]));
        }
        if (V == 8) {   // 64 wide MADs + 64 funnel shifts (the carry extraction of unsaturated limbs)
            rows8x8(acc, a, b);
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(r[j]) : "r"(r[j + 1]));
        }
        if (V == 9) {   // 64 DFMA + 64 DADD (the hi/lo split of the FP64 multiplier needs one subtraction per product)
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(x), "d"(y));
                    asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[(j + 4) & 7]) : "d"(y));
                }
        }
        if (V == 10 || V == 11) {   // 4 dependent 9 x 29-bit Montgomery multiplications (162 products + 9 IMAD each)
            uint32_t a9[9], b9[9], c9[9];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a9[i] = a[i] & MASK29; b9[i] = b[i] & MASK29; }
            a9[8] = r[0] & 0xfffff; b9[8] = r[1] & 0xfffff;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                if (V == 10) mul29(c9, a9, b9, 0x4866389u); else mul29_split(c9, a9, b9, 0x4866389u);
#pragma unroll
                for (int i = 0; i < 9; ++i) { a9[i] = b9[i]; b9[i] = c9[i]; }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] ^= a9[i]; b[i] ^= b9[i]; }
            r[0] ^= a9[8]; r[1] ^= b9[8];
        }
        if (V == 12) {   // 256 three-input adds, 8 independent chains, distinct operands: the ALU pipe alone
#pragma unroll
            for (int q = 0; q < 32; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = r[j] + a[(j + q) & 7] + b[(j + 3 * q + 1) & 7];
        }
        if (V == 13) {   // 64 pure products, each folded by ONE three-input logic op
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    u64 tt;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(tt) : "r"(a[j]), "r"(b[i]));
                    r[j] = r[j] ^ (uint32_t)tt ^ (uint32_t)(tt >> 32);
                }
        }
        if (V == 14 || V == 15 || V == 16) {   // 8 x 8 schoolbook on saturated limbs: 14 all split, 15 even products carry-chain + odd split, 16 all carry-chain
            uint32_t w[18];
#pragma unroll
            for (int i = 0; i < 17; ++i) w[i] = r[i];
            w[17] = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (V == 14) row_split4(w + i, a[0], a[2], a[4], a[6], b[i]); else row_carry4(w + i, a[0], a[2], a[4], a[6], b[i]);
                if (V == 16) row_carry4(w + i + 1, a[1], a[3], a[5], a[7], b[i]); else row_split4(w + i + 1, a[1], a[3], a[5], a[7], b[i]);
            }
#pragma unroll
            for (int i = 0; i < 17; ++i) r[i] = w[i];
            r[0] ^= w[17];
        }
        // keep the operands moving so that nothing is loop-invariant
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += (uint32_t)acc[i] | r[i + 1]; b[i] ^= (uint32_t)(acc[i + 8] >> 32) + r[i + 9]; }
    }
    u64 c1 = clock64(), g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32) ^ r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= (uint32_t)__double_as_longlong(d[i]);
    sink[t] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = c1 - c0; clk[1] = g1 - g0; }
}

template <int V>
void run(const char* name, double instr_per_iter, int sms, int blocks_per_sm) {
    uint32_t* sink; u64* clk;
    int blocks = sms * blocks_per_sm, threads = 256;
    cudaMalloc(&sink, (size_t)blocks * threads * 4);
    cudaMalloc(&clk, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<V><<<blocks, threads>>>(sink, 64, clk);
    cudaEventRecord(e0);
    kern<V><<<blocks, threads>>>(sink, ITERS, clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    u64 h_clk[2] = {0, 0}; cudaMemcpy(h_clk, clk, 16, cudaMemcpyDeviceToHost);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<V>, threads, 0);
    const double clock_hz = (double)h_clk[0] / ((double)h_clk[1] * 1e-9);          // SM clock while this kernel ran
    const double warp_instr = (double)ITERS * instr_per_iter * blocks * (threads / 32);
    const double clk_per_instr = ms * 1e-3 * clock_hz * sms * 4 / warp_instr;      // issue clocks per warp instruction per SM sub-partition
    printf("{\"variant\": \"%s\", \"blocks_per_sm\": %d, \"resident_blocks_per_sm\": %d, \"ms\": %.3f, \"sm_clock_mhz\": %.0f, \"clk_per_warp_instr_per_smsp\": %.3f}\n", name,
           blocks_per_sm, occ, ms, clock_hz / 1e6, clk_per_instr);
    cudaFree(sink); cudaFree(clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d}\n", p.name, sms);
    for (int bps : {16}) {
        run<0>("64 wide MAD, row-wise, 16 independent 64-bit accumulators, distinct operands", 64, sms, bps);
        run<1>("64 wide MAD, column-wise (dependent chain per column)", 64, sms, bps);
        run<2>("64 mul.wide + 64 xor", 64, sms, bps);
        run<3>("64 products as mad.lo.cc/madc.hi.cc chains (IMAD.WIDE.X) [counted as 32 wide]", 32, sms, bps);
        run<4>("64 DFMA (8 chains)", 64, sms, bps);
        run<5>("64 wide MAD + 64 DFMA interleaved [counted 64]", 64, sms, bps);
        run<6>("32 wide.X + 64 DFMA [counted 64 DFMA]", 64, sms, bps);
        run<7>("64 wide MAD + 128 IADD [counted 64]", 64, sms, bps);
        run<8>("64 wide MAD + 64 SHF [counted 64]", 64, sms, bps);
        run<9>("64 DFMA + 64 DADD [counted 128]", 128, sms, bps);
        run<10>("4 Montgomery mul 9x29 (ptxas: accumulate form) [counted per mul]", 4, sms, bps);
        run<11>("4 Montgomery mul 9x29 (products RZ + ALU adds) [counted per mul]", 4, sms, bps);
        run<12>("256 IADD3", 256, sms, bps);
        run<13>("64 mul.wide + 64 LOP3", 64, sms, bps);
        run<14>("8x8 saturated, all products mul.wide + add.cc chains [counted 64 products]", 64, sms, bps);
        run<15>("8x8 saturated, even products IMAD.WIDE.X chains, odd products mul.wide + add.cc chains [64]", 64, sms, bps);
        run<16>("8x8 saturated, all IMAD.WIDE.X chains [64]", 64, sms, bps);
    }
    return 0;
}
