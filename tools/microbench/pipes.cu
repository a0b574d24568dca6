// Integer-pipe micro-benchmarks for sm_100a: issue rates of the instruction variants a multi-limb
// Montgomery multiplication can be built from.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o pipes pipes.cu && ./pipes     (prints one JSON line per variant)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* sink, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = t | 1, b = t * 3 + 7;
    uint32_t r0 = t, r1 = t + 1, r2 = t + 2, r3 = t + 3, r4 = t + 4, r5 = t + 5, r6 = t + 6, r7 = t + 7;
    uint32_t s0 = t, s1 = t + 1, s2 = t + 2, s3 = t + 3, s4 = t + 4, s5 = t + 5, s6 = t + 6, s7 = t + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (V == 0) {        // IMAD (32-bit): 8 independent
                asm volatile("mad.lo.u32 %0,%0,%8,%9; mad.lo.u32 %1,%1,%8,%9; mad.lo.u32 %2,%2,%8,%9; mad.lo.u32 %3,%3,%8,%9;"
                             "mad.lo.u32 %4,%4,%8,%9; mad.lo.u32 %5,%5,%8,%9; mad.lo.u32 %6,%6,%8,%9; mad.lo.u32 %7,%7,%8,%9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 1) { // IMAD.WIDE.U32 without accumulate (mul.wide): 4 independent 64-bit results, inputs vary
                asm volatile("{ .reg .u64 t0,t1,t2,t3;\n"
                             "mul.wide.u32 t0,%0,%8; mul.wide.u32 t1,%2,%8; mul.wide.u32 t2,%4,%8; mul.wide.u32 t3,%6,%8;\n"
                             "mov.b64 {%0,%1},t0; mov.b64 {%2,%3},t1; mov.b64 {%4,%5},t2; mov.b64 {%6,%7},t3; }"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 2) { // IMAD.WIDE.U32 with 64-bit accumulate, no carry: 4 independent
                asm volatile("{ .reg .u64 t0,t1,t2,t3;\n"
                             "mov.b64 t0,{%0,%1}; mov.b64 t1,{%2,%3}; mov.b64 t2,{%4,%5}; mov.b64 t3,{%6,%7};\n"
                             "mad.wide.u32 t0,%8,%9,t0; mad.wide.u32 t1,%8,%9,t1; mad.wide.u32 t2,%8,%9,t2; mad.wide.u32 t3,%8,%9,t3;\n"
                             "mov.b64 {%0,%1},t0; mov.b64 {%2,%3},t1; mov.b64 {%4,%5},t2; mov.b64 {%6,%7},t3; }"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 3) { // carry chain of 4 wide MADs (1 plain + 3 .X), as in the Montgomery rows; two chains
                asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                             "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
                asm volatile("mad.lo.cc.u32 %0,%8,%9,%0; madc.hi.cc.u32 %1,%8,%9,%1; madc.lo.cc.u32 %2,%8,%9,%2; madc.hi.cc.u32 %3,%8,%9,%3;"
                             "madc.lo.cc.u32 %4,%8,%9,%4; madc.hi.cc.u32 %5,%8,%9,%5; madc.lo.cc.u32 %6,%8,%9,%6; madc.hi.u32 %7,%8,%9,%7;"
                             : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7) : "r"(a), "r"(b));
            } else if (V == 4) { // IADD3 (3-input add), 8 independent
                asm volatile("add.u32 %0,%0,%8; add.u32 %1,%1,%9; add.u32 %2,%2,%8; add.u32 %3,%3,%9; add.u32 %4,%4,%8; add.u32 %5,%5,%9; add.u32 %6,%6,%8; add.u32 %7,%7,%9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 5) { // add.cc chain of 8 (IADD3 + IADD3.X with predicate carries)
                asm volatile("add.cc.u32 %0,%0,%8; addc.cc.u32 %1,%1,%9; addc.cc.u32 %2,%2,%8; addc.cc.u32 %3,%3,%9; addc.cc.u32 %4,%4,%8; addc.cc.u32 %5,%5,%9; addc.cc.u32 %6,%6,%8; addc.u32 %7,%7,%9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 6) { // IMAD.HI.U32, 8 independent
                asm volatile("mad.hi.u32 %0,%0,%8,%9; mad.hi.u32 %1,%1,%8,%9; mad.hi.u32 %2,%2,%8,%9; mad.hi.u32 %3,%3,%8,%9;"
                             "mad.hi.u32 %4,%4,%8,%9; mad.hi.u32 %5,%5,%8,%9; mad.hi.u32 %6,%6,%8,%9; mad.hi.u32 %7,%7,%8,%9;"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
            } else if (V == 7) { // 4 mul.wide (FMA pipe) + 8 independent adds (ALU pipe): do the pipes overlap?
                asm volatile("{ .reg .u64 t0,t1,t2,t3;\n"
                             "mul.wide.u32 t0,%0,%8; mul.wide.u32 t1,%2,%8; mul.wide.u32 t2,%4,%8; mul.wide.u32 t3,%6,%8;\n"
                             "mov.b64 {%0,%1},t0; mov.b64 {%2,%3},t1; mov.b64 {%4,%5},t2; mov.b64 {%6,%7},t3; }"
                             : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));
                asm volatile("add.u32 %0,%0,%8; add.u32 %1,%1,%9; add.u32 %2,%2,%8; add.u32 %3,%3,%9; add.u32 %4,%4,%8; add.u32 %5,%5,%9; add.u32 %6,%6,%8; add.u32 %7,%7,%9;"
                             : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7) : "r"(a), "r"(b));
            } else if (V == 8) { // DFMA, 8 independent
                double d0 = __longlong_as_double(((long long)r0 << 32) | r1), x = __longlong_as_double(0x3ff0000000000001ll + a), y = 1e-9;
                double d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3, d4 = d0 + 4, d5 = d0 + 5, d6 = d0 + 6, d7 = d0 + 7;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    asm volatile("fma.rz.f64 %0,%0,%8,%9; fma.rz.f64 %1,%1,%8,%9; fma.rz.f64 %2,%2,%8,%9; fma.rz.f64 %3,%3,%8,%9;"
                                 "fma.rz.f64 %4,%4,%8,%9; fma.rz.f64 %5,%5,%8,%9; fma.rz.f64 %6,%6,%8,%9; fma.rz.f64 %7,%7,%8,%9;"
                                 : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(x), "d"(y));
                }
                r0 ^= (uint32_t)__double_as_longlong(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
            }
        }
    }
    sink[t] = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7;
}

template <int V>
void run(const char* name, double ops_per_inner, int sms) {
    uint32_t* sink;
    int blocks = sms * 8, threads = 256;
    cudaMalloc(&sink, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<blocks, threads>>>(sink, 64);
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(sink, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = ops_per_inner * UNROLL * (double)ITERS * blocks * threads;
    double per_smsp_clk = ops / (ms * 1e-3) / (sms * 4.0) / 1.965e9;    // thread-ops per SMSP per clock at max clock
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"gops\": %.1f, \"lanes_per_clk_per_smsp\": %.2f, \"clk_per_warp_instr\": %.2f}\n", name, ms,
           ops / ms / 1e6, per_smsp_clk, 32.0 / per_smsp_clk);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    run<0>("imad32", 8, sms);
    run<1>("imad_wide_noacc", 4, sms);
    run<2>("imad_wide_acc", 4, sms);
    run<3>("imad_wide_X_chain(8 wide per 2 chains)", 8, sms);
    run<4>("iadd", 8, sms);
    run<5>("iadd_carry_chain", 8, sms);
    run<6>("imad_hi", 8, sms);
    run<7>("4 mul.wide + 8 add (pipe overlap; count=12)", 12, sms);
    run<8>("dfma (x4 per inner)", 32, sms);
    return 0;
}
