// BN254 G1 multi-scalar multiplication for sm_100a -- the device side of `best_multiexp`
// ([UP] halo2_proofs/src/arithmetic.rs::best_multiexp / multiexp_serial @ v2023_02_02; SURVEY.md
// section 8 rows a1/a2; reached through ParamsKZG::commit{,_lagrange} from src/scaffold.rs:132,135,
// 191-199,207-214,223,284,287,322-346,354 and examples/standard_plonk.rs:33,34,41-49,57).
//
// Contract (identical to the reference): scalars n x 32 B (Fr, Montgomery), bases n x 64 B
// (G1Affine x|y Montgomery, (0,0) = identity) -> sum_i s_i * P_i as a Jacobian triple whose
// affine value is the unique group element the reference computes.
//
// Algorithm (all on the device, no host arithmetic):
//   1. decompose: leave Montgomery form, map s > (r-1)/2 to (r - s, -P) so that "small negative"
//      witness values stay small, recode into W signed c-bit digits |d| <= 2^(c-1), histogram the
//      (window, |d|) buckets with global reductions;
//   2. exclusive scan of the histogram; counting-sort scatter of (point index | sign) by bucket;
//   3. plan: buckets longer than a cap L are split into tasks so that skewed inputs (50 % of a
//      witness column equal to 1 ...) cannot serialise on one thread;
//   4. accumulate: one thread per task walks its slice of the sorted list and adds the gathered
//      affine bases into an XYZZ accumulator (8M + 2S per point, next point prefetched);
//   5. combine the partial sums of split buckets (one warp per split bucket, shuffle tree);
//   6. bucket reduction sum_b b * B_b as a hierarchy of chunked running sums (m buckets per thread
//      per level), then Horner over the windows and conversion to Jacobian.
#include "common.h"
#include "ec.cuh"

namespace h2b {

static const uint32_t SIGN_BIT = 0x80000000u;

struct MsmPlan {
    uint32_t n;          // points in this (sub-)MSM, <= 2^26
    uint32_t c;          // window bits
    uint32_t W;          // windows
    uint32_t Nb;         // buckets per window = 2^(c-1), ids 1..Nb
    uint32_t B;          // W * Nb
    uint32_t L;          // task length cap
    uint32_t max_overflow;
};

// (r - 1) / 2 and r as canonical 32-bit limbs
__device__ __forceinline__ bool fr_gt_half(const Fr& s) {
    const uint32_t H[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    // lexicographic compare from the top limb
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (s.l[i] > H[i]) return true;
        if (s.l[i] < H[i]) return false;
    }
    return false;
}

__device__ __forceinline__ uint32_t limb_bits(const Fr& s, uint32_t bit, uint32_t c) {
    // c <= 24 bits starting at `bit` (zero beyond bit 255)
    uint32_t w = bit >> 5, sh = bit & 31;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if ((uint32_t)i == w) lo = s.l[i];
        if ((uint32_t)i == w + 1) hi = s.l[i];
    }
    uint32_t v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
    return v & ((1u << c) - 1);
}

// ---- 1. decompose + histogram ------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_decompose_kernel(const uint4* __restrict__ scalars, MsmPlan pl,
                                                          uint32_t* __restrict__ digits, uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.n) return;
    Fr s = fp_from_mont(fp_load<FR>(scalars + 2 * (size_t)i));
    uint32_t neg = 0;
    if (fr_gt_half(s)) {
        // s <- r - s  (canonical, non-zero)
        Fr r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.l[k] = FpParams<FR>::P(k);
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t d = (uint64_t)r.l[k] - s.l[k] - borrow;
            s.l[k] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        neg = SIGN_BIT;
    }
    const uint32_t c = pl.c, half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < pl.W; ++w) {
        uint32_t d = limb_bits(s, w * c, c) + carry;
        uint32_t sign = neg;
        if (d > half) { d = (1u << c) - d; carry = 1; sign ^= SIGN_BIT; }
        else carry = 0;
        uint32_t entry = 0;
        if (d != 0) {
            entry = d | sign;
            atomicAdd(&counts[w * pl.Nb + d - 1], 1u);
        }
        digits[(size_t)w * pl.n + i] = entry;
    }
}

// ---- 2. exclusive scan (three small kernels) + scatter -------------------------------------------
static const uint32_t SCAN_BLOCK = 1024;   // elements per CTA (256 threads x 4)

__global__ void __launch_bounds__(256) scan_block_sums_kernel(const uint32_t* __restrict__ in, uint32_t count, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[8];
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < count) s += in[base + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 8; ++k) t += warp_sums[k];
        block_sums[blockIdx.x] = t;
    }
}
// single CTA: exclusive scan of block_sums in place; block_sums[nblocks] = total
__global__ void __launch_bounds__(1024) scan_top_kernel(uint32_t* block_sums, uint32_t nblocks) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t start = 0; start < nblocks; start += 1024) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < nblocks ? block_sums[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t o = 1; o < 1024; o <<= 1) {
            uint32_t add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        uint32_t incl = sh[threadIdx.x], carry = carry_sh;
        if (idx < nblocks) block_sums[idx] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry_sh;
}
// per CTA: local exclusive scan + block offset; writes `out` and a second copy `out2` (cursor)
__global__ void __launch_bounds__(256) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t count, const uint32_t* __restrict__ block_sums,
                                                       uint32_t* __restrict__ out, uint32_t* __restrict__ out2) {
    __shared__ uint32_t warp_sums[8];
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < count) ? in[base + k] : 0; s += v[k]; }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t k = 0; k < wid; ++k) woff += warp_sums[k];
    uint32_t run = block_sums[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < count) { out[base + k] = run; out2[base + k] = run; }
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) out[count] = block_sums[gridDim.x];
}

__global__ void __launch_bounds__(256) msm_scatter_kernel(MsmPlan pl, const uint32_t* __restrict__ digits, uint32_t* __restrict__ cursor,
                                                        uint32_t* __restrict__ sorted) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.n) return;
    for (uint32_t w = 0; w < pl.W; ++w) {
        uint32_t e = digits[(size_t)w * pl.n + i];
        if (e == 0) continue;
        uint32_t d = e & ~SIGN_BIT;
        uint32_t pos = atomicAdd(&cursor[w * pl.Nb + d - 1], 1u);
        sorted[pos] = i | (e & SIGN_BIT);
    }
}

// ---- 3. plan: split long buckets ------------------------------------------------------------------
// ctrl[0] = overflow tasks allocated, ctrl[1] = split buckets
__global__ void __launch_bounds__(256) msm_plan_kernel(MsmPlan pl, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ ctrl,
                                                     uint2* __restrict__ overflow_desc, uint2* __restrict__ bucket_extra, uint32_t* __restrict__ heavy_list) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pl.B) return;
    uint32_t cnt = offsets[b + 1] - offsets[b];
    uint32_t extra = cnt > pl.L ? (cnt - 1) / pl.L : 0;
    uint32_t slot0 = 0;
    if (extra) {
        slot0 = atomicAdd(&ctrl[0], extra);
        for (uint32_t s = 0; s < extra; ++s) overflow_desc[slot0 + s] = make_uint2(b, s + 1);
        heavy_list[atomicAdd(&ctrl[1], 1u)] = b;
    }
    bucket_extra[b] = make_uint2(slot0, extra);
}

// ---- 4. accumulate ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_accumulate_kernel(MsmPlan pl, const uint4* __restrict__ bases, const uint32_t* __restrict__ offsets,
                                                           const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ ctrl,
                                                           const uint2* __restrict__ overflow_desc, uint4* __restrict__ bucket_acc, uint4* __restrict__ partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b, seg;
    uint4* dst;
    if (t < pl.B) { b = t; seg = 0; dst = bucket_acc + 8 * (size_t)t; }
    else {
        uint32_t slot = t - pl.B;
        if (slot >= ctrl[0]) return;
        uint2 d = overflow_desc[slot];
        b = d.x; seg = d.y;
        dst = partial + 8 * (size_t)slot;
    }
    uint32_t beg = offsets[b], end = offsets[b + 1];
    beg += seg * pl.L;
    if (end - beg > pl.L) end = beg + pl.L;
    XYZZ acc = xyzz_identity();
    if (beg < end) {
        uint32_t e = sorted[beg];
        Affine p = affine_load(bases + 4 * (size_t)(e & ~SIGN_BIT));
        for (uint32_t j = beg; j < end; ++j) {
            uint32_t e_next = 0;
            Affine p_next = p;
            if (j + 1 < end) {      // prefetch the next point while this one is added
                e_next = sorted[j + 1];
                p_next = affine_load(bases + 4 * (size_t)(e_next & ~SIGN_BIT));
            }
            xyzz_add_affine(acc, p, (e & SIGN_BIT) != 0);
            e = e_next;
            p = p_next;
        }
    }
    xyzz_store(dst, acc);
}

// ---- 5. combine split buckets: one warp per split bucket ------------------------------------------
__device__ __forceinline__ XYZZ xyzz_shfl_down(const XYZZ& v, int delta) {
    XYZZ r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], delta);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], delta);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], delta);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], delta);
    }
    return r;
}

__global__ void __launch_bounds__(128) msm_combine_kernel(const uint32_t* __restrict__ ctrl, const uint32_t* __restrict__ heavy_list,
                                                        const uint2* __restrict__ bucket_extra, const uint4* __restrict__ partial, uint4* __restrict__ bucket_acc) {
    uint32_t lane = threadIdx.x & 31;
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    uint32_t nheavy = ctrl[1];
    for (uint32_t h = warp; h < nheavy; h += nwarps) {
        uint32_t b = heavy_list[h];
        uint2 ex = bucket_extra[b];
        XYZZ acc = xyzz_identity();
        for (uint32_t s = lane; s < ex.y; s += 32) {
            XYZZ q = xyzz_load(partial + 8 * (size_t)(ex.x + s));
            xyzz_add(acc, q);
        }
        for (int d = 16; d > 0; d >>= 1) {
            XYZZ o = xyzz_shfl_down(acc, d);
            if (lane < (uint32_t)d) xyzz_add(acc, o);
        }
        if (lane == 0) {
            XYZZ cur = xyzz_load(bucket_acc + 8 * (size_t)b);
            xyzz_add(cur, acc);
            xyzz_store(bucket_acc + 8 * (size_t)b, cur);
        }
    }
}

// ---- 6. bucket reduction ------------------------------------------------------------------------------
// One level of  S_w = sum_u (u+1) * Bw[u] + sum_u Dw[u]  over N items per window, m items per thread:
//   A_j = sum_i B[jm+i],  C_j = sum_i (i+1) B[jm+i] + sum_i D[jm+i]
//   S_w = sum_j C_j + m * sum_{j>=1} j * A_j  ->  next level: B'[j-1] = m*A_j, B'[J-1] = 0, D'[j] = C_j.
__global__ void __launch_bounds__(128) msm_reduce_level_kernel(const uint4* __restrict__ Bin, const uint4* __restrict__ Din, uint32_t N, uint32_t logm,
                                                             uint32_t W, uint4* __restrict__ Bout, uint4* __restrict__ Dout) {
    uint32_t m = 1u << logm;
    uint32_t J = (N + m - 1) >> logm;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * J) return;
    uint32_t w = t / J, j = t - w * J;
    const uint4* Bw = Bin + 8 * (size_t)w * N;
    const uint4* Dw = Din ? Din + 8 * (size_t)w * N : nullptr;
    uint32_t lo = j << logm, hi = lo + m;
    if (hi > N) hi = N;
    XYZZ running = xyzz_identity(), acc = xyzz_identity();
    for (uint32_t u = hi; u-- > lo;) {
        XYZZ bu = xyzz_load(Bw + 8 * (size_t)u);
        xyzz_add(running, bu);
        xyzz_add(acc, running);
        if (Dw) {
            XYZZ du = xyzz_load(Dw + 8 * (size_t)u);
            xyzz_add(acc, du);
        }
    }
    xyzz_store(Dout + 8 * ((size_t)w * J + j), acc);
    if (j >= 1) {
        for (uint32_t k = 0; k < logm; ++k) running = xyzz_double(running);
        xyzz_store(Bout + 8 * ((size_t)w * J + j - 1), running);
    } else {
        xyzz_store(Bout + 8 * ((size_t)w * J + J - 1), xyzz_identity());
    }
}

// Horner over the window sums (S[w] = Dfinal[w]) and conversion to a Jacobian triple
__global__ void msm_final_kernel(const uint4* __restrict__ S, uint32_t W, uint32_t c, uint4* __restrict__ out_jac, uint32_t accumulate) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    XYZZ acc = xyzz_load(S + 8 * (size_t)(W - 1));
    for (uint32_t w = W - 1; w-- > 0;) {
        for (uint32_t k = 0; k < c; ++k) acc = xyzz_double(acc);
        XYZZ sw = xyzz_load(S + 8 * (size_t)w);
        xyzz_add(acc, sw);
    }
    if (accumulate) {       // running total across sub-MSMs is kept in XYZZ right behind the Jacobian slot
        XYZZ prev = xyzz_load(out_jac + 6);
        xyzz_add(acc, prev);
    }
    xyzz_store(out_jac + 6, acc);
    Fq X, Y, Z;
    xyzz_to_jacobian(acc, X, Y, Z);
    fp_store<FQ>(out_jac, X);
    fp_store<FQ>(out_jac + 2, Y);
    fp_store<FQ>(out_jac + 4, Z);
}

// ---- host orchestration ----------------------------------------------------------------------------------
struct MsmScratch {
    DevBuf digits, counts, offsets, cursor, block_sums, sorted, ctrl, overflow_desc, bucket_extra, heavy, bucket_acc, partial, redA, redB, redC, redD, result;
};

static int g_forced_c = 0;
int msm_set_window(int c) {
    if (c != 0 && (c < 2 || c > 22)) { set_error("msm window must be 0 (auto) or in [2, 22]"); return H2B_ERR_BAD_ARGUMENT; }
    g_forced_c = c;
    return H2B_OK;
}

static uint32_t msm_pick_window(size_t n) {
    if (g_forced_c) return (uint32_t)g_forced_c;
    static int env_c = -1;
    if (env_c < 0) { const char* e = getenv("H2B_MSM_C"); env_c = e ? atoi(e) : 0; }
    if (env_c >= 2 && env_c <= 22) return (uint32_t)env_c;
    uint32_t best = 2;
    double best_cost = 1e300;
    for (uint32_t c = 2; c <= 20; ++c) {
        double W = 253 / c + 1;
        double cost = (double)n * W + 16.0 * W * (double)(1u << (c - 1));
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

static int exclusive_scan(MsmScratch& s, const uint32_t* in, uint32_t count, uint32_t* out, uint32_t* out2, cudaStream_t stream) {
    uint32_t nblocks = (count + SCAN_BLOCK - 1) / SCAN_BLOCK;
    H2B_TRY(s.block_sums.reserve(((size_t)nblocks + 1) * 4));
    uint32_t* bs = (uint32_t*)s.block_sums.p;
    H2B_LAUNCH(scan_block_sums_kernel, nblocks, 256, 0, stream, in, count, bs);
    H2B_LAUNCH(scan_top_kernel, 1, 1024, 0, stream, bs, nblocks);
    H2B_LAUNCH(scan_apply_kernel, nblocks, 256, 0, stream, in, count, (const uint32_t*)bs, out, out2);
    H2B_CUDA(cudaGetLastError());
    return H2B_OK;
}

static int msm_sub(DeviceCtx& ctx, MsmScratch& s, const void* d_scalars, const void* d_bases, uint32_t n, void* d_result, bool accumulate, cudaStream_t stream) {
    MsmPlan pl;
    pl.n = n;
    pl.c = msm_pick_window(n);
    pl.W = 253 / pl.c + 1;
    pl.Nb = 1u << (pl.c - 1);
    pl.B = pl.W * pl.Nb;
    uint32_t mean = (n + pl.Nb - 1) / pl.Nb;
    pl.L = 2 * mean < 32 ? 32 : 2 * mean;
    pl.max_overflow = (uint32_t)(((uint64_t)n * pl.W) / pl.L + 1);

    H2B_TRY(s.digits.reserve((size_t)n * pl.W * 4));
    H2B_TRY(s.counts.reserve((size_t)pl.B * 4));
    H2B_TRY(s.offsets.reserve(((size_t)pl.B + 1) * 4));
    H2B_TRY(s.cursor.reserve((size_t)pl.B * 4));
    H2B_TRY(s.sorted.reserve((size_t)n * pl.W * 4));
    H2B_TRY(s.ctrl.reserve(16));
    H2B_TRY(s.overflow_desc.reserve((size_t)pl.max_overflow * 8));
    H2B_TRY(s.bucket_extra.reserve((size_t)pl.B * 8));
    H2B_TRY(s.heavy.reserve((size_t)pl.max_overflow * 4));
    H2B_TRY(s.bucket_acc.reserve((size_t)pl.B * 128));
    H2B_TRY(s.partial.reserve((size_t)pl.max_overflow * 128));

    uint32_t* counts = (uint32_t*)s.counts.p;
    uint32_t* offsets = (uint32_t*)s.offsets.p;
    uint32_t* cursor = (uint32_t*)s.cursor.p;
    uint32_t* ctrl = (uint32_t*)s.ctrl.p;
    H2B_CUDA(cudaMemsetAsync(counts, 0, (size_t)pl.B * 4, stream));
    H2B_CUDA(cudaMemsetAsync(ctrl, 0, 16, stream));

    const uint32_t nblk = (n + 255) / 256;
    ctx.prof.mark(PROF_BEGIN, stream);
    H2B_LAUNCH(msm_decompose_kernel, nblk, 256, 0, stream, (const uint4*)d_scalars, pl, (uint32_t*)s.digits.p, counts);
    ctx.prof.mark(PROF_MSM_DECOMPOSE, stream);
    H2B_TRY(exclusive_scan(s, counts, pl.B, offsets, cursor, stream));
    ctx.prof.mark(PROF_MSM_SCAN, stream);
    H2B_LAUNCH(msm_scatter_kernel, nblk, 256, 0, stream, pl, (const uint32_t*)s.digits.p, cursor, (uint32_t*)s.sorted.p);
    ctx.prof.mark(PROF_MSM_SCATTER, stream);
    H2B_LAUNCH(msm_plan_kernel, (pl.B + 255) / 256, 256, 0, stream, pl, (const uint32_t*)offsets, ctrl, (uint2*)s.overflow_desc.p,
               (uint2*)s.bucket_extra.p, (uint32_t*)s.heavy.p);
    ctx.prof.mark(PROF_MSM_PLAN, stream);
    const uint32_t acc_threads = pl.B + pl.max_overflow;
    H2B_LAUNCH(msm_accumulate_kernel, (acc_threads + 255) / 256, 256, 0, stream, pl, (const uint4*)d_bases, (const uint32_t*)offsets,
               (const uint32_t*)s.sorted.p, (const uint32_t*)ctrl, (const uint2*)s.overflow_desc.p, (uint4*)s.bucket_acc.p, (uint4*)s.partial.p);
    ctx.prof.mark(PROF_MSM_ACCUMULATE, stream);
    H2B_LAUNCH(msm_combine_kernel, ctx.sm_count * 2, 128, 0, stream, (const uint32_t*)ctrl, (const uint32_t*)s.heavy.p,
               (const uint2*)s.bucket_extra.p, (const uint4*)s.partial.p, (uint4*)s.bucket_acc.p);
    H2B_CUDA(cudaGetLastError());
    ctx.prof.mark(PROF_MSM_COMBINE, stream);

    // bucket reduction hierarchy
    const uint32_t logm = 4;
    uint32_t N = pl.Nb;
    uint32_t J0 = (N + (1u << logm) - 1) >> logm;
    H2B_TRY(s.redA.reserve((size_t)pl.W * J0 * 128));
    H2B_TRY(s.redB.reserve((size_t)pl.W * J0 * 128));
    H2B_TRY(s.redC.reserve((size_t)pl.W * J0 * 128));
    H2B_TRY(s.redD.reserve((size_t)pl.W * J0 * 128));
    const uint4* Bin = (const uint4*)s.bucket_acc.p;
    const uint4* Din = nullptr;
    uint4* Bping[2] = {(uint4*)s.redA.p, (uint4*)s.redB.p};
    uint4* Dping[2] = {(uint4*)s.redC.p, (uint4*)s.redD.p};
    int pp = 0;
    for (;;) {
        uint32_t J = (N + (1u << logm) - 1) >> logm;
        uint32_t threads = pl.W * J;
        H2B_LAUNCH(msm_reduce_level_kernel, (threads + 127) / 128, 128, 0, stream, Bin, Din, N, logm, pl.W, Bping[pp], Dping[pp]);
        Bin = Bping[pp];
        Din = Dping[pp];
        pp ^= 1;
        N = J;
        if (J == 1) break;
    }
    ctx.prof.mark(PROF_MSM_REDUCE, stream);
    H2B_LAUNCH(msm_final_kernel, 1, 32, 0, stream, Din, pl.W, pl.c, (uint4*)d_result, accumulate ? 1u : 0u);
    H2B_CUDA(cudaGetLastError());
    ctx.prof.mark(PROF_MSM_FINAL, stream);
    return H2B_OK;
}

// d_out_jac: 96 bytes (x|y|z Montgomery). Internally a 224-byte result block is used (Jacobian + XYZZ total).
int msm_run(DeviceCtx& ctx, const void* d_scalars, const void* d_bases, size_t n, void* d_out_jac, bool with_xyzz, cudaStream_t stream) {
    if (!ctx.msm) ctx.msm = new MsmScratch();
    MsmScratch& s = *ctx.msm;
    H2B_TRY(s.result.reserve(256));
    const size_t out_bytes = with_xyzz ? 224 : 96;
    if (n == 0) {
        // identity: Jacobian (0, 1, 0), XYZZ all-zero
        uint32_t host[56];
        memset(host, 0, sizeof(host));
        for (int i = 0; i < 8; ++i) host[8 + i] = FpParams<FQ>::ONE(i);
        H2B_CUDA(cudaMemcpyAsync(d_out_jac, host, out_bytes, cudaMemcpyHostToDevice, stream));
        H2B_CUDA(cudaStreamSynchronize(stream));
        return H2B_OK;
    }
    if (!d_scalars || !d_bases || !d_out_jac) { set_error("msm: null pointer"); return H2B_ERR_BAD_ARGUMENT; }
    const size_t MAX_SUB = (size_t)1 << 26;
    bool first = true;
    for (size_t done = 0; done < n; done += MAX_SUB) {
        uint32_t m = (uint32_t)((n - done < MAX_SUB) ? (n - done) : MAX_SUB);
        H2B_TRY(msm_sub(ctx, s, (const char*)d_scalars + done * 32, (const char*)d_bases + done * 64, m, s.result.p, !first, stream));
        first = false;
    }
    H2B_CUDA(cudaMemcpyAsync(d_out_jac, s.result.p, out_bytes, cudaMemcpyDeviceToDevice, stream));
    return H2B_OK;
}

// sum of `count` partial results (224-byte blocks: Jacobian | XYZZ) -> Jacobian.  Used to fold the per-device
// partial sums of a point-range-sharded MSM.
__global__ void msm_sum_partials_kernel(const uint4* __restrict__ blocks, uint32_t count, uint4* __restrict__ out_jac) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    XYZZ acc = xyzz_identity();
    for (uint32_t i = 0; i < count; ++i) {
        XYZZ p = xyzz_load(blocks + 14 * (size_t)i + 6);
        xyzz_add(acc, p);
    }
    Fq X, Y, Z;
    xyzz_to_jacobian(acc, X, Y, Z);
    fp_store<FQ>(out_jac, X);
    fp_store<FQ>(out_jac + 2, Y);
    fp_store<FQ>(out_jac + 4, Z);
}
int msm_sum_partials_run(DeviceCtx& ctx, const void* d_blocks, uint32_t count, void* d_out_jac, cudaStream_t stream) {
    (void)ctx;
    H2B_LAUNCH(msm_sum_partials_kernel, 1, 32, 0, stream, (const uint4*)d_blocks, count, (uint4*)d_out_jac);
    H2B_CUDA(cudaGetLastError());
    return H2B_OK;
}

void msm_release(DeviceCtx& ctx) {
    if (!ctx.msm) return;
    MsmScratch& s = *ctx.msm;
    DevBuf* all[] = {&s.digits, &s.counts, &s.offsets, &s.cursor, &s.block_sums, &s.sorted, &s.ctrl, &s.overflow_desc, &s.bucket_extra,
                     &s.heavy, &s.bucket_acc, &s.partial, &s.redA, &s.redB, &s.redC, &s.redD, &s.result};
    for (DevBuf* b : all) b->release();
    delete ctx.msm;
    ctx.msm = nullptr;
}

}  // namespace h2b
