// BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM kernels.
//
// Bucket sums are kept in extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// the mixed addition of an affine SRS point costs 8M + 2S with no inversion. The identity is
// ZZ = 0.  Affine inputs follow halo2curves' G1Affine layout (x|y Montgomery, (0,0) = identity,
// [UP] halo2curves 0.3.x src/derive/curve.rs; SURVEY.md section 8 "Sizes").  Every routine is
// complete: P + P, P + (-P) and identity operands are handled, because real witness columns make
// equal points meet in a bucket (duplicate bases, repeated small scalars).
#pragma once
#include "field.cuh"

namespace h2b {

struct Affine {
    Fq x, y;
};
struct XYZZ {
    Fq x, y, zz, zzz;
};

__device__ __forceinline__ bool affine_is_identity(const Affine& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
__device__ __forceinline__ bool xyzz_is_identity(const XYZZ& p) { return fp_is_zero(p.zz); }
__device__ __forceinline__ XYZZ xyzz_identity() {
    XYZZ r;
    r.x = fp_zero<FQ>(); r.y = fp_zero<FQ>(); r.zz = fp_zero<FQ>(); r.zzz = fp_zero<FQ>();
    return r;
}
__device__ __forceinline__ Affine affine_load(const void* p) {
    Affine a;
    a.x = fp_load<FQ>(p);
    a.y = fp_load<FQ>(reinterpret_cast<const uint4*>(p) + 2);
    return a;
}
__device__ __forceinline__ void affine_store(void* p, const Affine& a) {
    fp_store<FQ>(p, a.x);
    fp_store<FQ>(reinterpret_cast<uint4*>(p) + 2, a.y);
}
__device__ __forceinline__ XYZZ xyzz_load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    XYZZ r;
    r.x = fp_load<FQ>(q); r.y = fp_load<FQ>(q + 2); r.zz = fp_load<FQ>(q + 4); r.zzz = fp_load<FQ>(q + 6);
    return r;
}
__device__ __forceinline__ void xyzz_store(void* p, const XYZZ& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    fp_store<FQ>(q, v.x); fp_store<FQ>(q + 2, v.y); fp_store<FQ>(q + 4, v.zz); fp_store<FQ>(q + 6, v.zzz);
}
__device__ __forceinline__ XYZZ xyzz_from_affine(const Affine& p) {
    XYZZ r;
    if (affine_is_identity(p)) return xyzz_identity();
    r.x = p.x; r.y = p.y; r.zz = fp_one<FQ>(); r.zzz = fp_one<FQ>();
    return r;
}

// 2 * (affine point), result in XYZZ  (EFD mdbl-2008-s-1, a = 0)
__device__ __forceinline__ XYZZ xyzz_double_affine(const Affine& p) {
    XYZZ r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P in XYZZ  (EFD dbl-2008-s-1, a = 0).  y = 0 never happens on this curve (odd order).
__device__ __forceinline__ XYZZ xyzz_double(const XYZZ& p) {
    if (xyzz_is_identity(p)) return p;
    XYZZ r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// acc += (neg ? -q : q), q affine  (EFD madd-2008-s: 8M + 2S)
__device__ __forceinline__ void xyzz_add_affine(XYZZ& acc, const Affine& q_in, bool neg) {
    if (affine_is_identity(q_in)) return;
    Affine q = q_in;
    if (neg) q.y = fp_neg(q.y);
    if (xyzz_is_identity(acc)) {
        acc.x = q.x; acc.y = q.y; acc.zz = fp_one<FQ>(); acc.zzz = fp_one<FQ>();
        return;
    }
    Fq u2 = fp_mul(q.x, acc.zz);
    Fq s2 = fp_mul(q.y, acc.zzz);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_double_affine(q);
        else acc = xyzz_identity();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq qq = fp_mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
    // Y3 = R*(Q - X3) - Y1*PPP.  fp_mul2 (two products, one Montgomery reduction) saves 72 of this addition's 1260
    // FMA-pipe instructions, but measured 2 % SLOWER here (profiles/r01_field_primitives.jsonl): at the 16 warps per
    // SM this kernel's 126 registers allow, the longer dependency chains of the separated product/reduction form
    // cost more than the saved dispatch cycles.  Kept behind a switch for kernels with more resident warps.
#ifdef H2B_MADD_MUL2
    Fq y3 = fp_mul2(r, fp_sub(qq, x3), fp_neg(acc.y), ppp);
#else
    Fq y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(acc.y, ppp));
#endif
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}

// acc += b, both XYZZ  (EFD add-2008-s: 12M + 2S)
__device__ __forceinline__ void xyzz_add(XYZZ& acc, const XYZZ& b) {
    if (xyzz_is_identity(b)) return;
    if (xyzz_is_identity(acc)) { acc = b; return; }
    Fq u1 = fp_mul(acc.x, b.zz);
    Fq u2 = fp_mul(b.x, acc.zz);
    Fq s1 = fp_mul(acc.y, b.zzz);
    Fq s2 = fp_mul(b.y, acc.zzz);
    Fq p = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_double(acc);
        else acc = xyzz_identity();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq qq = fp_mul(u1, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(fp_mul(acc.zz, b.zz), pp);
    acc.zzz = fp_mul(fp_mul(acc.zzz, b.zzz), ppp);
}

// ---- quad-cooperative group law (latency-bound kernels) ----------------------------------------------------------------
// The bucket reduction, the combine of cut buckets and the final Horner are chains of DEPENDENT group additions run by a
// handful of warps: a lone warp needs ~2300 clocks per field multiplication (every IMAD.WIDE.X waits for the carry of the one
// before), so a 14-multiplication addition is ~15 us deep and a 2^16-point MSM spends 0.7 of its 0.95 ms there.  Here FOUR
// consecutive lanes hold the same operands and each computes one of the (up to four) independent multiplications of a step;
// the products are exchanged by quad-wide shuffles.  An XYZZ addition becomes 4 multiplication steps instead of 14, a
// doubling 3 instead of 9.  Results are bit-identical (same field operations).  Every lane executes every shuffle: operands
// that need no arithmetic (identity, P + (-P)) are resolved after the last exchange, and callers give idle quads identities.
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << (threadIdx.x & 28u); }
#ifdef H2B_EMU
// The kernel-logic emulator implements a shuffle as two warp-wide fiber barriers: the exchanges below would make the CPU suite
// ~8x slower.  There every lane of a quad computes the whole operation itself unless H2B_EMU_QUAD=1 (one test exercises the
// shuffled form); the index math of the callers (one quad per item, idle quads on identities) is the same either way.
static inline bool emu_quad_exchange() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("H2B_EMU_QUAD"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
#endif

// lane q of the quad multiplies x[q] * y[q]; all four products are returned to every lane
__device__ __forceinline__ void quad_mul(uint32_t q, unsigned mask, const Fq (&x)[4], const Fq (&y)[4], Fq (&out)[4]) {
    Fq a, b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a.l[i] = q == 0 ? x[0].l[i] : (q == 1 ? x[1].l[i] : (q == 2 ? x[2].l[i] : x[3].l[i]));
        b.l[i] = q == 0 ? y[0].l[i] : (q == 1 ? y[1].l[i] : (q == 2 ? y[2].l[i] : y[3].l[i]));
    }
    const Fq m = fp_mul(a, b);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) out[s].l[i] = __shfl_sync(mask, m.l[i], s, 4);
}

// acc += b, both XYZZ, computed by the quad (acc and b replicated in its four lanes)
__device__ __forceinline__ void xyzz_add_quad(XYZZ& acc, const XYZZ& b, uint32_t q, unsigned mask) {
#ifdef H2B_EMU
    if (!emu_quad_exchange()) { xyzz_add(acc, b); return; }
#endif
    const bool b_id = xyzz_is_identity(b), a_id = xyzz_is_identity(acc);
    Fq o[4];
    {
        const Fq x[4] = {acc.x, b.x, acc.y, b.y}, y[4] = {b.zz, acc.zz, b.zzz, acc.zzz};
        quad_mul(q, mask, x, y, o);
    }
    const Fq u1 = o[0], s1 = o[2];
    const Fq p = fp_sub(o[1], u1), r = fp_sub(o[3], s1);
    {
        const Fq x[4] = {p, r, acc.zz, acc.zzz}, y[4] = {p, r, b.zz, b.zzz};
        quad_mul(q, mask, x, y, o);
    }
    const Fq pp = o[0], rr = o[1];
    {
        const Fq x[4] = {p, u1, o[2], o[3]}, y[4] = {pp, pp, pp, p};
        quad_mul(q, mask, x, y, o);
    }
    const Fq ppp = o[0], qq = o[1], zz3 = o[2], w = o[3];
    const Fq x3 = fp_sub(fp_sub(rr, ppp), fp_dbl(qq));
    {
        const Fq x[4] = {r, s1, w, w}, y[4] = {fp_sub(qq, x3), ppp, pp, pp};
        quad_mul(q, mask, x, y, o);
    }
    if (b_id) return;
    if (a_id) { acc = b; return; }
    if (fp_is_zero(p)) {
        if (fp_is_zero(r)) acc = xyzz_double(acc);      // P + P: the plain doubling, computed by every lane (rare)
        else acc = xyzz_identity();
        return;
    }
    acc.x = x3;
    acc.y = fp_sub(o[0], o[1]);
    acc.zz = zz3;
    acc.zzz = o[2];
}

// 2 * P by the quad
__device__ __forceinline__ XYZZ xyzz_double_quad(const XYZZ& p, uint32_t q, unsigned mask) {
#ifdef H2B_EMU
    if (!emu_quad_exchange()) return xyzz_double(p);
#endif
    Fq o[4];
    const Fq u = fp_dbl(p.y);
    {
        const Fq x[4] = {u, p.x, u, p.x}, y[4] = {u, p.x, u, p.x};
        quad_mul(q, mask, x, y, o);
    }
    const Fq v = o[0], xx = o[1];
    const Fq m = fp_add(fp_dbl(xx), xx);
    {
        const Fq x[4] = {u, p.x, m, v}, y[4] = {v, v, m, p.zz};
        quad_mul(q, mask, x, y, o);
    }
    const Fq w = o[0], s = o[1], zz3 = o[3];
    const Fq x3 = fp_sub(o[2], fp_dbl(s));
    {
        const Fq x[4] = {m, w, w, w}, y[4] = {fp_sub(s, x3), p.y, p.zzz, p.zzz};
        quad_mul(q, mask, x, y, o);
    }
    if (xyzz_is_identity(p)) return p;
    XYZZ r;
    r.x = x3;
    r.y = fp_sub(o[0], o[1]);
    r.zz = zz3;
    r.zzz = o[2];
    return r;
}

// XYZZ -> a Jacobian representative (X', Y', Z') of the same point without an inversion:
// Z' = ZZ*ZZZ, X' = X*ZZ*ZZZ^2, Y' = Y*ZZ^3*ZZZ^2  (then X'/Z'^2 = X/ZZ and Y'/Z'^3 = Y/ZZZ).
// Identity -> (0, 1, 0), matching halo2curves' G1::identity().
__device__ __forceinline__ void xyzz_to_jacobian(const XYZZ& p, Fq& X, Fq& Y, Fq& Z) {
    if (xyzz_is_identity(p)) { X = fp_zero<FQ>(); Y = fp_one<FQ>(); Z = fp_zero<FQ>(); return; }
    Fq zzz2 = fp_sqr(p.zzz);
    Fq t = fp_mul(p.zz, zzz2);          // ZZ * ZZZ^2
    X = fp_mul(p.x, t);
    Y = fp_mul(fp_mul(p.y, fp_sqr(p.zz)), t);
    Z = fp_mul(p.zz, p.zzz);
}

// XYZZ -> affine (one field inversion; used for a handful of points only)
__device__ __forceinline__ Affine xyzz_to_affine(const XYZZ& p) {
    Affine a;
    if (xyzz_is_identity(p)) { a.x = fp_zero<FQ>(); a.y = fp_zero<FQ>(); return a; }
    Fq zi = fp_inv(fp_mul(p.zz, p.zzz));     // 1/(ZZ*ZZZ)
    a.x = fp_mul(p.x, fp_mul(p.zzz, zi));    // X/ZZ
    a.y = fp_mul(p.y, fp_mul(p.zz, zi));     // Y/ZZZ
    return a;
}

}  // namespace h2b
