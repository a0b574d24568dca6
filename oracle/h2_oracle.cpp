// ORACLE (test infrastructure only -- never linked or loaded by the product path).
//
// C++17 CPU restatement of the hot path that DCMMC/halo2-scaffold reaches through
// halo2_proofs: `best_multiexp` / `multiexp_serial` and
// `best_fft` / `recursive_butterfly_arithmetic`, plus the BN254 Fr / Fq / G1
// arithmetic of halo2curves that they run on.
//
// PARITY UNPINNED: the algorithm lives in third-party git dependencies that are absent
// from /root/reference (halo2_proofs @ PSE tag v2023_02_02, Cargo.toml:13; Axiom fork
// `axiom/dev` via halo2-base, Cargo.toml:16; halo2curves 0.3.x transitive) and the
// reference tree holds no golden vector for this path. This file restates
// the published algorithm; it is pinned by oracle/bn254.py
// (independent big-int implementation, naive MSM and O(n^2) DFT) and by the external
// anchors listed there. Reference call sites: src/scaffold.rs:132,135,191-199,207-214,
// 223-230,284,287,322-346,354-361; examples/standard_plonk.rs:33,34,41-49,57-64.
//
// Also used (by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs only) as the timed CPU baseline: "C++ restatement of halo2_proofs
// v2023_02_02", std::thread in place of rayon.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include <map>
#include <algorithm>

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------
// 4x64 Montgomery field, R = 2^256  ([UP] halo2curves/src/derive/field.rs, bn256/{fr,fq}.rs)
// ---------------------------------------------------------------------------------
struct FqParams {
    static constexpr u64 M[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr u64 INV = 0x87d20782e4866389ULL;
    static constexpr u64 R1[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
    static constexpr u64 R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
};
struct FrParams {
    static constexpr u64 M[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static constexpr u64 INV = 0xc2e1f593efffffffULL;
    static constexpr u64 R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
    static constexpr u64 R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
};

template <class P>
struct Fp {
    u64 l[4];

    static Fp zero() { Fp r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
    static Fp one() { Fp r; memcpy(r.l, P::R1, 32); return r; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }

    static inline bool geq_mod(const u64* a) {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > P::M[i]) return true;
            if (a[i] < P::M[i]) return false;
        }
        return true;
    }
    static inline void sub_mod(u64* a) {
        u128 b = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)a[i] - P::M[i] - (u64)b;
            a[i] = (u64)d;
            b = (d >> 64) & 1;
        }
    }
    Fp operator+(const Fp& o) const {
        Fp r; u128 c = 0;
        for (int i = 0; i < 4; ++i) { c += (u128)l[i] + o.l[i]; r.l[i] = (u64)c; c >>= 64; }
        if (geq_mod(r.l)) sub_mod(r.l);       // 2p < 2^256: no carry out
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; u128 b = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)l[i] - o.l[i] - (u64)b;
            r.l[i] = (u64)d; b = (d >> 64) & 1;
        }
        if (b) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + P::M[i]; r.l[i] = (u64)c; c >>= 64; } }
        return r;
    }
    Fp neg() const { return zero() - *this; }
    Fp dbl() const { return *this + *this; }
    // CIOS Montgomery product
    Fp operator*(const Fp& o) const {
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) { c += (u128)l[j] * o.l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
            c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
            u64 m = t[0] * P::INV;
            c = (u128)m * P::M[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) { c += (u128)m * P::M[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
            c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
        }
        Fp r; memcpy(r.l, t, 32);
        if (t[4] || geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    Fp to_mont() const { Fp r2; memcpy(r2.l, P::R2, 32); return *this * r2; }
    Fp from_mont() const { Fp o; o.l[0] = 1; o.l[1] = o.l[2] = o.l[3] = 0; return *this * o; }
    Fp pow(const u64 e[4]) const {
        Fp r = one();
        for (int i = 255; i >= 0; --i) { r = r.sqr(); if ((e[i / 64] >> (i % 64)) & 1) r = r * *this; }
        return r;
    }
    Fp inv() const {   // Fermat; 0 -> 0
        u64 e[4]; memcpy(e, P::M, 32); e[0] -= 2;
        return pow(e);
    }
};
typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

// ---------------------------------------------------------------------------------
// G1: y^2 = x^3 + 3.  Affine (0,0) = identity; Jacobian z = 0 = identity.
// ([UP] halo2curves/src/derive/curve.rs; any complete group law gives the same affine value)
// ---------------------------------------------------------------------------------
struct G1Affine { Fq x, y; bool is_identity() const { return x.is_zero() && y.is_zero(); } };
struct G1 {
    Fq x, y, z;
    static G1 identity() { G1 r; r.x = Fq::zero(); r.y = Fq::one(); r.z = Fq::zero(); return r; }
    bool is_identity() const { return z.is_zero(); }
};

static G1 g1_double(const G1& p) {
    if (p.is_identity()) return p;
    Fq a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
    Fq d = ((p.x + b).sqr() - a - c).dbl();
    Fq e = a + a + a, f = e.sqr();
    G1 r;
    r.z = (p.y * p.z).dbl();
    r.x = f - d.dbl();
    r.y = e * (d - r.x) - c.dbl().dbl().dbl();
    return r;
}
static G1 g1_add_mixed(const G1& p, const G1Affine& q) {
    if (q.is_identity()) return p;
    if (p.is_identity()) { G1 r; r.x = q.x; r.y = q.y; r.z = Fq::one(); return r; }
    Fq z1z1 = p.z.sqr(), u2 = q.x * z1z1, s2 = q.y * p.z * z1z1;
    if (u2 == p.x) {
        if (s2 == p.y) return g1_double(p);
        return G1::identity();
    }
    Fq h = u2 - p.x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - p.y).dbl(), v = p.x * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (p.y * j).dbl();
    r.z = (p.z + h).sqr() - z1z1 - hh;
    return r;
}
static G1 g1_add(const G1& p, const G1& q) {
    if (p.is_identity()) return q;
    if (q.is_identity()) return p;
    Fq z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
    Fq u1 = p.x * z2z2, u2 = q.x * z1z1, s1 = p.y * z2z2 * q.z, s2 = q.y * z1z1 * p.z;
    if (u1 == u2) {
        if (s1 == s2) return g1_double(p);
        return G1::identity();
    }
    Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (s1 * j).dbl();
    r.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
    return r;
}
static G1Affine g1_to_affine(const G1& p) {
    G1Affine r;
    if (p.is_identity()) { r.x = Fq::zero(); r.y = Fq::zero(); return r; }
    Fq zi = p.z.inv(), zi2 = zi.sqr();
    r.x = p.x * zi2; r.y = p.y * zi2 * zi;
    return r;
}

// ---------------------------------------------------------------------------------
// multiexp_serial / best_multiexp   ([UP] halo2_proofs/src/arithmetic.rs @ v2023_02_02)
// ---------------------------------------------------------------------------------
static inline u64 get_at(size_t segment, size_t c, const uint8_t* bytes /*32 LE*/) {
    size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t len = std::min<size_t>(8, 32 - skip_bytes);
    memcpy(v, bytes + skip_bytes, len);
    u64 tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    return tmp % ((u64)1 << c);
}

struct Bucket {   // [UP] enum Bucket { None, Affine(C), Projective(C::Curve) }
    int state; G1Affine a; G1 p;
    void add_assign(const G1Affine& o) {
        if (state == 0) { a = o; state = 1; }
        else if (state == 1) { G1 t; t.x = a.x; t.y = a.y; t.z = a.is_identity() ? Fq::zero() : Fq::one(); p = g1_add_mixed(t, o); state = 2; }
        else p = g1_add_mixed(p, o);
    }
    G1 add_to(const G1& acc) const {
        if (state == 0) return acc;
        if (state == 1) return g1_add_mixed(acc, a);
        return g1_add(acc, p);
    }
};

static void multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t m, G1& acc) {
    std::vector<uint8_t> reprs(m * 32);
    for (size_t i = 0; i < m; ++i) { Fr c = coeffs[i].from_mont(); memcpy(&reprs[i * 32], c.l, 32); }
    size_t c;
    if (m < 4) c = 1; else if (m < 32) c = 3; else c = (size_t)std::ceil(std::log((double)m));
    size_t segments = 256 / c + 1;
    std::vector<Bucket> buckets(((size_t)1 << c) - 1);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t k = 0; k < c; ++k) acc = g1_double(acc);
        for (auto& b : buckets) b.state = 0;
        for (size_t i = 0; i < m; ++i) {
            u64 d = get_at(seg, c, &reprs[i * 32]);
            if (d != 0) buckets[d - 1].add_assign(bases[i]);
        }
        G1 running = G1::identity();
        for (size_t b = buckets.size(); b-- > 0;) {
            running = buckets[b].add_to(running);
            acc = g1_add(acc, running);
        }
    }
}

static G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if (n > (size_t)threads) {
        size_t chunk = n / threads;
        size_t nchunks = (n + chunk - 1) / chunk;
        std::vector<G1> parts(nchunks, G1::identity());
        std::vector<std::thread> th;
        for (size_t t = 0; t < nchunks; ++t) {
            size_t lo = t * chunk, len = std::min(chunk, n - lo);
            th.emplace_back([=, &parts] { multiexp_serial(coeffs + lo, bases + lo, len, parts[t]); });
        }
        for (auto& x : th) x.join();
        G1 acc = G1::identity();
        for (auto& p : parts) acc = g1_add(acc, p);
        return acc;
    }
    G1 acc = G1::identity();
    multiexp_serial(coeffs, bases, n, acc);
    return acc;
}

// ---------------------------------------------------------------------------------
// best_fft / recursive_butterfly_arithmetic   ([UP] halo2_proofs/src/arithmetic.rs)
// ---------------------------------------------------------------------------------
static inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static void recursive_butterfly(Fr* a, size_t n, size_t twiddle_chunk, const Fr* tw, int par_depth) {
    if (n == 2) { Fr t = a[1]; a[1] = a[0] - t; a[0] = a[0] + t; return; }
    size_t half = n / 2;
    if (par_depth > 0) {       // rayon::join
        std::thread other([=] { recursive_butterfly(a + half, half, twiddle_chunk * 2, tw, par_depth - 1); });
        recursive_butterfly(a, half, twiddle_chunk * 2, tw, par_depth - 1);
        other.join();
    } else {
        recursive_butterfly(a, half, twiddle_chunk * 2, tw, 0);
        recursive_butterfly(a + half, half, twiddle_chunk * 2, tw, 0);
    }
    Fr* l = a; Fr* r = a + half;
    // case i = 0: twiddle is one
    { Fr t = r[0]; r[0] = l[0] - t; l[0] = l[0] + t; }
    auto body = [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) { Fr t = r[i] * tw[i * twiddle_chunk]; r[i] = l[i] - t; l[i] = l[i] + t; }
    };
    if (par_depth > 0 && half >= 4096) {     // the combine loop of the top levels, split like `parallelize`
        int T = 1 << par_depth;
        std::vector<std::thread> th;
        size_t per = (half + T - 1) / T;
        for (int t = 0; t < T; ++t) {
            size_t lo = std::max<size_t>(1, t * per), hi = std::min(half, (t + 1) * per);
            if (lo < hi) th.emplace_back(body, lo, hi);
        }
        for (auto& x : th) x.join();
    } else body(1, half);
}
static void best_fft(Fr* a, const Fr& omega, unsigned log_n, int threads) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; ++k) { size_t rk = bitreverse(k, log_n); if (k < rk) std::swap(a[k], a[rk]); }
    if (n < 2) return;
    std::vector<Fr> tw(n / 2);
    tw[0] = Fr::one();
    for (size_t i = 1; i < n / 2; ++i) tw[i] = tw[i - 1] * omega;
    int depth = 0;
    while ((1 << (depth + 1)) <= threads) ++depth;
    if ((unsigned)depth >= log_n) depth = log_n > 1 ? (int)log_n - 1 : 0;
    recursive_butterfly(a, n, 1, tw.data(), depth);
}

// ---------------------------------------------------------------------------------
// deterministic synthetic inputs
// ---------------------------------------------------------------------------------
static inline u64 splitmix64_at(u64 seed, u64 index) {   // value #index (0-based) of the SplitMix64 stream
    u64 z = seed + (index + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

extern "C" {

int orc_hardware_threads() { int t = (int)std::thread::hardware_concurrency(); return t > 0 ? t : 1; }

// out_jac: x|y|z Montgomery (12 x u64); z == 0 <=> identity
void orc_best_multiexp(const u64* scalars, const u64* bases, size_t n, int threads, u64* out_jac) {
    G1 r = best_multiexp((const Fr*)scalars, (const G1Affine*)bases, n, threads);
    memcpy(out_jac, &r, 96);
}
void orc_best_fft(u64* a, const u64* omega, uint32_t log_n, int threads) {
    Fr w; memcpy(w.l, omega, 32);
    best_fft((Fr*)a, w, log_n, threads);
}
void orc_g1_to_affine(const u64* jac, u64* out_aff) {
    G1 p; memcpy(&p, jac, 96);
    G1Affine a = g1_to_affine(p);
    memcpy(out_aff, &a, 64);
}
// sum of `count` Jacobian points
void orc_g1_sum(const u64* jacs, size_t count, u64* out_jac) {
    G1 acc = G1::identity();
    for (size_t i = 0; i < count; ++i) { G1 p; memcpy(&p, jacs + 12 * i, 96); acc = g1_add(acc, p); }
    memcpy(out_jac, &acc, 96);
}
// element-wise helpers used to check the device field arithmetic: op 0 add, 1 sub, 2 mul; field 0 = Fr, 1 = Fq
void orc_field_op(int field, int op, const u64* a, const u64* b, size_t n, u64* out) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) {
            Fr x, y, z; memcpy(x.l, a + 4 * i, 32); memcpy(y.l, b + 4 * i, 32);
            z = op == 0 ? x + y : op == 1 ? x - y : x * y; memcpy(out + 4 * i, z.l, 32);
        } else {
            Fq x, y, z; memcpy(x.l, a + 4 * i, 32); memcpy(y.l, b + 4 * i, 32);
            z = op == 0 ? x + y : op == 1 ? x - y : x * y; memcpy(out + 4 * i, z.l, 32);
        }
    }
}
void orc_fr_to_mont(const u64* in, size_t n, u64* out) {
    for (size_t i = 0; i < n; ++i) { Fr x; memcpy(x.l, in + 4 * i, 32); x = x.to_mont(); memcpy(out + 4 * i, x.l, 32); }
}
void orc_fr_from_mont(const u64* in, size_t n, u64* out) {
    for (size_t i = 0; i < n; ++i) { Fr x; memcpy(x.l, in + 4 * i, 32); x = x.from_mont(); memcpy(out + 4 * i, x.l, 32); }
}
// a[i] *= s  (Montgomery), used by the EvaluationDomain wrappers' checks
void orc_fr_scale(u64* a, size_t n, const u64* s) {
    Fr w; memcpy(w.l, s, 32);
    for (size_t i = 0; i < n; ++i) { Fr x; memcpy(x.l, a + 4 * i, 32); x = x * w; memcpy(a + 4 * i, x.l, 32); }
}
// ---- grand-product building blocks ([UP] halo2_proofs plonk/permutation/prover.rs and
// plonk/lookup/prover.rs build z(X) as a running product of numerator / denominator, the denominators inverted with
// ff::BatchInvert, which leaves zeros untouched) ---------------------------------------------------------------------
// a[i] <- 1 / a[i]; a[i] == 0 stays 0
void orc_fr_batch_invert(u64* a, size_t n) {
    std::vector<Fr> pre(n);
    Fr run = Fr::one();
    for (size_t i = 0; i < n; ++i) {
        Fr x; memcpy(x.l, a + 4 * i, 32);
        pre[i] = run;
        if (!x.is_zero()) run = run * x;
    }
    Fr inv = run.inv();
    for (size_t i = n; i-- > 0;) {
        Fr x; memcpy(x.l, a + 4 * i, 32);
        if (x.is_zero()) continue;
        Fr xi = inv * pre[i];
        inv = inv * x;
        memcpy(a + 4 * i, xi.l, 32);
    }
}
// out[0] = 1, out[i] = a[0] * ... * a[i-1]  (n outputs: the exclusive running product that z(omega^i) is)
void orc_fr_prefix_product(const u64* a, size_t n, u64* out) {
    Fr run = Fr::one();
    for (size_t i = 0; i < n; ++i) {
        memcpy(out + 4 * i, run.l, 32);
        Fr x; memcpy(x.l, a + 4 * i, 32);
        run = run * x;
    }
}
// [UP] halo2_proofs::arithmetic::eval_polynomial: sum_i poly[i] * point^i (Horner from the top)
void orc_fr_eval_polynomial(const u64* a, size_t n, const u64* x, u64* out) {
    Fr xx; memcpy(xx.l, x, 32);
    Fr r = Fr::zero();
    for (size_t i = n; i-- > 0;) { Fr c; memcpy(c.l, a + 4 * i, 32); r = r * xx + c; }
    memcpy(out, r.l, 32);
}
// [UP] halo2_proofs::arithmetic::kate_division(a, b): the quotient of a(X) by (X - b), n - 1 coefficients:
// q[n-2] = a[n-1], q[i-1] = a[i] + b * q[i]
void orc_fr_kate_division(const u64* a, size_t n, const u64* b, u64* q) {
    if (n <= 1) return;
    Fr bb; memcpy(bb.l, b, 32);
    Fr run = Fr::zero();
    for (size_t i = n - 1; i >= 1; --i) {
        Fr c; memcpy(c.l, a + 4 * i, 32);
        run = c + bb * run;
        memcpy(q + 4 * (i - 1), run.l, 32);
    }
}
}  // extern "C"

// ---------------------------------------------------------------------------------
// Quotient evaluation ([UP] halo2_proofs/src/plonk/evaluation.rs), restated as upstream
// runs it: one `intermediates` vector, the calculations walked in order for every row.
// Flat encoding shared with tests: a value source is 3 u32 (kind, index, rotation), a calculation 10 u32
// (op, target, a[3], b[3], parts_offset, parts_len); kinds / ops numbered as the enums are declared upstream:
//   ValueSource: 0 Constant, 1 Intermediate, 2 Fixed, 3 Advice, 4 Instance, 5 Challenge, 6 Beta, 7 Gamma, 8 Theta, 9 Y,
//                10 PreviousValue;   Calculation: 0 Add, 1 Sub, 2 Mul, 3 Square, 4 Double, 5 Negate, 6 Horner, 7 Store
// ---------------------------------------------------------------------------------
struct OrcGraph {
    const u64* constants; uint32_t n_constants;
    const int32_t* rotations; uint32_t n_rotations;
    const uint32_t* calcs; uint32_t n_calcs;
    const uint32_t* parts; uint32_t n_parts;
    uint32_t n_intermediates;
};
struct OrcColumns {
    const u64* const* fixed; const u64* const* advice; const u64* const* instance;
    const u64* challenges; const u64 *beta, *gamma, *theta, *y;
};
static inline Fr fr_at(const u64* col, size_t i) { Fr x; memcpy(x.l, col + 4 * i, 32); return x; }
// [UP] evaluation.rs get_rotation_idx
static inline size_t get_rotation_idx(size_t idx, int32_t rot, int32_t rot_scale, int64_t isize) {
    int64_t v = ((int64_t)idx + (int64_t)rot * rot_scale) % isize;
    if (v < 0) v += isize;
    return (size_t)v;
}
struct OrcEvalData { std::vector<Fr> intermediates; std::vector<size_t> rotations; };
// [UP] evaluation.rs ValueSource::get
static Fr value_source_get(const uint32_t* vs, const OrcGraph& g, const OrcColumns& c, const OrcEvalData& d, const Fr& previous_value) {
    switch (vs[0]) {
        case 0: return fr_at(g.constants, vs[1]);
        case 1: return d.intermediates[vs[1]];
        case 2: return fr_at(c.fixed[vs[1]], d.rotations[vs[2]]);
        case 3: return fr_at(c.advice[vs[1]], d.rotations[vs[2]]);
        case 4: return fr_at(c.instance[vs[1]], d.rotations[vs[2]]);
        case 5: return fr_at(c.challenges, vs[1]);
        case 6: return fr_at(c.beta, 0);
        case 7: return fr_at(c.gamma, 0);
        case 8: return fr_at(c.theta, 0);
        case 9: return fr_at(c.y, 0);
        default: return previous_value;
    }
}
// [UP] evaluation.rs GraphEvaluator::evaluate
static Fr graph_evaluate(const OrcGraph& g, const OrcColumns& c, OrcEvalData& d, const Fr& previous_value, size_t idx, int32_t rot_scale, int64_t isize) {
    for (uint32_t r = 0; r < g.n_rotations; ++r) d.rotations[r] = get_rotation_idx(idx, g.rotations[r], rot_scale, isize);
    for (uint32_t i = 0; i < g.n_calcs; ++i) {
        const uint32_t* k = g.calcs + 10 * (size_t)i;
        Fr a = value_source_get(k + 2, g, c, d, previous_value), v;
        switch (k[0]) {
            case 0: v = a + value_source_get(k + 5, g, c, d, previous_value); break;
            case 1: v = a - value_source_get(k + 5, g, c, d, previous_value); break;
            case 2: v = a * value_source_get(k + 5, g, c, d, previous_value); break;
            case 3: v = a.sqr(); break;
            case 4: v = a.dbl(); break;
            case 5: v = a.neg(); break;
            case 6: {
                Fr factor = value_source_get(k + 5, g, c, d, previous_value);
                v = a;
                for (uint32_t j = 0; j < k[9]; ++j) v = v * factor + value_source_get(g.parts + 3 * (size_t)(k[8] + j), g, c, d, previous_value);
                break;
            }
            default: v = a;
        }
        d.intermediates[k[1]] = v;
    }
    return g.n_calcs ? d.intermediates[g.calcs[10 * (size_t)(g.n_calcs - 1) + 1]] : Fr::zero();
}
static OrcGraph make_graph(const u64* constants, uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations, const uint32_t* calcs, uint32_t n_calcs,
                           const uint32_t* parts, uint32_t n_parts, uint32_t n_intermediates) {
    return OrcGraph{constants, n_constants, rotations, n_rotations, calcs, n_calcs, parts, n_parts, n_intermediates};
}

extern "C" {
// the "Custom gates" loop of evaluate_h: values[idx] = custom_gates.evaluate(.., previous_value = values[idx], idx, rot_scale, isize)
void orc_evaluate_graph(const u64* constants, uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations, const uint32_t* calcs, uint32_t n_calcs,
                        const uint32_t* parts, uint32_t n_parts, uint32_t n_intermediates, const u64* const* fixed, const u64* const* advice,
                        const u64* const* instance, const u64* challenges, const u64* beta, const u64* gamma, const u64* theta, const u64* y, u64* values,
                        uint32_t size, int32_t rot_scale) {
    OrcGraph g = make_graph(constants, n_constants, rotations, n_rotations, calcs, n_calcs, parts, n_parts, n_intermediates);
    OrcColumns c{fixed, advice, instance, challenges, beta, gamma, theta, y};
    OrcEvalData d{std::vector<Fr>(n_intermediates, Fr::zero()), std::vector<size_t>(n_rotations, 0)};
    for (size_t idx = 0; idx < size; ++idx) {
        Fr v = graph_evaluate(g, c, d, fr_at(values, idx), idx, rot_scale, size);
        memcpy(values + 4 * idx, v.l, 32);
    }
}
// one iteration of the "Lookups" loop of evaluate_h
void orc_evaluate_h_lookup(const u64* constants, uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations, const uint32_t* calcs, uint32_t n_calcs,
                           const uint32_t* parts, uint32_t n_parts, uint32_t n_intermediates, const u64* const* fixed, const u64* const* advice,
                           const u64* const* instance, const u64* challenges, const u64* beta_w, const u64* gamma_w, const u64* theta, const u64* y_w,
                           u64* values, uint32_t size, int32_t rot_scale, const u64* product_coset, const u64* permuted_input_coset,
                           const u64* permuted_table_coset, const u64* l0, const u64* l_last, const u64* l_active_row) {
    OrcGraph g = make_graph(constants, n_constants, rotations, n_rotations, calcs, n_calcs, parts, n_parts, n_intermediates);
    OrcColumns c{fixed, advice, instance, challenges, beta_w, gamma_w, theta, y_w};
    OrcEvalData d{std::vector<Fr>(n_intermediates, Fr::zero()), std::vector<size_t>(n_rotations, 0)};
    const Fr beta = fr_at(beta_w, 0), gamma = fr_at(gamma_w, 0), y = fr_at(y_w, 0), one = Fr::one();
    for (size_t idx = 0; idx < size; ++idx) {
        Fr table_value = graph_evaluate(g, c, d, Fr::zero(), idx, rot_scale, size);
        size_t r_next = get_rotation_idx(idx, 1, rot_scale, size), r_prev = get_rotation_idx(idx, -1, rot_scale, size);
        Fr a_minus_s = fr_at(permuted_input_coset, idx) - fr_at(permuted_table_coset, idx);
        Fr value = fr_at(values, idx);
        value = value * y + ((one - fr_at(product_coset, idx)) * fr_at(l0, idx));
        value = value * y + ((fr_at(product_coset, idx) * fr_at(product_coset, idx) - fr_at(product_coset, idx)) * fr_at(l_last, idx));
        value = value * y + ((fr_at(product_coset, r_next) * (fr_at(permuted_input_coset, idx) + beta) * (fr_at(permuted_table_coset, idx) + gamma) -
                              fr_at(product_coset, idx) * table_value) * fr_at(l_active_row, idx));
        value = value * y + (a_minus_s * fr_at(l0, idx));
        value = value * y + (a_minus_s * (fr_at(permuted_input_coset, idx) - fr_at(permuted_input_coset, r_prev)) * fr_at(l_active_row, idx));
        memcpy(values + 4 * idx, value.l, 32);
    }
}
// the "Permutations" loop of evaluate_h
void orc_evaluate_h_permutation(u64* values, uint32_t size, int32_t rot_scale, const u64* const* product_cosets, uint32_t n_sets, const u64* const* columns,
                                const u64* const* perm_cosets, uint32_t n_columns, uint32_t chunk_len, int32_t last_rotation, const u64* l0, const u64* l_last,
                                const u64* l_active_row, const u64* beta_w, const u64* gamma_w, const u64* y_w, const u64* delta_w, const u64* zeta_w,
                                const u64* extended_omega_w) {
    if (n_sets == 0) return;
    const Fr beta = fr_at(beta_w, 0), gamma = fr_at(gamma_w, 0), y = fr_at(y_w, 0), delta = fr_at(delta_w, 0), one = Fr::one();
    const Fr extended_omega = fr_at(extended_omega_w, 0);
    const Fr delta_start = beta * fr_at(zeta_w, 0);
    Fr beta_term = Fr::one();                                   // extended_omega^start with start = 0 (one chunk)
    for (size_t idx = 0; idx < size; ++idx) {
        size_t r_next = get_rotation_idx(idx, 1, rot_scale, size), r_last = get_rotation_idx(idx, last_rotation, rot_scale, size);
        Fr value = fr_at(values, idx);
        value = value * y + ((one - fr_at(product_cosets[0], idx)) * fr_at(l0, idx));
        const u64* last = product_cosets[n_sets - 1];
        value = value * y + ((fr_at(last, idx) * fr_at(last, idx) - fr_at(last, idx)) * fr_at(l_last, idx));
        for (uint32_t s = 1; s < n_sets; ++s)
            value = value * y + ((fr_at(product_cosets[s], idx) - fr_at(product_cosets[s - 1], r_last)) * fr_at(l0, idx));
        Fr current_delta = delta_start * beta_term;
        for (uint32_t s = 0, c0 = 0; s < n_sets && c0 < n_columns; ++s, c0 += chunk_len) {      // sets.zip(columns.chunks(chunk_len))
            uint32_t c1 = std::min(c0 + chunk_len, n_columns);
            Fr left = fr_at(product_cosets[s], r_next);
            for (uint32_t c = c0; c < c1; ++c) left = left * (fr_at(columns[c], idx) + beta * fr_at(perm_cosets[c], idx) + gamma);
            Fr right = fr_at(product_cosets[s], idx);
            for (uint32_t c = c0; c < c1; ++c) { right = right * (fr_at(columns[c], idx) + current_delta + gamma); current_delta = current_delta * delta; }
            value = value * y + ((left - right) * fr_at(l_active_row, idx));
        }
        beta_term = beta_term * extended_omega;
        memcpy(values + 4 * idx, value.l, 32);
    }
}
}  // extern "C"

extern "C" {
// [UP] plonk/permutation/prover.rs Argument::commit, one set: modified_values = prod_j (beta s_j + gamma + v_j); batch_invert;
// *= prod_j (deltaomega beta + gamma + v_j) with deltaomega running over omega^i and delta^j; z[0] = last_z, z[row] = z[row-1] * modified[row-1]
void orc_permutation_product(const u64* const* values, const u64* const* sigma, uint32_t m, size_t n, const u64* beta_w, const u64* gamma_w,
                             const u64* delta_w, const u64* deltaomega_w, const u64* omega_w, const u64* last_z_w, u64* z) {
    const Fr beta = fr_at(beta_w, 0), gamma = fr_at(gamma_w, 0), delta = fr_at(delta_w, 0), omega = fr_at(omega_w, 0);
    std::vector<Fr> modified(n, Fr::one());
    for (uint32_t j = 0; j < m; ++j)
        for (size_t i = 0; i < n; ++i) modified[i] = modified[i] * (beta * fr_at(sigma[j], i) + gamma + fr_at(values[j], i));
    if (n) orc_fr_batch_invert((u64*)modified.data(), n);
    Fr deltaomega_col = fr_at(deltaomega_w, 0);
    for (uint32_t j = 0; j < m; ++j) {
        Fr deltaomega = deltaomega_col;
        for (size_t i = 0; i < n; ++i) {
            modified[i] = modified[i] * (deltaomega * beta + gamma + fr_at(values[j], i));
            deltaomega = deltaomega * omega;
        }
        deltaomega_col = deltaomega_col * delta;
    }
    Fr run = fr_at(last_z_w, 0);
    for (size_t row = 0; row < n; ++row) { memcpy(z + 4 * row, run.l, 32); run = run * modified[row]; }
}
// [UP] plonk/lookup/prover.rs Permuted::commit_product
void orc_lookup_product(const u64* compressed_input, const u64* compressed_table, const u64* permuted_input, const u64* permuted_table, size_t n,
                        const u64* beta_w, const u64* gamma_w, u64* z) {
    const Fr beta = fr_at(beta_w, 0), gamma = fr_at(gamma_w, 0);
    std::vector<Fr> lookup_product(n);
    for (size_t i = 0; i < n; ++i) lookup_product[i] = (fr_at(permuted_input, i) + beta) * (fr_at(permuted_table, i) + gamma);
    if (n) orc_fr_batch_invert((u64*)lookup_product.data(), n);
    for (size_t i = 0; i < n; ++i) lookup_product[i] = lookup_product[i] * ((fr_at(compressed_input, i) + beta) * (fr_at(compressed_table, i) + gamma));
    Fr run = Fr::one();
    for (size_t i = 0; i < n; ++i) { memcpy(z + 4 * i, run.l, 32); run = run * lookup_product[i]; }
}
// [UP] plonk/lookup/prover.rs permute_expression_pair on the first `usable` rows.  Fr's Ord compares canonical integers.
// Returns 0, or 1 when an input value is not in the table (upstream: Err(ConstraintSystemFailure)).
int orc_lookup_permute(const u64* input, const u64* table, size_t usable, u64* permuted_input, u64* permuted_table) {
    struct Canon {
        u64 l[4];
        bool operator<(const Canon& o) const { for (int i = 3; i >= 0; --i) if (l[i] != o.l[i]) return l[i] < o.l[i]; return false; }
        bool operator==(const Canon& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    };
    auto canon = [](const u64* col, size_t i) { Fr x = fr_at(col, i).from_mont(); Canon c; memcpy(c.l, x.l, 32); return c; };
    auto store = [](u64* col, size_t i, const Canon& c) { Fr x; memcpy(x.l, c.l, 32); x = x.to_mont(); memcpy(col + 4 * i, x.l, 32); };
    std::vector<Canon> permuted_input_expression(usable);
    for (size_t i = 0; i < usable; ++i) permuted_input_expression[i] = canon(input, i);
    std::sort(permuted_input_expression.begin(), permuted_input_expression.end());
    std::map<Canon, uint32_t> leftover_table_map;
    for (size_t i = 0; i < usable; ++i) leftover_table_map[canon(table, i)] += 1;
    std::vector<Canon> permuted_table_coeffs(usable, Canon{{0, 0, 0, 0}});
    std::vector<size_t> repeated_input_rows;
    for (size_t row = 0; row < usable; ++row) {
        const Canon& input_value = permuted_input_expression[row];
        if (row == 0 || !(input_value == permuted_input_expression[row - 1])) {
            permuted_table_coeffs[row] = input_value;
            auto it = leftover_table_map.find(input_value);
            if (it == leftover_table_map.end() || it->second == 0) return 1;
            it->second -= 1;
        } else {
            repeated_input_rows.push_back(row);
        }
    }
    for (auto& kv : leftover_table_map)
        for (uint32_t c = 0; c < kv.second; ++c) { permuted_table_coeffs[repeated_input_rows.back()] = kv.first; repeated_input_rows.pop_back(); }
    for (size_t i = 0; i < usable; ++i) { store(permuted_input, i, permuted_input_expression[i]); store(permuted_table, i, permuted_table_coeffs[i]); }
    return 0;
}
// ---------------------------------------------------------------------------------
// SRS point encodings ([UP] halo2curves 0.3.x GroupEncoding::{to_bytes, from_bytes} for G1Affine; upstream's
// ParamsKZG::read_custom decompresses with `parallelize`, here std::thread).  Returns the index of the first invalid
// encoding, or n.
// ---------------------------------------------------------------------------------
size_t orc_g1_from_bytes(const unsigned char* bytes, size_t n, int threads, u64* out_aff) {
    static const u64 E[4] = {0x4f082305b61f3f52ULL, 0x65e05aa45a1c72a3ULL, 0x6e14116da0605617ULL, 0x0c19139cb84c680aULL};   // (p + 1) / 4
    if (threads < 1) threads = 1;
    std::vector<size_t> bad(threads, n);
    std::vector<std::thread> th;
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=, &bad] {
            const Fq three = Fq::one() + Fq::one() + Fq::one();
            for (size_t i = lo; i < hi; ++i) {
                Fq x; memcpy(x.l, bytes + 32 * i, 32);
                const bool ysign = (x.l[3] >> 63) & 1;
                x.l[3] &= ~(1ULL << 63);
                G1Affine p; p.x = Fq::zero(); p.y = Fq::zero();
                bool ok = !Fq::geq_mod(x.l);
                if (ok && !(x.is_zero() && !ysign)) {
                    Fq xm = x.to_mont();
                    Fq rhs = xm.sqr() * xm + three;
                    Fq y = rhs.pow(E);
                    ok = y.sqr() == rhs;
                    if (ok) {
                        if (((y.from_mont().l[0] & 1) != 0) != ysign) y = y.neg();
                        p.x = xm; p.y = y;
                    }
                }
                if (!ok && bad[t] == n) bad[t] = i;
                memcpy(out_aff + 8 * i, &p, 64);
            }
        });
    }
    for (auto& x : th) x.join();
    size_t first = n;
    for (size_t b : bad) first = std::min(first, b);
    return first;
}
void orc_g1_to_bytes(const u64* aff, size_t n, unsigned char* out) {
    for (size_t i = 0; i < n; ++i) {
        Fq x, y; memcpy(x.l, aff + 8 * i, 32); memcpy(y.l, aff + 8 * i + 4, 32);
        Fq c = Fq::zero();
        if (!(x.is_zero() && y.is_zero())) { c = x.from_mont(); c.l[3] |= (y.from_mont().l[0] & 1) << 63; }
        memcpy(out + 32 * i, c.l, 32);
    }
}
// uniform scalars in Montgomery form: 512-bit SplitMix64 draw reduced mod r (same stream as oracle/bn254.py random_fr)
void orc_random_fr(u64 seed, size_t n, u64* out) {
    Fr two64; two64.l[0] = 0; two64.l[1] = 1; two64.l[2] = two64.l[3] = 0; two64 = two64.to_mont();
    for (size_t i = 0; i < n; ++i) {
        Fr acc = Fr::zero();
        for (int w = 0; w < 8; ++w) {           // big-endian word order: v = (v << 64) | w
            Fr x; x.l[0] = splitmix64_at(seed, 8 * i + w); x.l[1] = x.l[2] = x.l[3] = 0;
            acc = acc * two64 + x.to_mont();
        }
        memcpy(out + 4 * i, acc.l, 32);
    }
}
// synthetic bases: P_i = [z_i] G, z_i = SplitMix64 value #i of stream `seed` (64-bit multiples of the generator),
// written as affine Montgomery x|y. Multi-threaded, batch-normalised.
void orc_gen_points(u64 seed, size_t n, int threads, u64* out_aff) {
    G1Affine g; g.x = Fq::one(); g.y = Fq::one().dbl();
    std::vector<G1Affine> pow2(64);
    {
        G1 t; t.x = g.x; t.y = g.y; t.z = Fq::one();
        for (int j = 0; j < 64; ++j) { pow2[j] = g1_to_affine(t); t = g1_double(t); }
    }
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=, &pow2] {
            std::vector<G1> jac(hi - lo);
            for (size_t i = lo; i < hi; ++i) {
                u64 z = splitmix64_at(seed, i);
                G1 acc = G1::identity();
                for (int j = 0; j < 64; ++j) if ((z >> j) & 1) acc = g1_add_mixed(acc, pow2[j]);
                jac[i - lo] = acc;
            }
            // batch inversion of z
            size_t m = hi - lo;
            std::vector<Fq> pre(m);
            Fq run = Fq::one();
            for (size_t i = 0; i < m; ++i) { pre[i] = run; if (!jac[i].z.is_zero()) run = run * jac[i].z; }
            Fq inv = run.inv();
            for (size_t i = m; i-- > 0;) {
                G1Affine a;
                if (jac[i].z.is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }
                else {
                    Fq zi = inv * pre[i]; inv = inv * jac[i].z;
                    Fq zi2 = zi.sqr(); a.x = jac[i].x * zi2; a.y = jac[i].y * zi2 * zi;
                }
                memcpy(out_aff + 8 * (lo + i), &a, 64);
            }
        });
    }
    for (auto& x : th) x.join();
}

}  // extern "C"
