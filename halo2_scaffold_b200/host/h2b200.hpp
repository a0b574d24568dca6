// h2b200.hpp -- C++ host-side mirror of the reference's operator interface for the hot path, over the C ABI of
// include/h2b200.h.  The reference is Rust (halo2_proofs / halo2curves, un-vendored git dependencies of
// DCMMC/halo2-scaffold: Cargo.toml:13,16); this image has no Rust toolchain, so the host side above the C ABI is
// written in C++ with the SAME names, argument meaning and error behaviour as the Rust items it mirrors:
//
//   h2b200::arithmetic::best_multiexp   <- [UP] halo2_proofs::arithmetic::best_multiexp   
//   h2b200::arithmetic::best_fft        <- [UP] halo2_proofs::arithmetic::best_fft        (row a3)
//   h2b200::poly::EvaluationDomain      <- [UP] halo2_proofs::poly::EvaluationDomain      (row a6, Appendix B)
//   h2b200::poly::kzg::ParamsKZG        <- [UP] halo2_proofs::poly::kzg::commitment::ParamsKZG::{commit, commit_lagrange} (row a7)
//
// Rust `assert!`/`panic!` become h2b200::Panic (a std::logic_error); a non-zero return of the C ABI becomes
// h2b200::Panic carrying h2b_last_error(), exactly what the Rust shim of INTEGRATION.md does.
// Vector work runs on the GPU; only the per-domain scalar constants (omega, divisors, zeta powers) are computed on
// the host, as EvaluationDomain::new does in the reference.  Header-only; link with -lh2b200.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/h2b200.h"

namespace h2b200 {

struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};

inline void check(int rc, const char* what) {
    if (rc != H2B_OK) throw Panic(std::string(what) + ": " + h2b_last_error());
}

// halo2curves 0.3.x memory layouts (SURVEY.md section 8 "Sizes")
struct Fr { uint64_t l[4]; };                 // Montgomery, R = 2^256, fully reduced
struct Fq { uint64_t l[4]; };
struct G1Affine { Fq x, y; };                 // (0, 0) = identity
struct G1 { Fq x, y, z; };                    // Jacobian, z == 0 = identity
static_assert(sizeof(Fr) == 32 && sizeof(G1Affine) == 64 && sizeof(G1) == 96, "layouts must match halo2curves");

inline bool operator==(const Fr& a, const Fr& b) { return std::memcmp(&a, &b, sizeof(Fr)) == 0; }

// ---- scalar Fr arithmetic on the host (domain constants only) ----------------------------------------------------
namespace fr {
typedef unsigned __int128 u128;
static const uint64_t MODULUS[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t INV = 0xc2e1f593efffffffull;                 // -r^-1 mod 2^64
static const uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
static const uint32_t S = 28;                                      // two-adicity
// canonical (non-Montgomery) integers
static const uint64_t ROOT_OF_UNITY_CANON[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
static const uint64_t ZETA_CANON[4] = {0xb8ca0b2d36636f23ull, 0xcc37a73fec2bc5e9ull, 0x048b6e193fd84104ull, 0x30644e72e131a029ull};

inline bool geq_modulus(const uint64_t t[4]) {
    for (int i = 3; i >= 0; --i) {
        if (t[i] > MODULUS[i]) return true;
        if (t[i] < MODULUS[i]) return false;
    }
    return true;
}
// Montgomery product a * b / R mod r (coarsely integrated operand scanning)
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 carry = 0;
        for (int j = 0; j < 4; ++j) {
            u128 cur = (u128)a.l[j] * b.l[i] + t[j] + carry;
            t[j] = (uint64_t)cur;
            carry = cur >> 64;
        }
        u128 top = (u128)t[4] + carry;
        t[4] = (uint64_t)top;
        t[5] = (uint64_t)(top >> 64);
        uint64_t m = t[0] * INV;
        carry = ((u128)m * MODULUS[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) {
            u128 cur = (u128)m * MODULUS[j] + t[j] + carry;
            t[j - 1] = (uint64_t)cur;
            carry = cur >> 64;
        }
        top = (u128)t[4] + carry;
        t[3] = (uint64_t)top;
        t[4] = t[5] + (uint64_t)(top >> 64);
    }
    if (t[4] || geq_modulus(t)) {
        u128 borrow = 0;
        for (int i = 0; i < 4; ++i) {
            u128 d = (u128)t[i] - MODULUS[i] - borrow;
            t[i] = (uint64_t)d;
            borrow = (d >> 64) & 1;
        }
    }
    Fr r;
    std::memcpy(r.l, t, 32);
    return r;
}
inline Fr from_canonical(const uint64_t c[4]) {
    Fr a, r2;
    std::memcpy(a.l, c, 32);
    std::memcpy(r2.l, R2, 32);
    return mul(a, r2);
}
inline Fr from_u64(uint64_t v) {
    uint64_t c[4] = {v, 0, 0, 0};
    return from_canonical(c);
}
inline Fr one() { return from_u64(1); }
inline Fr square(const Fr& a) { return mul(a, a); }
inline Fr pow(const Fr& a, const uint64_t e[4]) {
    Fr r = one();
    for (int w = 3; w >= 0; --w)
        for (int i = 63; i >= 0; --i) {
            r = square(r);
            if ((e[w] >> i) & 1) r = mul(r, a);
        }
    return r;
}
inline Fr invert(const Fr& a) {                       // a^(r-2)
    uint64_t e[4] = {MODULUS[0] - 2, MODULUS[1], MODULUS[2], MODULUS[3]};
    return pow(a, e);
}
inline Fr root_of_unity() { return from_canonical(ROOT_OF_UNITY_CANON); }
inline Fr zeta() { return from_canonical(ZETA_CANON); }
}  // namespace fr

// ---- process-wide initialisation ---------------------------------------------------------------------------------
inline void ensure_init() {
    if (h2b_device_count() == 0) check(h2b_init(0), "h2b_init");
}

// ---- halo2_proofs::arithmetic -----------------------------------------------------------------------------------
namespace arithmetic {

// pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve
inline G1 best_multiexp(const Fr* coeffs, size_t coeffs_len, const G1Affine* bases, size_t bases_len) {
    if (coeffs_len != bases_len) throw Panic("assertion failed: `(left == right)` coeffs.len() == bases.len()");
    ensure_init();
    G1 out;
    check(h2b_msm_bn254_g1(reinterpret_cast<const uint64_t*>(coeffs), reinterpret_cast<const uint64_t*>(bases), coeffs_len,
                           reinterpret_cast<uint64_t*>(&out)), "best_multiexp");
    return out;
}
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    return best_multiexp(coeffs.data(), coeffs.size(), bases.data(), bases.size());
}

// pub fn best_fft<G: Group>(a: &mut [G], omega: G::Scalar, log_n: u32)      (G = Fr)
inline void best_fft(Fr* a, size_t len, const Fr& omega, uint32_t log_n) {
    if (log_n > 63 || len != ((size_t)1 << log_n)) throw Panic("assertion failed: a.len() == 1 << log_n");
    ensure_init();
    check(h2b_ntt_bn254_fr(reinterpret_cast<uint64_t*>(a), omega.l, log_n), "best_fft");
}
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) { best_fft(a.data(), a.size(), omega, log_n); }

}  // namespace arithmetic

// ---- a device buffer (RAII) for the device-resident callers --------------------------------------------------------
class DeviceBuffer {
  public:
    DeviceBuffer(int device, size_t bytes) : device_(device) { check(h2b_dev_alloc(device, bytes, &p_), "h2b_dev_alloc"); }
    ~DeviceBuffer() { if (p_) h2b_dev_free(device_, p_); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void* get() const { return p_; }
  private:
    int device_;
    void* p_ = nullptr;
};

namespace poly {

// pub struct EvaluationDomain<G: Group>; EvaluationDomain::new(j, k)
class EvaluationDomain {
  public:
    EvaluationDomain(uint32_t j, uint32_t k, int device = 0) : device_(device), k_(k) {
        ensure_init();
        if (j < 2) throw Panic("EvaluationDomain::new: j must be at least 2");
        quotient_poly_degree_ = j - 1;
        n_ = (uint64_t)1 << k;
        extended_k_ = k;
        while (((uint64_t)1 << extended_k_) < n_ * quotient_poly_degree_) ++extended_k_;
        if (extended_k_ > fr::S) throw Panic("EvaluationDomain::new: extended domain exceeds the two-adicity of the field");
        extended_omega_ = fr::root_of_unity();
        for (uint32_t i = extended_k_; i < fr::S; ++i) extended_omega_ = fr::square(extended_omega_);
        omega_ = extended_omega_;
        for (uint32_t i = k; i < extended_k_; ++i) omega_ = fr::square(omega_);
        omega_inv_ = fr::invert(omega_);
        extended_omega_inv_ = fr::invert(extended_omega_);
        ifft_divisor_ = fr::invert(fr::from_u64(n_));
        extended_ifft_divisor_ = fr::invert(fr::from_u64((uint64_t)1 << extended_k_));
        g_coset_ = fr::zeta();
        g_coset_inv_ = fr::square(g_coset_);
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    size_t extended_len() const { return (size_t)1 << extended_k_; }
    uint64_t get_quotient_poly_degree() const { return quotient_poly_degree_; }
    const Fr& get_omega() const { return omega_; }
    const Fr& get_omega_inv() const { return omega_inv_; }
    const Fr& get_extended_omega() const { return extended_omega_; }

    // pub fn lagrange_to_coeff(&self, a: Polynomial<_, LagrangeCoeff>) -> Polynomial<_, Coeff>
    std::vector<Fr> lagrange_to_coeff(const std::vector<Fr>& a) const {
        if (a.size() != n_) throw Panic("assertion failed: a.len() == 1 << self.k");
        return run(a, n_, n_, n_, [&](void* d) {
            check(h2b_lagrange_to_coeff_dev(device_, d, k_, omega_inv_.l, ifft_divisor_.l, nullptr), "lagrange_to_coeff");
        });
    }
    // pub fn coeff_to_extended(&self, a: Polynomial<_, Coeff>) -> Polynomial<_, ExtendedLagrangeCoeff>
    std::vector<Fr> coeff_to_extended(const std::vector<Fr>& a) const {
        if (a.size() != n_) throw Panic("assertion failed: a.len() == 1 << self.k");
        const Fr z[3] = {fr::one(), g_coset_, g_coset_inv_};
        // only the 2^k coefficients cross PCIe; scaling by zeta powers, zero-padding and the extended FFT run on the device
        return run(a, extended_len(), n_, extended_len(), [&](void* d) {
            check(h2b_coeff_to_extended_dev(device_, d, k_, extended_k_, extended_omega_.l, z[0].l, nullptr), "coeff_to_extended");
        });
    }
    // pub fn extended_to_coeff(&self, a: Polynomial<_, ExtendedLagrangeCoeff>) -> Vec<G>
    std::vector<Fr> extended_to_coeff(const std::vector<Fr>& a) const {
        if (a.size() != extended_len()) throw Panic("assertion failed: a.len() == self.extended_len()");
        const Fr z[3] = {extended_ifft_divisor_, fr::mul(extended_ifft_divisor_, g_coset_inv_), fr::mul(extended_ifft_divisor_, g_coset_)};
        return run(a, n_ * quotient_poly_degree_, extended_len(), extended_len(), [&](void* d) {
            check(h2b_extended_to_coeff_dev(device_, d, extended_k_, extended_omega_inv_.l, z[0].l, nullptr), "extended_to_coeff");
        });
    }

  private:
    // upload `upload_len` elements of `a` into a device buffer of `alloc_len`, run `steps`, download `out_len`
    template <class F>
    std::vector<Fr> run(const std::vector<Fr>& a, size_t out_len, size_t upload_len, size_t alloc_len, F steps) const {
        DeviceBuffer d(device_, alloc_len * sizeof(Fr));
        check(h2b_memcpy_h2d(device_, d.get(), a.data(), upload_len * sizeof(Fr)), "h2d");
        steps(d.get());
        check(h2b_dev_sync(device_), "sync");
        std::vector<Fr> out(out_len);
        check(h2b_memcpy_d2h(device_, out.data(), d.get(), out_len * sizeof(Fr)), "d2h");
        return out;
    }
    int device_;
    uint32_t k_, extended_k_;
    uint64_t n_, quotient_poly_degree_;
    Fr omega_, omega_inv_, extended_omega_, extended_omega_inv_, ifft_divisor_, extended_ifft_divisor_, g_coset_, g_coset_inv_;
};

namespace kzg {

// pub struct ParamsKZG<E: Engine> { k, n, g: Vec<G1Affine>, g_lagrange: Vec<G1Affine>, ... }
// Both SRS vectors are registered once (device resident, window tables precomputed).
class ParamsKZG {
  public:
    ParamsKZG(uint32_t k, const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange) : k_(k), n_((uint64_t)1 << k) {
        ensure_init();
        if (g.size() != n_ || g_lagrange.size() != n_) throw Panic("ParamsKZG: g and g_lagrange must hold 2^k points");
        check(h2b_register_bases(reinterpret_cast<const uint64_t*>(g.data()), g.size(), &g_), "register g");
        check(h2b_register_bases(reinterpret_cast<const uint64_t*>(g_lagrange.data()), g_lagrange.size(), &g_lagrange_), "register g_lagrange");
    }
    ~ParamsKZG() {
        if (g_) h2b_unregister_bases(g_);
        if (g_lagrange_) h2b_unregister_bases(g_lagrange_);
    }
    ParamsKZG(const ParamsKZG&) = delete;
    ParamsKZG& operator=(const ParamsKZG&) = delete;
    uint32_t k() const { return k_; }
    uint64_t n() const { return n_; }
    // fn commit(&self, poly: &Polynomial<_, Coeff>, _: Blind<_>) -> G1     (best_multiexp(&scalars, &self.g[..len]))
    G1 commit(const std::vector<Fr>& poly) const { return msm(poly, g_); }
    // fn commit_lagrange(&self, poly: &Polynomial<_, LagrangeCoeff>, _: Blind<_>) -> G1
    G1 commit_lagrange(const std::vector<Fr>& poly) const { return msm(poly, g_lagrange_); }

  private:
    G1 msm(const std::vector<Fr>& poly, uint64_t handle) const {
        if (poly.size() > n_) throw Panic("assertion failed: bases.len() >= size");
        G1 out;
        check(h2b_msm_bn254_g1_registered(reinterpret_cast<const uint64_t*>(poly.data()), handle, 0, poly.size(), reinterpret_cast<uint64_t*>(&out)), "commit");
        return out;
    }
    uint32_t k_;
    uint64_t n_;
    uint64_t g_ = 0, g_lagrange_ = 0;
};

}  // namespace kzg
}  // namespace poly
}  // namespace h2b200
