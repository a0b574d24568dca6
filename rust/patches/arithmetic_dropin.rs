// Replacement bodies for `best_multiexp` and `best_fft` in halo2_proofs/src/arithmetic.rs -- applied identically to
//   (1) privacy-scaling-explorations/halo2 @ tag v2023_02_02   (reference Cargo.toml:13; used by examples/standard_plonk.rs)
//   (2) axiom-crypto/halo2 @ branch axiom/dev                   (reference Cargo.toml:16 via halo2-base; used by src/scaffold.rs)
// Public signatures are unchanged, so keygen_vk / keygen_pk / create_proof / verify_proof, src/scaffold.rs and every
// example compile unchanged.  The original generic bodies are kept under new names for non-BN254 instantiations
// (`best_multiexp_generic`, `best_fft_generic`), which the scaffold never reaches.
//
// Written to spec; NOT compiled in the build container (no cargo/rustc).  See INTEGRATION.md for the [patch] stanzas.
use std::any::TypeId;
use std::mem::size_of;

use group::Group as _;
use halo2curves::bn256::{Fr, G1Affine, G1};
use halo2curves::CurveAffine;

pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve {
    assert_eq!(coeffs.len(), bases.len());
    if TypeId::of::<C>() == TypeId::of::<G1Affine>() {
        // Fr = [u64; 4] Montgomery, G1Affine = x | y, G1 = x | y | z: identical to the device layouts
        debug_assert_eq!(size_of::<C::Scalar>(), 32);
        debug_assert_eq!(size_of::<C>(), 64);
        debug_assert_eq!(size_of::<C::Curve>(), 96);
        let scalars = unsafe { std::slice::from_raw_parts(coeffs.as_ptr() as *const [u64; 4], coeffs.len()) };
        let points = unsafe { std::slice::from_raw_parts(bases.as_ptr() as *const [u64; 8], bases.len()) };
        let out: [u64; 12] = h2b200_sys::msm_bn254_g1(scalars, points);
        // SAFETY: C::Curve == G1 here (checked through TypeId above), 96 bytes, plain old data
        return unsafe { std::mem::transmute_copy::<[u64; 12], C::Curve>(&out) };
    }
    best_multiexp_generic(coeffs, bases)
}

pub fn best_fft<G: Group>(a: &mut [G], omega: G::Scalar, log_n: u32) {
    if TypeId::of::<G>() == TypeId::of::<Fr>() {
        debug_assert_eq!(size_of::<G>(), 32);
        let data = unsafe { std::slice::from_raw_parts_mut(a.as_mut_ptr() as *mut [u64; 4], a.len()) };
        let w: [u64; 4] = unsafe { std::mem::transmute_copy::<G::Scalar, [u64; 4]>(&omega) };
        h2b200_sys::ntt_bn254_fr(data, &w, log_n);
        return;
    }
    best_fft_generic(a, omega, log_n)
}

// `best_multiexp_generic` / `best_fft_generic`: the upstream bodies, renamed, unchanged.
