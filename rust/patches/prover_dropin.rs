// SURVEY.md 8f rank 1 -- the device-resident column pipeline inside the patched `halo2_proofs` (both forks), written to spec
// (no cargo in the build container).  Two small changes; the scaffold and every circuit stay untouched.
//
// (1) poly/kzg/commitment.rs [UP]: ParamsKZG registers its two vectors once and commits through the handles, so that
//     `commit` / `commit_lagrange` skip the per-call digest of the implicit cache.
//
//         pub struct ParamsKZG<E: Engine> { ..., h2b_g: OnceCell<u64>, h2b_g_lagrange: OnceCell<u64> }
//
//         fn commit_lagrange(&self, poly: &Polynomial<E::Scalar, LagrangeCoeff>, _: Blind<E::Scalar>) -> E::G1 {
//             let h = *self.h2b_g_lagrange.get_or_init(|| h2b200_sys::register_bases(cast_points(&self.g_lagrange)));
//             from_jacobian_words(h2b200_sys::msm_registered(cast_scalars(&poly.values), h, 0))
//         }
//
// (2) plonk/prover.rs [UP]: where create_proof commits an advice / permuted / product column in Lagrange form and then converts
//     it twice (`domain.lagrange_to_coeff`, later `domain.coeff_to_extended`), ONE call does all three with a single upload and
//     leaves the extended form on the device for evaluate_h (rust/patches/evaluation_dropin.rs):
use std::os::raw::c_void;

pub struct ColumnOnDevice {
    pub commitment: [u64; 12],        // Jacobian x | y | z, Montgomery: transmute to G1 exactly as in arithmetic_dropin.rs
    pub coeff: Vec<[u64; 4]>,         // coefficient form (kept on the host for the evaluations at x and the opening)
    pub d_extended: *mut c_void,      // 2^extended_k evaluations over the zeta coset, resident on `device`
    pub device: i32,
}
impl Drop for ColumnOnDevice {
    fn drop(&mut self) { unsafe { h2b200_sys::h2b_dev_free(self.device, self.d_extended); } }
}

/// `domain`: the EvaluationDomain's own constants (omega_inv, ifft_divisor, extended_omega, g_coset), as Montgomery words
pub fn commit_and_extend(device: i32, lagrange: &[[u64; 4]], handle_g_lagrange: u64, k: u32, extended_k: u32, omega_inv: &[u64; 4],
                         ifft_divisor: &[u64; 4], extended_omega: &[u64; 4], zeta_powers: &[[u64; 4]; 3]) -> ColumnOnDevice {
    assert_eq!(lagrange.len(), 1usize << k);
    h2b200_sys::ensure_init();
    let mut out = ColumnOnDevice { commitment: [0; 12], coeff: vec![[0u64; 4]; 1 << k], d_extended: std::ptr::null_mut(), device };
    let rc = unsafe {
        h2b200_sys::h2b_column_pipeline(device, lagrange.as_ptr() as *const u64, handle_g_lagrange, k, extended_k, omega_inv.as_ptr(), ifft_divisor.as_ptr(),
                                        extended_omega.as_ptr(), zeta_powers.as_ptr() as *const u64, out.commitment.as_mut_ptr(),
                                        out.coeff.as_mut_ptr() as *mut u64, std::ptr::null_mut(), &mut out.d_extended)
    };
    if rc != 0 { panic!("h2b_column_pipeline: {}", h2b200_sys::last_error_string()); }
    out
}

// The independent commitments of one prover phase (all advice columns; both permuted columns of every lookup; all grand products)
// go through `h2b_msm_bn254_g1_batch_registered` in one call: see INTEGRATION.md section 6.
