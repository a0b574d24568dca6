//! `extern "C"` declarations of include/h2b200.h plus the two safe wrappers the `halo2_proofs` patches call.
//!
//! Layout contract (halo2curves 0.3.x): `Fr`/`Fq` are `#[repr(transparent)]`-compatible `[u64; 4]` in Montgomery form,
//! `G1Affine` is `x | y` (64 bytes, (0,0) = identity), `G1` is `x | y | z` Jacobian (96 bytes).  Slices of those types are
//! therefore passed by pointer cast, no conversion and no copy on the host.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

/// Row range of a sharded `evaluate_h` call (include/h2b200.h `h2b_eval_shard`).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct h2b_eval_shard { pub row0: u32, pub rows: u32, pub halo: u32 }
/// `enum ValueSource` of halo2_proofs/src/plonk/evaluation.rs, flattened: kind = variant index in declaration order
/// (Constant, Intermediate, Fixed, Advice, Instance, Challenge, Beta, Gamma, Theta, Y, PreviousValue).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct h2b_value_source { pub kind: u32, pub index: u32, pub rotation: u32 }
/// `enum Calculation` + `CalculationInfo::target`: op = variant index (Add, Sub, Mul, Square, Double, Negate, Horner, Store).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct h2b_calculation { pub op: u32, pub target: u32, pub a: h2b_value_source, pub b: h2b_value_source, pub parts_offset: u32, pub parts_len: u32 }
#[repr(C)]
pub struct h2b_graph {
    pub constants: *const u64, pub n_constants: u32,
    pub rotations: *const i32, pub n_rotations: u32,
    pub calculations: *const h2b_calculation, pub n_calculations: u32,
    pub parts: *const h2b_value_source, pub n_parts: u32,
    pub n_intermediates: u32,
}
#[repr(C)]
pub struct h2b_eval_columns {
    pub fixed: *const *const c_void, pub n_fixed: u32,
    pub advice: *const *const c_void, pub n_advice: u32,
    pub instance: *const *const c_void, pub n_instance: u32,
    pub challenges: *const u64, pub n_challenges: u32,
    pub beta: *const u64, pub gamma: *const u64, pub theta: *const u64, pub y: *const u64,
}

extern "C" {
    pub fn h2b_init(n_devices: c_int) -> c_int;
    pub fn h2b_init_device(device: c_int) -> c_int;
    pub fn h2b_shutdown();
    pub fn h2b_device_count() -> c_int;
    pub fn h2b_last_error() -> *const c_char;
    pub fn h2b_msm_bn254_g1(scalars: *const u64, bases: *const u64, n: usize, out_jac: *mut u64) -> c_int;
    pub fn h2b_ntt_bn254_fr(a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn h2b_register_bases(bases: *const u64, n: usize, handle: *mut u64) -> c_int;
    pub fn h2b_unregister_bases(handle: u64) -> c_int;
    pub fn h2b_msm_bn254_g1_registered(scalars: *const u64, handle: u64, offset: usize, n: usize, out_jac: *mut u64) -> c_int;
    pub fn h2b_ntt_bn254_fr_dev(device: c_int, d_a: *mut c_void, omega: *const u64, log_n: u32, stream: *mut c_void) -> c_int;
    pub fn h2b_msm_bn254_g1_dev(device: c_int, d_scalars: *const c_void, d_bases: *const c_void, n: usize, d_out_jac: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_msm_bn254_g1_dev_registered(device: c_int, d_scalars: *const c_void, handle: u64, offset: usize, n: usize, d_out_block: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_scale_dev(device: c_int, d_a: *mut c_void, n: usize, factors: *const u64, count: c_int, stream: *mut c_void) -> c_int;
    // EvaluationDomain conversions and the prover's column primitives on device-resident columns
    pub fn h2b_lagrange_to_coeff_dev(device: c_int, d_a: *mut c_void, k: u32, omega_inv: *const u64, ifft_divisor: *const u64, stream: *mut c_void) -> c_int;
    pub fn h2b_coeff_to_extended_dev(device: c_int, d_a: *mut c_void, k: u32, extended_k: u32, extended_omega: *const u64, zeta_powers: *const u64, stream: *mut c_void) -> c_int;
    pub fn h2b_extended_to_coeff_dev(device: c_int, d_a: *mut c_void, extended_k: u32, extended_omega_inv: *const u64, factors: *const u64, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_batch_invert_dev(device: c_int, d_a: *mut c_void, n: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_prefix_product_dev(device: c_int, d_in: *const c_void, d_out: *mut c_void, n: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_eval_polynomial_dev(device: c_int, d_coeffs: *const c_void, n: usize, x: *const u64, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_kate_division_dev(device: c_int, d_a: *const c_void, n: usize, b: *const u64, d_q: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_transpose_dev(device: c_int, d_in: *const c_void, d_out: *mut c_void, rows: u32, cols: u32, stream: *mut c_void) -> c_int;
    pub fn h2b_fr_lincomb_dev(device: c_int, d_cols: *const *const c_void, coeffs: *const u64, m: u32, n: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_permutation_product_dev(device: c_int, d_values: *const *const c_void, d_permutations: *const *const c_void, n_columns: u32, n: usize,
                                       beta: *const u64, gamma: *const u64, delta: *const u64, deltaomega: *const u64, omega: *const u64,
                                       last_z: *const u64, d_z: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_lookup_permute_dev(device: c_int, d_input: *const c_void, d_table: *const c_void, usable_rows: u32, d_permuted_input: *mut c_void,
                                  d_permuted_table: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_lookup_permute_async_dev(device: c_int, d_input: *const c_void, d_table: *const c_void, usable_rows: u32, d_permuted_input: *mut c_void,
                                        d_permuted_table: *mut c_void, d_status: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_lookup_product_dev(device: c_int, d_compressed_input: *const c_void, d_compressed_table: *const c_void, d_permuted_input: *const c_void,
                                  d_permuted_table: *const c_void, n: usize, beta: *const u64, gamma: *const u64, d_z: *mut c_void, stream: *mut c_void) -> c_int;
    // plonk::evaluation (evaluate_h)
    pub fn h2b_evaluate_graph_dev(device: c_int, graph: *const h2b_graph, cols: *const h2b_eval_columns, d_values: *mut c_void, size: u32, rot_scale: i32,
                                  stream: *mut c_void) -> c_int;
    pub fn h2b_evaluate_h_permutation_dev(device: c_int, d_values: *mut c_void, size: u32, rot_scale: i32, d_product_cosets: *const *const c_void, n_sets: u32,
                                          d_columns: *const *const c_void, d_perm_cosets: *const *const c_void, n_columns: u32, chunk_len: u32,
                                          last_rotation: i32, d_l0: *const c_void, d_l_last: *const c_void, d_l_active_row: *const c_void, beta: *const u64,
                                          gamma: *const u64, y: *const u64, delta: *const u64, zeta: *const u64, extended_omega: *const u64,
                                          stream: *mut c_void) -> c_int;
    pub fn h2b_evaluate_h_lookup_dev(device: c_int, graph: *const h2b_graph, cols: *const h2b_eval_columns, d_values: *mut c_void, size: u32, rot_scale: i32,
                                     d_product_coset: *const c_void, d_permuted_input_coset: *const c_void, d_permuted_table_coset: *const c_void,
                                     d_l0: *const c_void, d_l_last: *const c_void, d_l_active_row: *const c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_evaluate_graph_shard_dev(device: c_int, graph: *const h2b_graph, cols: *const h2b_eval_columns, d_values: *mut c_void, size: u32, rot_scale: i32,
                                        shard: *const h2b_eval_shard, stream: *mut c_void) -> c_int;
    pub fn h2b_evaluate_h_permutation_shard_dev(device: c_int, d_values: *mut c_void, size: u32, rot_scale: i32, d_product_cosets: *const *const c_void,
                                                n_sets: u32, d_columns: *const *const c_void, d_perm_cosets: *const *const c_void, n_columns: u32, chunk_len: u32,
                                                last_rotation: i32, d_l0: *const c_void, d_l_last: *const c_void, d_l_active_row: *const c_void,
                                                beta: *const u64, gamma: *const u64, y: *const u64, delta: *const u64, zeta: *const u64,
                                                extended_omega: *const u64, shard: *const h2b_eval_shard, stream: *mut c_void) -> c_int;
    pub fn h2b_evaluate_h_lookup_shard_dev(device: c_int, graph: *const h2b_graph, cols: *const h2b_eval_columns, d_values: *mut c_void, size: u32,
                                           rot_scale: i32, d_product_coset: *const c_void, d_permuted_input_coset: *const c_void,
                                           d_permuted_table_coset: *const c_void, d_l0: *const c_void, d_l_last: *const c_void,
                                           d_l_active_row: *const c_void, shard: *const h2b_eval_shard, stream: *mut c_void) -> c_int;
    // SRS file -> resident base sets (ParamsKZG::read_custom / write_custom)
    pub fn h2b_g1_decode_dev(device: c_int, d_bytes: *const c_void, n: usize, format: c_int, d_out_affine: *mut c_void, first_invalid: *mut u64, stream: *mut c_void) -> c_int;
    pub fn h2b_g1_encode_dev(device: c_int, d_affine: *const c_void, n: usize, d_out_bytes: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_srs_read(path: *const c_char, format: c_int, k: *mut u32, out_g: *mut u64, out_g_lagrange: *mut u64, g2_bytes: *mut u8, g2_cap: usize,
                        g2_len: *mut usize, handle_g: *mut u64, handle_g_lagrange: *mut u64) -> c_int;
    pub fn h2b_srs_write(path: *const c_char, format: c_int, k: u32, g: *const u64, g_lagrange: *const u64, g2_bytes: *const u8, g2_len: usize) -> c_int;
    pub fn h2b_srs_cache_clear() -> c_int;
    pub fn h2b_dev_alloc(device: c_int, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn h2b_dev_free(device: c_int, p: *mut c_void) -> c_int;
    pub fn h2b_memcpy_h2d(device: c_int, d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn h2b_memcpy_h2d_async(device: c_int, d_dst: *mut c_void, h_src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_memcpy_d2h(device: c_int, h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn h2b_dev_sync(device: c_int) -> c_int;
    // round 2: sharded residency, batched columns / polynomials, one-upload column pipeline, column copies
    pub fn h2b_register_bases_sharded(bases: *const u64, n: usize, handle: *mut u64) -> c_int;
    pub fn h2b_msm_bn254_g1_batch_registered(scalars: *const *const u64, lens: *const usize, count: usize, handle: u64, out_jac: *mut u64) -> c_int;
    pub fn h2b_ntt_bn254_fr_batch(a: *const *mut u64, count: usize, omega: *const u64, log_n: u32) -> c_int;
    pub fn h2b_msm_bn254_g1_dev_batch_registered(device: c_int, d_scalars: *const *const c_void, lens: *const usize, count: usize, handle: u64,
                                                 d_out_blocks: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_ntt_bn254_fr_dev_batch(device: c_int, d_polys: *const *mut c_void, count: usize, omega: *const u64, log_n: u32, stream: *mut c_void) -> c_int;
    pub fn h2b_lagrange_to_coeff_dev_batch(device: c_int, d_cols: *const *mut c_void, count: usize, k: u32, omega_inv: *const u64, ifft_divisor: *const u64,
                                           stream: *mut c_void) -> c_int;
    pub fn h2b_coeff_to_extended_dev_batch(device: c_int, d_cols: *const *mut c_void, count: usize, k: u32, extended_k: u32, extended_omega: *const u64,
                                           zeta_powers: *const u64, stream: *mut c_void) -> c_int;
    pub fn h2b_column_pipeline(device: c_int, lagrange: *const u64, handle_g_lagrange: u64, k: u32, extended_k: u32, omega_inv: *const u64,
                               ifft_divisor: *const u64, extended_omega: *const u64, zeta_powers: *const u64, out_commitment_jac: *mut u64,
                               out_coeff: *mut u64, out_extended: *mut u64, d_extended: *mut *mut c_void) -> c_int;
    pub fn h2b_memcpy_d2d_async(device: c_int, d_dst: *mut c_void, d_src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_memset_zero_async(device: c_int, d_dst: *mut c_void, bytes: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_implicit_cache_stats(out: *mut u64) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(h2b_last_error()).to_string_lossy().into_owned() }
}

/// Initialise every visible B200 once; panics (like the upstream functions' asserts) when no device is usable --
/// there is no CPU fallback.
pub fn ensure_init() {
    use std::sync::Once;
    static INIT: Once = Once::new();
    INIT.call_once(|| {
        let rc = unsafe { h2b_init(0) };
        if rc != 0 {
            panic!("h2b200: initialisation failed ({rc}): {}", last_error());
        }
    });
}

/// `best_multiexp::<G1Affine>`: `scalars` = `&[Fr]` as `n x 4` u64, `bases` = `&[G1Affine]` as `n x 8` u64;
/// returns the Jacobian triple `x | y | z`.
pub fn msm_bn254_g1(scalars: &[[u64; 4]], bases: &[[u64; 8]]) -> [u64; 12] {
    assert_eq!(scalars.len(), bases.len());
    ensure_init();
    let mut out = [0u64; 12];
    let rc = unsafe { h2b_msm_bn254_g1(scalars.as_ptr() as *const u64, bases.as_ptr() as *const u64, scalars.len(), out.as_mut_ptr()) };
    if rc != 0 {
        panic!("h2b_msm_bn254_g1 failed ({rc}): {}", last_error());
    }
    out
}

/// `best_fft::<Fr>`: in place, natural order, no scaling.
pub fn ntt_bn254_fr(a: &mut [[u64; 4]], omega: &[u64; 4], log_n: u32) {
    assert_eq!(a.len(), 1usize << log_n);
    ensure_init();
    let rc = unsafe { h2b_ntt_bn254_fr(a.as_mut_ptr() as *mut u64, omega.as_ptr(), log_n) };
    if rc != 0 {
        panic!("h2b_ntt_bn254_fr failed ({rc}): {}", last_error());
    }
}
