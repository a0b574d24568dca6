//! `extern "C"` declarations of include/h2b200.h plus the two safe wrappers the `halo2_proofs` patches call.
//!
//! Layout contract (halo2curves 0.3.x): `Fr`/`Fq` are `#[repr(transparent)]`-compatible `[u64; 4]` in Montgomery form,
//! `G1Affine` is `x | y` (64 bytes, (0,0) = identity), `G1` is `x | y | z` Jacobian (96 bytes).  Slices of those types are
//! therefore passed by pointer cast, no conversion and no copy on the host.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

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
    pub fn h2b_dev_alloc(device: c_int, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn h2b_dev_free(device: c_int, p: *mut c_void) -> c_int;
    pub fn h2b_memcpy_h2d(device: c_int, d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn h2b_memcpy_d2h(device: c_int, h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn h2b_dev_sync(device: c_int) -> c_int;
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
