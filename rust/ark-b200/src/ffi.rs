//! Raw bindings of include/algebra_b200.h (what `bindgen` would emit for the entry points used here).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const B200_CURVE_BLS12_381: c_int = 0;
pub const B200_CURVE_BN254: c_int = 1;
pub const B200_FIELD_BLS12_381_FR: c_int = 0;
pub const B200_FIELD_BN254_FR: c_int = 1;

extern "C" {
    pub fn b200_last_error() -> *const c_char;
    pub fn b200_msm_sw_g1(curve: c_int, bases: *const u64, scalars: *const u64, n: usize, out_xyz: *mut u64) -> c_int;
    pub fn b200_msm_sw_g1_dev(curve: c_int, d_bases: *const c_void, d_scalars: *const c_void, n: usize,
                              out_xyz: *mut u64, stream: *mut c_void) -> c_int;
    pub fn b200_ntt_fr(field: c_int, data: *mut u64, log_n: u32, inverse: c_int, coset_offset: *const u64) -> c_int;
    pub fn b200_g1_sum(curve: c_int, points_xyz: *const u64, k: usize, out_xyz: *mut u64) -> c_int;
}

pub(crate) fn check(rc: c_int) {
    if rc != 0 {
        // fft_in_place returns () and SWCurveConfig::msm's Err is reserved for length mismatch, so a CUDA failure
        // cannot be propagated through the reference's trait signatures: panic (release profile is panic = abort).
        let msg = unsafe { core::ffi::CStr::from_ptr(b200_last_error()) };
        panic!("algebra_b200 failed with code {rc}: {}", msg.to_string_lossy());
    }
}
