//! The whole FFI layer: declarations of include/maxpro_b200.h and the two helpers every wrapper
//! uses.  Designs cross the boundary row-major (`flat[i * d + k] == design[i][k]`), buffers are
//! caller-owned, a non-zero status carries the reference's own panic text.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

pub const METRIC_MAXPRO: c_int = 0;
pub const METRIC_MAXIMIN: c_int = 1;

unsafe extern "C" {
    pub fn maxpro_last_error() -> *const c_char;
    pub fn maxpro_release_device_cache(bytes: *mut u64) -> c_int;
    /// Devices the build_lhd / anneal_lhd calls of this thread spread over (count == 0: every device of the box).
    pub fn maxpro_set_devices(devices: *const c_int, count: c_int) -> c_int;
    pub fn maxpro_generate_lhd(n: u64, d: u64, seed: u64, out: *mut f64) -> c_int;
    pub fn maxpro_criterion(design: *const f64, n: u64, d: u64, out: *mut f64) -> c_int;
    pub fn maxpro_maximin_criterion(design: *const f64, n: u64, d: u64, out: *mut f64) -> c_int;
    pub fn maxpro_l2_distance(a: *const f64, b: *const f64, d: u64, out: *mut f64) -> c_int;
    pub fn maxpro_build_lhd(n: u64, d: u64, iters: u64, metric: c_int, seed: u64,
                            out: *mut f64, score: *mut f64, index: *mut u64) -> c_int;
    pub fn maxpro_anneal_lhd(design: *const f64, n: u64, d: u64, iters: u64, t0: f64, cooling: f64,
                             metric: c_int, minimize: c_int, seed: u64, swap: c_int, n_chains: u64,
                             out: *mut f64, out_metric: *mut f64, out_chain: *mut u64) -> c_int;
    pub fn maxpro_order_design(design: *const f64, n: u64, d: u64, metric: c_int, minimize: c_int,
                               out: *mut f64, perm: *mut u64) -> c_int;
}

/// The reference fails through `assert!`; the engine returns a status and keeps the same text.
pub fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(maxpro_last_error()) }.to_string_lossy().into_owned();
        panic!("{msg}");
    }
}

pub fn flatten(design: &[Vec<f64>]) -> (Vec<f64>, u64, u64) {
    let n = design.len();
    let d = design.first().map_or(0, |row| row.len());
    let mut flat = Vec::with_capacity(n * d);
    for row in design {
        assert!(row.len() == d, "All points must have the same number of dimensions");
        flat.extend_from_slice(row);
    }
    (flat, n as u64, d as u64)
}

pub fn unflatten(flat: &[f64], n: u64, d: u64) -> Vec<Vec<f64>> {
    if d == 0 {
        return vec![Vec::new(); n as usize];
    }
    flat.chunks(d as usize).take(n as usize).map(|row| row.to_vec()).collect()
}
