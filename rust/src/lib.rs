//! `maxpro` on a B200: the reference crate's public surface (src/lib.rs, lhd.rs, maxpro_utils.rs,
//! maximin_utils.rs, anneal.rs, order.rs, enums.rs) over the C ABI of libmaxpro_b200.
//! There is no CPU implementation behind these functions: without an sm_100 device every call
//! panics with the library's error text.
pub mod anneal;
pub mod enums;
pub mod ffi;
pub mod lhd;
pub mod maximin_utils;
pub mod maxpro_utils;
pub mod order;

use std::os::raw::c_int;

/// Random search for a good Latin hypercube (src/lib.rs:30-100): candidate `i` is the design of
/// stream `seed + i`, the best by `metric` wins, ties go to the later candidate; `None` returns
/// the single design of stream `seed`.  Generation, scoring and the fold run in one fused kernel;
/// only the winner's index leaves the GPU and the winner is regenerated from it.
pub fn build_lhd(
    n_samples: u64,
    n_dim: u64,
    n_iterations: u64,
    metric: Option<enums::Metrics>,
    seed: u64,
) -> Vec<Vec<f64>> {
    assert!(n_samples > 0, "n_samples must be positive and nonzero");
    assert!(n_dim > 0, "n_dim must be positive and nonzero");
    assert!(n_iterations > 0, "n_iterations must be positive and nonzero");
    let code = match &metric {
        Some(m) => m.code(),
        None => return lhd::generate_lhd_seeded(n_samples, n_dim, seed),
    };
    let mut flat = vec![0.0f64; (n_samples * n_dim) as usize];
    ffi::check(unsafe {
        ffi::maxpro_build_lhd(n_samples, n_dim, n_iterations, code, seed, flat.as_mut_ptr(),
                              std::ptr::null_mut(), std::ptr::null_mut())
    });
    ffi::unflatten(&flat, n_samples, n_dim)
}

/// Devices the `build_lhd` / `anneal_lhd` calls of this thread spread over (empty slice = every device of the box):
/// the counterpart of rayon's global pool in the reference (src/lib.rs:65-98).  Not calling it keeps device 0.
pub fn set_devices(devices: &[i32]) {
    let ptr = if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() };
    ffi::check(unsafe { ffi::maxpro_set_devices(ptr, devices.len() as c_int) });
}

/// Which of the two criteria is this closure?  Decided on the reference's own known-answer
/// designs (src/maxpro_utils.rs:117-133, src/maximin_utils.rs:85-101):
/// [[0,1],[1,0]] -> MaxPro 1, maximin sqrt 2;  [[0,2],[1,0]] -> MaxPro 0.5, maximin sqrt 5.
pub(crate) fn metric_code_of<F: Fn(&Vec<Vec<f64>>) -> f64>(metric: &F) -> c_int {
    let a = metric(&vec![vec![0.0, 1.0], vec![1.0, 0.0]]);
    let b = metric(&vec![vec![0.0, 2.0], vec![1.0, 0.0]]);
    let near = |x: f64, y: f64| (x - y).abs() < 1e-9;
    if near(a, 1.0) && near(b, 0.5) {
        ffi::METRIC_MAXPRO
    } else if near(a, 2f64.sqrt()) && near(b, 5f64.sqrt()) {
        ffi::METRIC_MAXIMIN
    } else {
        panic!("metric must be maxpro_criterion or maximin_criterion: the GPU engine cannot run an arbitrary closure");
    }
}

#[cfg(feature = "pyo3-bindings")]
mod python {
    use super::*;
    use pyo3::exceptions::PyValueError;
    use pyo3::prelude::*;

    #[pyfunction(name = "build_lhd")]
    #[pyo3(signature = (n_samples, n_dim, n_iterations, metric, seed=None))]
    pub fn py_build_lhd(n_samples: u64, n_dim: u64, n_iterations: u64, metric: String, seed: Option<u64>)
                        -> PyResult<Vec<Vec<f64>>> {
        if n_samples == 0 {
            return Err(PyValueError::new_err("n_samples must be positive and nonzero"));
        }
        if n_dim == 0 {
            return Err(PyValueError::new_err("n_dim must be positive and nonzero"));
        }
        if n_iterations == 0 {
            return Err(PyValueError::new_err("n_iterations must be positive and nonzero"));
        }
        let m = match metric.as_str() {  // case-sensitive here, as in src/lib.rs:147-149
            "maxpro" => enums::Metrics::MaxPro,
            "maximin" => enums::Metrics::MaxiMin,
            _ => return Err(PyValueError::new_err("Metric must be 'maxpro' or 'maximin'.")),
        };
        Ok(build_lhd(n_samples, n_dim, n_iterations, Some(m), seed.unwrap_or_else(rand::random::<u64>)))
    }

    #[pymodule]
    fn maxpro(m: &Bound<'_, PyModule>) -> PyResult<()> {
        m.add_function(wrap_pyfunction!(py_build_lhd, m)?)?;
        m.add_function(wrap_pyfunction!(crate::lhd::py_generate_lhd, m)?)?;
        m.add_function(wrap_pyfunction!(crate::maxpro_utils::py_maxpro_criterion, m)?)?;
        m.add_function(wrap_pyfunction!(crate::maximin_utils::py_maximin_criterion, m)?)?;
        m.add_function(wrap_pyfunction!(crate::anneal::py_anneal_lhd, m)?)?;
        m.add_function(wrap_pyfunction!(crate::order::py_order_design, m)?)?;
        Ok(())
    }
}
