use crate::ffi;

/// psi(D) = (mean over i<j of 1/(prod_k (x_ik - x_jk)^2 + 1e-12))^(1/d), on the GPU.
/// Same edge cases as the reference (src/maxpro_utils.rs:72-83): empty design panics,
/// fewer than two points give 0.
pub fn maxpro_criterion(design: &Vec<Vec<f64>>) -> f64 {
    assert!(!design.is_empty(), "Design must have at least one point");
    let (flat, n, d) = ffi::flatten(design);
    let mut out = 0.0f64;
    ffi::check(unsafe { ffi::maxpro_criterion(flat.as_ptr(), n, d, &mut out) });
    out
}

#[cfg(feature = "pyo3-bindings")]
#[pyo3::pyfunction(name = "maxpro_criterion")]
pub fn py_maxpro_criterion(design: Vec<Vec<f64>>) -> pyo3::PyResult<f64> {
    if design.is_empty() {
        return Err(pyo3::exceptions::PyValueError::new_err("Design cannot be empty"));
    }
    Ok(maxpro_criterion(&design))
}
