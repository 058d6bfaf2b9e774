use crate::ffi;

/// Euclidean distance of two points (src/maximin_utils.rs:26-45); host arithmetic inside the library.
pub fn calculate_l2_distance(a: &Vec<f64>, b: &Vec<f64>) -> f64 {
    assert!(a.len() == b.len(), "Points must have the same number of dimensions");
    assert!(!a.is_empty(), "Points must have at least one dimension");
    let mut out = 0.0f64;
    ffi::check(unsafe { ffi::maxpro_l2_distance(a.as_ptr(), b.as_ptr(), a.len() as u64, &mut out) });
    out
}

/// Minimum pairwise distance (src/maximin_utils.rs:54-67); a single point gives +inf.
pub fn maximin_criterion(design: &Vec<Vec<f64>>) -> f64 {
    assert!(!design.is_empty(), "Design must have at least one point");
    let (flat, n, d) = ffi::flatten(design);
    let mut out = 0.0f64;
    ffi::check(unsafe { ffi::maxpro_maximin_criterion(flat.as_ptr(), n, d, &mut out) });
    out
}

#[cfg(feature = "pyo3-bindings")]
#[pyo3::pyfunction(name = "maximin_criterion")]
pub fn py_maximin_criterion(design: Vec<Vec<f64>>) -> pyo3::PyResult<f64> {
    if design.is_empty() {
        return Err(pyo3::exceptions::PyValueError::new_err("Design cannot be empty"));
    }
    Ok(maximin_criterion(&design))
}
