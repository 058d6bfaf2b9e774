use crate::{ffi, metric_code_of};

/// Greedy run order (src/order.rs:34-105): the point nearest the centre first, then repeatedly
/// the remaining point that gives the best criterion of the ordered prefix.  The engine keeps a
/// running per-point sum / minimum instead of re-scoring every prefix (O(n^2 d) for the
/// reference's O(n^4 d)) and reproduces its `swap_remove` pool order and first-position ties.
pub fn order_design<F>(lhd: Vec<Vec<f64>>, metric: F, minimize: bool) -> Vec<Vec<f64>>
where
    F: Fn(&Vec<Vec<f64>>) -> f64,
{
    if lhd.is_empty() || lhd[0].is_empty() {
        return lhd;
    }
    order_design_code(lhd, metric_code_of(&metric), minimize)
}

pub fn order_design_code(lhd: Vec<Vec<f64>>, metric_code: std::os::raw::c_int, minimize: bool) -> Vec<Vec<f64>> {
    let (flat, n, d) = ffi::flatten(&lhd);
    let mut out = vec![0.0f64; flat.len()];
    ffi::check(unsafe {
        ffi::maxpro_order_design(flat.as_ptr(), n, d, metric_code, minimize as i32, out.as_mut_ptr(), std::ptr::null_mut())
    });
    ffi::unflatten(&out, n, d)
}

#[cfg(feature = "pyo3-bindings")]
#[pyo3::pyfunction(name = "order_design")]
pub fn py_order_design(design: Vec<Vec<f64>>, metric_name: String) -> pyo3::PyResult<Vec<Vec<f64>>> {
    if design.is_empty() {
        return Ok(design);
    }
    let (code, minimize) = match metric_name.to_lowercase().as_str() {
        "maxpro" => (ffi::METRIC_MAXPRO, true),
        "maximin" => (ffi::METRIC_MAXIMIN, false),
        _ => {
            return Err(pyo3::exceptions::PyValueError::new_err(format!(
                "Unknown metric: '{}'. Available metrics are 'maxpro' and 'maximin'.", metric_name)));
        }
    };
    Ok(order_design_code(design, code, minimize))
}
