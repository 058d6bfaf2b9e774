use crate::{ffi, metric_code_of};

/// Simulated annealing of a design (src/anneal.rs:56-143): coordinate swaps (`swap`) or jitter,
/// Metropolis acceptance at a geometrically cooled temperature, the best design seen is returned.
///
/// `metric` keeps the reference's generic signature so existing call sites compile unchanged;
/// it must be one of this crate's two criteria (every call site of the reference passes one of
/// them).  It is recognised by evaluating it on two probe designs -- there is no CPU path that
/// could run an arbitrary closure, and none is wanted.
pub fn anneal_lhd<F>(
    design: &Vec<Vec<f64>>,
    n_iterations: u64,
    initial_temp: f64,
    cooling_rate: f64,
    metric: F,
    minimize: bool,
    seed: u64,
    swap: bool,
) -> Vec<Vec<f64>>
where
    F: Fn(&Vec<Vec<f64>>) -> f64,
{
    if design.is_empty() {
        return design.to_vec();
    }
    assert!(n_iterations > 0, "n_iterations must be positive and nonzero");
    assert!(initial_temp > 0.0, "initial_temp must be positive and nonzero");
    anneal_lhd_chains(design, n_iterations, initial_temp, cooling_rate, metric_code_of(&metric), minimize, seed, swap, 1)
}

/// Engine extension: `n_chains` independent chains (stream `seed + chain`) from the same start,
/// the best of them returned.  `n_chains == 1` is the reference's single chain.
pub fn anneal_lhd_chains(
    design: &Vec<Vec<f64>>,
    n_iterations: u64,
    initial_temp: f64,
    cooling_rate: f64,
    metric_code: std::os::raw::c_int,
    minimize: bool,
    seed: u64,
    swap: bool,
    n_chains: u64,
) -> Vec<Vec<f64>> {
    let (flat, n, d) = ffi::flatten(design);
    let mut out = vec![0.0f64; flat.len()];
    ffi::check(unsafe {
        ffi::maxpro_anneal_lhd(flat.as_ptr(), n, d, n_iterations, initial_temp, cooling_rate, metric_code,
                               minimize as i32, seed, swap as i32, n_chains, out.as_mut_ptr(),
                               std::ptr::null_mut(), std::ptr::null_mut())
    });
    ffi::unflatten(&out, n, d)
}

#[cfg(feature = "pyo3-bindings")]
#[pyo3::pyfunction(name = "anneal_lhd")]
#[pyo3(signature = (design, n_iterations, initial_temp, cooling_rate, metric_name, minimize, swap, seed=None))]
pub fn py_anneal_lhd(
    design: Vec<Vec<f64>>,
    n_iterations: u64,
    initial_temp: f64,
    cooling_rate: f64,
    metric_name: String,
    minimize: bool,
    swap: bool,
    seed: Option<u64>,
) -> pyo3::PyResult<Vec<Vec<f64>>> {
    use pyo3::exceptions::PyValueError;
    if n_iterations == 0 {
        return Err(PyValueError::new_err("n_iterations must be positive and nonzero"));
    }
    if !(initial_temp > 0.0) {
        return Err(PyValueError::new_err("initial_temp must be positive"));
    }
    let code = match metric_name.to_lowercase().as_str() {
        "maxpro" => ffi::METRIC_MAXPRO,
        "maximin" => ffi::METRIC_MAXIMIN,
        _ => {
            return Err(PyValueError::new_err(format!(
                "Unknown metric: '{}'. Available metrics are 'maxpro' and 'maximin'.", metric_name)));
        }
    };
    if design.is_empty() {
        return Ok(design);
    }
    let seed = seed.unwrap_or_else(rand::random::<u64>);
    Ok(anneal_lhd_chains(&design, n_iterations, initial_temp, cooling_rate, code, minimize, seed, swap, 1))
}
