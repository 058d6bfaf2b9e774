use crate::ffi;
use rand::Rng;
use rand::rngs::StdRng;

/// One random Latin hypercube (src/lhd.rs:97-132).  The reference threads a `StdRng`; the engine
/// generates from a 64-bit stream id (LHD-Philox v1), so one `u64` is drawn
/// from the caller's generator and used as that id: same state in, same design out.
pub fn generate_lhd(n_samples: u64, n_dim: u64, rng: &mut StdRng) -> Vec<Vec<f64>> {
    generate_lhd_seeded(n_samples, n_dim, rng.random::<u64>())
}

/// The engine's native form: the design of stream `seed` (what `build_lhd(.., seed)[i]` uses
/// with `seed + i`).
pub fn generate_lhd_seeded(n_samples: u64, n_dim: u64, seed: u64) -> Vec<Vec<f64>> {
    if n_samples == 0 {
        return Vec::new();
    }
    assert!(n_dim > 0, "Must have finite, positive number of dimensions");
    let mut flat = vec![0.0f64; (n_samples * n_dim) as usize];
    ffi::check(unsafe { ffi::maxpro_generate_lhd(n_samples, n_dim, seed, flat.as_mut_ptr()) });
    ffi::unflatten(&flat, n_samples, n_dim)
}

#[cfg(feature = "pyo3-bindings")]
#[pyo3::pyfunction(name = "generate_lhd")]
#[pyo3(signature = (n_samples, n_dim, seed=None))]
pub fn py_generate_lhd(n_samples: u64, n_dim: u64, seed: Option<u64>) -> pyo3::PyResult<Vec<Vec<f64>>> {
    use pyo3::exceptions::PyValueError;
    if n_samples == 0 {
        return Err(PyValueError::new_err("n_samples must be a positive, nonzero integer"));
    }
    if n_dim == 0 {
        return Err(PyValueError::new_err("n_dim must be a positive, nonzero integer"));
    }
    Ok(generate_lhd_seeded(n_samples, n_dim, seed.unwrap_or_else(rand::random::<u64>)))
}
