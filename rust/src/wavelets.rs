// wavelets.rs -- the wavelet generator functions src/ssqueeze/_rs.pyi:91-132 declares (reference:
// rust/src/wavelets/morlet.rs:59-146, gmw.rs:226-358; never registered there), over libssqcuda's host-side double
// implementations (ssq_wavelet_*).  `dtype` is accepted and unused, as in the reference.
use ndarray::Array1;
use num_complex::Complex64;
use numpy::{IntoPyArray, PyReadonlyArray1};
use pyo3::prelude::*;
use std::os::raw::c_int;

use crate::ffi;

fn bandpass(norm: &str) -> c_int {
    (norm.to_lowercase() == "bandpass") as c_int // gmw.rs:24, :44
}

fn run_morlet(py: Python<'_>, kind: c_int, w: Option<&[f64]>, n: usize, scale: f64, mu: f64) -> PyResult<PyObject> {
    let mut out = Array1::<Complex64>::zeros(n);
    let wp = w.map(|s| s.as_ptr()).unwrap_or(std::ptr::null());
    let st = unsafe { ffi::ssq_wavelet_morlet(kind, wp, n as i64, scale, mu, out.as_mut_ptr() as *mut f64) };
    ffi::check(st, std::ptr::null())?;
    Ok(out.into_pyarray(py).into_py(py))
}

#[allow(clippy::too_many_arguments)]
fn run_gmw(py: Python<'_>, kind: c_int, w: Option<&[f64]>, n: usize, scale: f64, gamma: f64, beta: f64, norm: &str,
           order: i32) -> PyResult<PyObject> {
    let mut out = Array1::<Complex64>::zeros(n);
    let wp = w.map(|s| s.as_ptr()).unwrap_or(std::ptr::null());
    let st = unsafe {
        ffi::ssq_wavelet_gmw(kind, wp, n as i64, scale, gamma, beta, bandpass(norm), order, out.as_mut_ptr() as *mut f64)
    };
    ffi::check(st, std::ptr::null())?; // SSQ_EINVAL -> ValueError: gamma <= 0, beta < 0, order < 0 (gmw.rs:238-246)
    Ok(out.into_pyarray(py).into_py(py))
}

#[pyfunction]
#[pyo3(signature = (w, mu=6.0, dtype="float64"))]
pub fn morlet(py: Python<'_>, w: PyReadonlyArray1<f64>, mu: f64, dtype: &str) -> PyResult<PyObject> {
    let _ = dtype;
    let v = w.as_array().to_vec();
    run_morlet(py, 0, Some(&v), v.len(), 1.0, mu)
}

#[pyfunction]
#[pyo3(signature = (n=1024, scale=1.0, mu=6.0, dtype="float64"))]
pub fn morlet_freq(py: Python<'_>, n: usize, scale: f64, mu: f64, dtype: &str) -> PyResult<PyObject> {
    let _ = dtype;
    run_morlet(py, 1, None, n, scale, mu)
}

#[pyfunction]
#[pyo3(signature = (n=1024, scale=1.0, mu=6.0, dtype="float64"))]
pub fn morlet_time(py: Python<'_>, n: usize, scale: f64, mu: f64, dtype: &str) -> PyResult<PyObject> {
    let _ = dtype;
    run_morlet(py, 2, None, n, scale, mu)
}

#[pyfunction]
#[pyo3(signature = (w, gamma=3.0, beta=60.0, norm="bandpass", order=0, dtype="float64"))]
pub fn gmw(py: Python<'_>, w: PyReadonlyArray1<f64>, gamma: f64, beta: f64, norm: &str, order: i32, dtype: &str)
    -> PyResult<PyObject> {
    let _ = dtype;
    let v = w.as_array().to_vec();
    run_gmw(py, 0, Some(&v), v.len(), 1.0, gamma, beta, norm, order)
}

#[pyfunction]
#[pyo3(signature = (n=1024, scale=1.0, gamma=3.0, beta=60.0, norm="bandpass", order=0, dtype="float64"))]
#[allow(clippy::too_many_arguments)]
pub fn gmw_freq(py: Python<'_>, n: usize, scale: f64, gamma: f64, beta: f64, norm: &str, order: i32, dtype: &str)
    -> PyResult<PyObject> {
    let _ = dtype;
    run_gmw(py, 1, None, n, scale, gamma, beta, norm, order)
}

#[pyfunction]
#[pyo3(signature = (n=1024, scale=1.0, gamma=3.0, beta=60.0, norm="bandpass", order=0, dtype="float64"))]
#[allow(clippy::too_many_arguments)]
pub fn gmw_time(py: Python<'_>, n: usize, scale: f64, gamma: f64, beta: f64, norm: &str, order: i32, dtype: &str)
    -> PyResult<PyObject> {
    let _ = dtype;
    run_gmw(py, 2, None, n, scale, gamma, beta, norm, order)
}

#[pyfunction]
#[pyo3(signature = (gamma=3.0, beta=60.0, kind="peak"))]
pub fn gmw_center_frequency(gamma: f64, beta: f64, kind: &str) -> PyResult<f64> {
    let k = match kind {
        "peak" => 0,
        "energy" => 1,
        _ => return Err(pyo3::exceptions::PyValueError::new_err(format!("Unknown center frequency kind: {}", kind))),
    };
    let mut out = 0.0f64;
    let st = unsafe { ffi::ssq_wavelet_gmw_center_frequency(gamma, beta, k, &mut out) };
    ffi::check(st, std::ptr::null())?;
    Ok(out)
}
