// lib.rs -- the `ssqueeze._rs` extension module (reference: rust/src/lib.rs:16-35) with the compute bodies replaced
// by calls into libssqcuda.  Same callables, argument order, defaults and return arity as the reference registers
// (stft, ssq_stft, cwt, cwt_simd, ssq_cwt, hello_from_bin); plus what the north star and src/ssqueeze/_rs.pyi name
// and the reference never registers: istft, issq_stft, icwt, issq_cwt, adm_ssq, extract_ridges and the wavelet
// generator functions.
use pyo3::prelude::*;

mod ffi;
mod spectral;
mod wavelets;

#[pyfunction]
fn hello_from_bin() -> String {
    // lib.rs:16-19, verbatim through the library
    unsafe { std::ffi::CStr::from_ptr(ffi::ssq_hello_from_bin()) }.to_string_lossy().into_owned()
}

#[pymodule]
fn _rs(py: Python<'_>, m: &Bound<'_, PyModule>) -> PyResult<()> {
    m.add_function(wrap_pyfunction!(hello_from_bin, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::stft, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::ssq_stft, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::ssq_stft_batch, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::stft_batch, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::cwt, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::cwt_simd, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::ssq_cwt, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::istft, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::issq_stft, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::icwt, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::issq_cwt, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::adm_ssq, py)?)?;
    m.add_function(wrap_pyfunction!(spectral::extract_ridges, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::morlet, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::morlet_freq, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::morlet_time, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::gmw, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::gmw_freq, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::gmw_time, py)?)?;
    m.add_function(wrap_pyfunction!(wavelets::gmw_center_frequency, py)?)?;
    Ok(())
}
