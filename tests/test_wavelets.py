"""Wavelet generator functions (src/ssqueeze/_rs.pyi:91-132, rust/src/wavelets/{morlet,gmw}.rs): host
arithmetic in libssqcuda (no device needed), against the NumPy restatement and against upstream's own values
(tests/golden/upstream_wavelets.npz) where the two definitions coincide."""
import os

import numpy as np
import pytest

from oracle import wavelets_oracle as W

G = os.path.join(os.path.dirname(__file__), "golden")


def _rs(built_lib):
    from ssqueeze_rs_b200 import _rs
    return _rs


def test_oracle_equals_upstream_where_definitions_coincide():
    z = np.load(os.path.join(G, "upstream_wavelets.npz"))
    w = z["w"]
    for mu in (6.0, 13.4, 5.0):
        assert np.abs(W.morlet(w, mu) - z[f"morlet_mu{mu}"]).max() < 1e-14
    for g, b in ((3.0, 60.0), (3.0, 20.0), (2.0, 7.5)):
        for norm in ("bandpass", "energy"):
            ref = z[f"gmw_{g}_{b}_{norm}"]
            assert np.abs(W.gmw(w, g, b, norm, 0) - ref).max() < 1e-13 * np.abs(ref).max()
        assert abs(W.gmw_center_frequency(g, b, "peak") - float(z[f"wc_peak_{g}_{b}"])) < 1e-15
        assert abs(W.gmw_center_frequency(g, b, "energy") - float(z[f"wc_energy_{g}_{b}"])) < 1e-13
    for n in (64, 101):
        assert np.abs(W.gmw_freq(n, 1.5) - z[f"gmw_freq_{n}"]).max() < 1e-13
        assert np.abs(W.gmw_time(n, 1.5) - z[f"gmw_time_{n}"]).max() < 1e-15


def test_library_functions_against_the_oracle(built_lib):
    rs = _rs(built_lib)
    rng = np.random.default_rng(3)
    w = np.concatenate([rng.uniform(-3, 15, 200), [0.0, 6.0, 2.7144176165949063]])
    for mu in (6.0, 13.4):
        a, b = rs.morlet(w, mu), W.morlet(w, mu)
        assert a.dtype == np.complex128 and np.abs(a - b).max() < 1e-14
    for norm in ("bandpass", "energy", "BandPass", "anything-else-is-L2"):
        for order in (0, 1, 2, 3):
            for g, be in ((3.0, 60.0), (2.0, 7.5)):
                a, b = rs.gmw(w, g, be, norm, order), W.gmw(w, g, be, norm, order)
                assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300), (norm, order, g, be)
    for n in (1, 2, 7, 64, 101, 1024):
        for scale in (1.0, 3.7):
            assert np.abs(rs.morlet_freq(n, scale) - W.morlet_freq(n, scale)).max() < 1e-14
            assert np.abs(rs.morlet_time(n, scale) - W.morlet_time(n, scale)).max() < 1e-14
            assert np.abs(rs.gmw_freq(n, scale, order=1) - W.gmw_freq(n, scale, order=1)).max() < 1e-12
            assert np.abs(rs.gmw_time(n, scale, norm="energy") - W.gmw_time(n, scale, norm="energy")).max() < 1e-13
    assert rs.gmw_time().shape == (1024,) and rs.morlet_freq().shape == (1024,)
    for kind in ("peak", "energy"):
        assert abs(rs.gmw_center_frequency(3.0, 60.0, kind) - W.gmw_center_frequency(3.0, 60.0, kind)) < 1e-14
    with pytest.raises(ValueError):
        rs.gmw_center_frequency(kind="median")
    for bad in (dict(gamma=0.0), dict(beta=-1.0), dict(order=-1)):
        with pytest.raises(ValueError):
            rs.gmw(w, **bad)
    with pytest.raises(TypeError):
        rs.morlet(w.astype(np.float32))


def test_library_against_upstream_goldens(built_lib):
    rs = _rs(built_lib)
    z = np.load(os.path.join(G, "upstream_wavelets.npz"))
    w = np.ascontiguousarray(z["w"])
    assert np.abs(rs.morlet(w, 6.0) - z["morlet_mu6.0"]).max() < 1e-14
    assert np.abs(rs.gmw(w) - z["gmw_3.0_60.0_bandpass"]).max() < 1e-13 * 2
    assert np.abs(rs.gmw(w, norm="energy") - z["gmw_3.0_60.0_energy"]).max() < 1e-12 * np.abs(z["gmw_3.0_60.0_energy"]).max()
    assert np.abs(rs.gmw_time(101, 1.5) - z["gmw_time_101"]).max() < 1e-15
