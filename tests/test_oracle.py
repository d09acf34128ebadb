"""CPU: the oracle against the committed golden vectors (tests/golden/*.npz,
made by tests/golden/make_golden.py from the vendored upstream implementation)
and against the reference's documented quirks."""
import os

import numpy as np
import pytest

from oracle import ssq_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_matches_upstream_golden():
    z = np.load(os.path.join(G, "upstream_odd.npz"))
    for ci, (N, n_fft, hop) in enumerate(z["cases"]):
        p = f"c{ci}_"
        x, win = z[p + "x"], z[p + "window"]
        Tx, sf, aux = O.ssq_stft(x, win, n_fft=int(n_fft), hop_len=int(hop), fs=1.0, return_aux=True)
        assert Tx.shape == z[p + "Tx"].shape
        sc = np.abs(z[p + "Tx"]).max()
        assert np.abs(Tx - z[p + "Tx"]).max() <= 1e-12 * sc
        assert np.abs(aux["Sx"] - z[p + "Sx"]).max() <= 1e-12 * np.abs(z[p + "Sx"]).max()
        assert np.abs(aux["dSx"] - z[p + "dSx"]).max() <= 1e-11 * np.abs(z[p + "dSx"]).max()
        assert np.allclose(sf, z[p + "ssq_freqs"], rtol=1e-13, atol=0)
        Sx, _ = O.stft(x, int(n_fft), int(hop), win, "reflect")
        xr = O.istft(Sx, win, n_fft=int(n_fft), hop_len=int(hop), N=int(N))
        assert np.abs(xr - z[p + "istft"]).max() < 1e-13
        assert np.abs(xr - x).mean() < 1e-14  # old/tests/reconstruction_test.py:165,179


def test_oracle_readme_regression():
    z = np.load(os.path.join(G, "readme_cases.npz"))
    x = z["x"]
    win = np.hanning(256)
    Sx, freqs = O.stft(x, 256, 64, win, "reflect")
    assert Sx.shape == (129, 16) and freqs.shape == (129,)  # tests/stft_test.py:137-151
    assert np.array_equal(Sx, z["stft_Sx"])
    Tx, sf = O.ssq_stft(x, win, n_fft=256, hop_len=64, fs=1000.0)
    assert Tx.shape == (129, 16)
    assert np.allclose(Tx, z["ssq_stft_Tx"], rtol=1e-12, atol=1e-12)
    Tq, sfq = O.ssq_cwt(x, "gmw", z["scales"], fs=1000.0, nv=16)
    assert Tq.shape == (32, 1000)  # tests/ssq_cwt_test.py:19-57
    assert np.allclose(np.abs(Tq).sum(axis=1), z["ssq_cwt_gmw_Tx_abs_rowsum"], rtol=1e-9)


def test_pad_conventions():
    x = np.arange(1.0, 11.0)
    p = O.pad_reflect_stft(x, 6)  # pad 5: left 2, right 3 (stft_utils.rs:21-23)
    assert list(p[:2]) == [3.0, 2.0] and list(p[-3:]) == [9.0, 8.0, 7.0]
    assert len(p) == 15
    # pad longer than the signal: remainder stays zero (guards :35, :43)
    q = O.pad_reflect_stft(np.array([1.0, 2.0, 3.0]), 12)
    assert len(q) == 14
    left = 5
    assert list(q[left:left + 3]) == [1.0, 2.0, 3.0]
    assert list(q[:left]) == [0.0, 0.0, 0.0, 3.0, 2.0]
    assert list(q[left + 3:]) == [2.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    z = O.pad_zeros_stft(x, 6)
    assert z[:2].sum() == 0 and z[-3:].sum() == 0


def test_diff_window_and_fit():
    w = np.hanning(64)
    dw = O.diff_window(w)
    # derivative of a smooth window: antisymmetric-ish, zero mean
    assert abs(dw.sum()) < 1e-12
    num = np.gradient(w)
    assert np.abs(dw[4:-4] - num[4:-4]).max() < 2e-3
    assert O.fit_window(np.ones(4), 8).tolist() == [0, 0, 1, 1, 1, 1, 0, 0]
    assert O.fit_window(np.arange(8.0), 4).tolist() == [2, 3, 4, 5]


def test_reassign_ties_and_clamp():
    # a w value exactly between two grid points goes to the LOWER index
    # (strict '<', ssq_stft.rs:285); out of range clamps; NaN -> bin 0
    sf = O.ssq_freqs_stft(5, 8.0)  # 0,1,2,3,4
    Sx = np.ones((5, 1), dtype=np.complex128)
    w = np.array([[1.5], [100.0], [np.nan], [np.inf], [2.49]])
    Tx, kk = O.reassign_stft(Sx, w, sf, "sum")
    assert kk[:, 0].tolist() == [1, 4, 0, -1, 2]
    assert Tx[:, 0].real.tolist() == [1.0, 1.0, 1.0, 0.0, 1.0]


def test_cwt_quirks():
    assert O.next_power_of_2(1500) == 2048 and O.next_power_of_2(1024) == 1024
    xi = O.xifn(1.0, 8)
    assert xi[4] > 0 and xi[5] < 0  # Nyquist positive (wavelets/base.rs:23-30)
    s = O.generate_log_scales(1000, 16)
    assert len(s) == int(np.ceil((np.log2(500.0) - 1) * 16)) and abs(s[0] - 2.0) < 1e-15
    s2 = O.generate_log_scales(1000, 16, simd=True)
    assert np.allclose(s, s2, rtol=1e-14)
    # default nv=32 grid has ratio 2^(1/32) < 1.1 -> binned LINEARLY (ssq_cwt.rs:135-139)
    x = np.sin(2 * np.pi * 100 * np.linspace(0, 1, 1000, endpoint=False))
    Tx, sf, aux = O.ssq_cwt(x, "gmw", None, fs=1000.0, nv=32, maprange="maximal", return_aux=True)
    assert Tx.shape == (len(aux["scales"]), 1000)
    assert sf[1] / sf[0] < 1.1
    Wx, sc, dWx = O.cwt(x, "morlet", np.logspace(1, 5, 32) / 1000, fs=1000.0, nv=16)
    assert Wx.shape == (32, 1000) and dWx is None  # 3-tuple always (cwt.rs:60,143)


def test_error_behaviour():
    x = np.zeros(10)
    with pytest.raises(ValueError):
        O.ssq_stft(x, np.ones(16), n_fft=8)  # win_len > n_fft (ssq_stft.rs:96-101)
    with pytest.raises(ValueError):
        O.cwt(x, t=np.array([0.0]))  # cwt.rs:68-70


def test_oracle_icwt_formula_and_errors():
    """cwt.rs:548-627 (one-integral icwt): constants, norms, x_len / x_mean handling."""
    rng = np.random.default_rng(0)
    Wx = rng.standard_normal((6, 50)) + 1j * rng.standard_normal((6, 50))
    sc = 2.0 ** np.arange(1, 7, dtype=np.float64)
    x = O.icwt(Wx, "gmw", sc, x_mean=0.5)
    assert np.allclose(x, 2.0 * np.log(2.0) * Wx.real.sum(0) + 0.5, rtol=1e-14)          # adm = 1, dj = ln 2
    xm = O.icwt(Wx, "morlet", sc)
    assert np.allclose(xm, (2.0 / 0.776) * np.log(2.0) * Wx.real.sum(0), rtol=1e-14)     # adm = 0.776
    x2 = O.icwt(Wx, "gmw", sc, l1_norm=False, x_len=20)
    assert x2.shape == (20,)
    assert np.allclose(x2, 2.0 * np.log(2.0) * (Wx.real[:, :20] / np.sqrt(sc)[:, None]).sum(0), rtol=1e-14)
    assert np.allclose(O.icwt(Wx, "gmw", sc[::-1].copy()), 2.0 * 0.1 * Wx.real.sum(0), rtol=1e-14)  # dj default 0.1
    with pytest.raises(ValueError):
        O.icwt(Wx)
    with pytest.raises(IndexError):
        O.icwt(Wx, "gmw", sc, x_len=51)


def test_oracle_admissibility_pinned_to_upstream():
    """adm_ssq of the Rust wavelets (cwt.rs:492-547) = upstream's adm_ssq (utils/cwt_utils.py:28-47) times the
    constant ratio between the two definitions of the same wavelet (golden: tests/golden/upstream_adm.npz)."""
    import math
    z = np.load(os.path.join(G, "upstream_adm.npz"))
    for wav in ("gmw", "morlet"):
        w = z[f"{wav}_w"]
        ratio = O.generate_wavelet_fourier(w, 1.0, wav).real / z[f"{wav}_psih"]
        assert np.allclose(ratio, ratio[1], rtol=1e-9), (wav, ratio)      # same wavelet up to a constant
        assert np.isclose(O.adm_ssq(wav), ratio[1] * float(z[f"{wav}_css"]), rtol=2e-6), wav
    assert np.isclose(O.adm_ssq("gmw"), (2.0 / 3.0) * math.gamma(20.0), rtol=1e-12)  # 2 int w^59 exp(-w^3) dw


def test_oracle_cwt_round_trips_with_exact_admissibility():
    """cwt -> icwt(exact_adm) and ssq_cwt(maximal) -> issq_cwt return x (old/tests/reconstruction_test.py style);
    with the reference's placeholder constants (cwt.rs:579-583) they do not."""
    N = 2048
    t = np.arange(N)
    x = np.cos(2 * np.pi * 0.05 * t) + 0.5 * np.cos(2 * np.pi * 0.11 * t + 1.0)
    for wav in ("gmw", "morlet"):
        Wx, sc, _ = O.cwt(x, wav, nv=32)
        xr = O.icwt(Wx, wav, sc, exact_adm=True)
        assert np.abs(xr - x).mean() < 5e-3, wav
        assert np.abs(O.icwt(Wx, wav, sc) - x).mean() > 0.2, wav
        Tx, _ = O.ssq_cwt(x, wav, nv=32, maprange="maximal")
        xs = O.issq_cwt(Tx, wav, sc)
        assert np.abs(xs - x).mean() < 5e-3, wav
    with pytest.raises(ValueError):
        O.issq_cwt(Tx)


def test_oracle_cwt_pinned_to_upstream():
    """cwt.rs:169-326 / ssq_cwt.rs:339-435 (padding, xi grid, psi-hat, inverse FFT normalisation, derivative,
    unpadding) and phase_cwt (ssq_cwt.rs:15-47) against upstream `cwt` / `phase_cwt` outputs
    (tests/golden/upstream_cwt.npz; the wavelets differ by one constant each, stored as `*_ratio`)."""
    z = np.load(os.path.join(G, "upstream_cwt.npz"))
    x, sc = z["x"], z["scales"]
    for wav, tol in (("gmw", 1e-13), ("morlet", 1e-7)):
        rows = z[f"{wav}_rows_ok"]
        assert len(rows) >= 10
        Wo, _, dWo = O.cwt(x, wav, sc, fs=1.0, derivative=True)
        r = float(z[f"{wav}_ratio"])
        Wu, dWu = z[f"{wav}_Wx"], z[f"{wav}_dWx"]
        assert np.abs(r * Wo[rows] - Wu[rows]).max() < tol * np.abs(Wu).max(), wav
        assert np.abs(r * dWo[rows] - dWu[rows]).max() < tol * np.abs(dWu).max(), wav
        # phase transform on upstream's own Wx, dWx: same formula, same gate
        w = O.phase_cwt(Wu, dWu, float(z[f"{wav}_gamma"]))
        wu = z[f"{wav}_w"]
        assert np.array_equal(np.isinf(w), np.isinf(wu)), wav
        fin = np.isfinite(wu)
        assert np.allclose(w[fin], wu[fin], rtol=1e-10, atol=1e-14), wav


def test_oracle_benchmark_geometry_pinned_to_upstream():
    """n_fft=512, hop=32 (BASELINE configs[1]): Sx, dSx and Tx of every frame that touches no padding equal
    upstream's on the one-sample-shifted signal (tests/golden/upstream_even512.npz)."""
    z = np.load(os.path.join(G, "upstream_even512.npz"))
    j0, j1 = z["cols"]
    Tx, sf, aux = O.ssq_stft(z["x"], z["window"], n_fft=512, hop_len=32, fs=1.0, return_aux=True)
    assert j1 - j0 >= 50
    for got, key in ((aux["Sx"], "Sx"), (aux["dSx"], "dSx"), (Tx, "Tx")):
        ref = z[key]
        assert np.abs(got[:, j0:j1] - ref).max() < 1e-12 * np.abs(ref).max(), key
    assert np.allclose(sf, z["ssq_freqs"], rtol=0, atol=1e-15)


def test_oracle_component_inversion_pinned_to_upstream():
    """invert_components (the curve-band branch of issq_cwt) against upstream `_invert_components`
    (old/ssqueezepy/_ssq_cwt.py:380-402; tests/golden/upstream_components.npz)."""
    z = np.load(os.path.join(G, "upstream_components.npz"))
    x = O.invert_components(z["Tx"], z["cc"], z["cw"])
    assert x.shape == z["x"].shape == (4, 300)
    assert np.abs(x - z["x"]).max() < 1e-12
    # bands cover disjoint or overlapping rows; band sums plus residual exceed the plain sum only by the overlaps
    sc = 2.0 ** np.linspace(1, 5, 40)
    full = O.issq_cwt(z["Tx"], "gmw", sc)
    comp = O.issq_cwt(z["Tx"], "gmw", sc, z["cc"], z["cw"])
    assert comp.shape == (4, 300) and np.all(np.isfinite(comp)) and full.shape == (300,)
