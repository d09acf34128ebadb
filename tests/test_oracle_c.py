"""CPU: the C restatement (timed CPU baseline) against the NumPy oracle."""
import numpy as np
import pytest

from oracle import cref
from oracle import ssq_oracle as O


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n_fft,hop,N,fs", [(256, 64, 1000, 1000.0), (512, 32, 3000, 30000.0), (65, 3, 400, 1.0)])
def test_c_ref_matches_numpy_oracle(mode, n_fft, hop, N, fs):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N)
    win = np.hanning(n_fft + 2)[1:-1].copy()
    Tx, sf, Sx = cref.ssq_stft(x, win, n_fft, hop, fs, mode=mode, want_Sx=True)
    To, sfo, aux = O.ssq_stft(x, win, n_fft=n_fft, hop_len=hop, fs=fs, return_aux=True)
    assert np.allclose(sf, sfo, rtol=1e-15)
    assert np.abs(Sx - aux["Sx"]).max() < 1e-11 * np.abs(aux["Sx"]).max()
    bad = np.abs(Tx - To) > 1e-9 * np.abs(To).max()
    assert bad.mean() < 1e-4, bad.mean()  # f64 rounding can flip a bin that sits exactly on an edge


def test_c_ref_options():
    rng = np.random.default_rng(2)
    x = rng.standard_normal(700)
    win = np.hanning(128)
    for kw in (dict(padtype="zero"), dict(squeezing="lebesgue"), dict(gamma=3.0)):
        Tx, _ = cref.ssq_stft(x, win, 128, 16, 100.0, mode=0, **kw)
        To, _ = O.ssq_stft(x, win, n_fft=128, hop_len=16, fs=100.0, **kw)
        assert np.abs(Tx - To).max() < 1e-9 * np.abs(To).max()
    assert cref.num_threads() >= 1


@pytest.mark.parametrize("mode", [0, 1])
def test_c_ref_benchmark_geometry_vs_upstream_golden(mode):
    """The timed CPU baseline itself against upstream's output at n_fft=512, hop=32 (frames without padding)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "upstream_even512.npz"))
    j0, j1 = z["cols"]
    Tx, sf, Sx = cref.ssq_stft(z["x"], z["window"], 512, 32, 1.0, mode=mode, want_Sx=True)
    assert np.abs(Sx[:, j0:j1] - z["Sx"]).max() < 1e-11 * np.abs(z["Sx"]).max()
    bad = np.abs(Tx[:, j0:j1] - z["Tx"]) > 1e-9 * np.abs(z["Tx"]).max()
    assert bad.mean() < 1e-4, bad.mean()
