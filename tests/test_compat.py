"""Upstream-compatible mode: ssqueeze_rs_b200.compat against upstream ssqueezepy's OWN outputs
(tests/golden/upstream_compat.npz, generated from old/ssqueezepy by tests/golden/make_golden.py): even and odd n_fft,
every frame including the padded ones, modulated and not, array / string / default (dpss) windows, fs != 1, the
inverses, and the thresholds of old/tests/reconstruction_test.py:160-206 re-hosted."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4
WK = {0: "array", 1: "hann", 2: None, 3: "hamming"}


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _cases():
    z = np.load(os.path.join(G, "upstream_compat.npz"))
    for ci, (N, n_fft, hop, mod, wk, fs) in enumerate(z["cases"]):
        p = f"k{ci}_"
        d = {k[len(p):]: z[k] for k in z.files if k.startswith(p)}
        n_fft = int(n_fft)
        window = np.hanning(n_fft + 2)[1:-1].copy() if WK[int(wk)] == "array" else WK[int(wk)]
        yield dict(d, N=int(N), n_fft=n_fft, hop=int(hop), mod=bool(mod), window=window, fs=float(fs))


def test_compat_windows_match_upstream():
    from ssqueeze_rs_b200 import compat
    for c in _cases():
        wl = len(c["window"]) if isinstance(c["window"], np.ndarray) else c["n_fft"]
        assert np.allclose(compat.get_window(c["window"], wl, c["n_fft"]), c["window_fit"], rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
def test_compat_against_upstream_outputs():
    from ssqueeze_rs_b200 import compat
    n = 0
    for c in _cases():
        kw = dict(window=c["window"], n_fft=c["n_fft"], hop_len=c["hop"], modulated=c["mod"])
        Tx, Sx, ssqf, Sfs, w, dSx = compat.ssq_stft(c["x"], fs=c["fs"], get_w=True, get_dWx=True, **kw)
        assert Tx.shape == c["Tx"].shape and np.allclose(ssqf, c["ssq_freqs"]) and np.allclose(Sfs, c["Sfs"])
        assert rel(Sx, c["Sx"]) < RTOL, ("Sx", kw, rel(Sx, c["Sx"]))          # every frame, padded ones included
        assert rel(dSx, c["dSx"]) < 2 * RTOL, ("dSx", kw, rel(dSx, c["dSx"]))
        To = c["Tx"]
        sc = np.abs(To).max()
        assert np.abs(Tx.sum(0) - To.sum(0)).max() < 20 * RTOL * sc, kw     # flip-invariant
        bad = np.abs(Tx - To) > RTOL * sc
        assert bad.mean() < 4e-3, (kw, float(bad.mean()))
        # w where the transform carries energy (upstream leaves w undefined / huge elsewhere)
        strong = np.abs(c["Sx"]) > 1e-3 * np.abs(c["Sx"]).max()
        fin = strong & np.isfinite(c["w"]) & np.isfinite(w)
        dw = Sfs[1] - Sfs[0]
        assert np.abs(w[fin] - np.abs(c["w"][fin])).max() < 0.05 * dw, kw
        # inverses
        xr = compat.istft(c["Sx"], N=c["N"], **kw)
        assert np.abs(xr - c["istft"]).max() < RTOL * max(1.0, np.abs(c["istft"]).max()), kw
        assert np.abs(xr - c["x"]).mean() < 5e-6, kw
        if "issq" in c:
            y = compat.issq_stft(c["Tx"], window=c["window"], n_fft=c["n_fft"])
            assert np.abs(y - c["issq"]).max() < RTOL * np.abs(c["issq"]).max(), kw
        n += 1
    assert n == 7


@pytest.mark.gpu
def test_upstream_reconstruction_tests_rehosted():
    """old/tests/reconstruction_test.py:160-206 (`test_stft`, `test_ssq_stft`) with the CUDA path behind upstream's
    signatures; the stft threshold is fp32's (5e-6 instead of 1e-14), the ssq_stft one is upstream's (0.1)."""
    from ssqueeze_rs_b200.compat import get_window, issq_stft, istft, ssq_stft, stft
    rng = np.random.default_rng(0)
    for N in (128, 129):
        x = rng.standard_normal(N)
        for n_fft in (120, 121):
            for hop_len in (1, 2, 3):
                for modulated in (True, False):
                    kw = dict(hop_len=hop_len, n_fft=n_fft, modulated=modulated)
                    Sx = stft(x, dtype="float64", **kw)
                    xr = istft(Sx, N=len(x), **kw)
                    assert len(x) == len(xr)
                    assert np.abs(x - xr).mean() < 5e-6, (N, n_fft, hop_len, modulated)
    for N in (128, 129):
        x = rng.standard_normal(N)
        for n_fft in (120, 121):
            for window_scaling in (1.0, 0.5):
                window = None if window_scaling == 1 else get_window(None, win_len=n_fft, n_fft=n_fft) * window_scaling
                Tx, *_ = ssq_stft(x, window=window, n_fft=n_fft)
                xr = issq_stft(Tx, window=window, n_fft=n_fft)
                assert len(x) == len(xr)
                assert np.abs(x - xr).mean() < 1e-1, (N, n_fft, window_scaling)
