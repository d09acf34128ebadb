"""Regenerates the committed golden vectors.  Run in the BUILD container only:

    PYTHONPATH=/root/reference/old python tests/golden/make_golden.py

* `upstream_odd.npz`  -- outputs of the vendored upstream implementation
  (`/root/reference/old/ssqueezepy`, v0.6.6-dev) for the cases where it must
  coincide with the Rust path (odd n_fft so both pad conventions agree, fs=1,
  modulated=False, float64, explicit window of length n_fft).  These pin the
  oracle's STFT / dSTFT / phase / reassignment / istft against code that is
  part of the reference tree.
* `readme_cases.npz` -- the input recipes of the reference's smoke scripts
  (tests/stft_test.py:137-151, stft_ssq_test.py:132-152, cwt_test.py:19-57,
  ssq_cwt_test.py:19-57,410-419) with the ORACLE's outputs (no reference
  binary exists to produce them): a regression pin, not a reference pin.

* `upstream_adm.npz` -- upstream's synchrosqueezing admissibility constants
  (`utils/cwt_utils.py:28-47`) for gmw(3, 60) and morlet(6), with upstream's
  wavelet values at the points that fix the ratio to the Rust wavelets
  (cwt.rs:492-547 differ from upstream's by a constant factor each).

* `upstream_even512.npz` -- the BENCHMARK geometry (n_fft=512, hop=32, hann, fs=1): upstream frames the
  signal one sample earlier than the Rust code for even n_fft (pad n_fft/2 against (n_fft-1)/2,
  stft_utils.rs:19-49), so upstream run on x[1:] yields, on every frame that touches no padding, exactly
  the Rust frames of x.  Columns [j0, j1) of upstream's Sx, dSx and Tx are stored.
* `upstream_components.npz` -- upstream `_invert_components` (`_ssq_cwt.py:380-402`: the curve-band inversion of
  `issq_cwt`) on a random Tx with three bands (one with gaps, one random incl. out-of-range and -1 centres).
* `upstream_cwt.npz` -- upstream `cwt(..., derivative=True)` and `phase_cwt` for an even N whose padded
  length is the same in both code bases (N=600 -> 1024), explicit scales, l1 norm, reflect padding.
  The Rust wavelets equal upstream's up to one constant each (`*_ratio`), so Wx/dWx must agree to
  rounding for gmw; for morlet only where psi-hat is negligible at the Nyquist bin (the Rust grid
  keeps +pi there, `wavelets/base.rs:18-33`, upstream -pi) -- rows `morlet_rows_ok`.

* `upstream_compat.npz` -- upstream's stft / ssq_stft / istft / issq_stft with upstream's own conventions (left pad
  n_fft // 2, modulated, dpss / string / array windows), every frame: pins the upstream-compatible mode.
* `upstream_wavelets.npz` -- upstream's morlet / gmw (order 0, both norms) / centre frequencies / time-domain gmw
  at the points where the Rust generator functions (rust/src/wavelets/{morlet,gmw}.rs) coincide with them.
* `upstream_ridges.npz` -- upstream `extract_ridges(..., parallel=False, get_params=True)`
  (`ridge_extraction.py:11-232`; the sequential JIT code: the parallel one races on the ridge index) on a
  two-chirp CWT-like map and an STFT-like map, float64 and float32, 1-3 ridges: pins oracle/ridge_oracle.py.

Neither the GPU tests nor bench.py read /root/reference; they read these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ssq_oracle as O  # noqa: E402


def upstream_cases():
    import ssqueezepy as S  # from /root/reference/old
    out = {}
    rng = np.random.default_rng(20261018)
    cases = [(400, 65, 3), (300, 129, 1), (513, 33, 5)]
    out["cases"] = np.array(cases, dtype=np.int64)
    for ci, (N, n_fft, hop) in enumerate(cases):
        x = rng.standard_normal(N)
        win = np.hanning(n_fft + 2)[1:-1].copy()
        Tx, Sx, ssqf, Sfs, w, dSx = S.ssq_stft(
            x, window=win, n_fft=n_fft, hop_len=hop, fs=1.0, modulated=False,
            dtype="float64", get_w=True, get_dWx=True)
        xr = S.istft(Sx, window=win, n_fft=n_fft, hop_len=hop, N=N, modulated=False)
        p = f"c{ci}_"
        out[p + "x"] = x
        out[p + "window"] = win
        out[p + "Tx"] = Tx
        out[p + "Sx"] = Sx
        out[p + "dSx"] = dSx
        out[p + "ssq_freqs"] = np.asarray(ssqf)
        out[p + "istft"] = xr
        # self-check at generation time
        Tx_o, sf_o, aux = O.ssq_stft(x, win, n_fft=n_fft, hop_len=hop, fs=1.0, return_aux=True)
        assert np.abs(Tx_o - Tx).max() <= 1e-12 * np.abs(Tx).max(), "oracle != upstream"
    return out


def readme_cases():
    out = {}
    fs = 1000
    t = np.linspace(0, 1, fs, endpoint=False)
    x = np.sin(2 * np.pi * 100 * t)
    out["x"] = x
    n_fft, hop = 256, 64
    win = np.hanning(n_fft)
    Sx, freqs = O.stft(x, n_fft, hop, win, "reflect")
    out["stft_Sx"], out["stft_freqs"] = Sx, freqs
    out["istft_x"] = O.istft(Sx, win, n_fft=n_fft, hop_len=hop, N=len(x))
    Tx, sf = O.ssq_stft(x, win, n_fft=n_fft, hop_len=hop, fs=float(fs))
    out["ssq_stft_Tx"], out["ssq_stft_freqs"] = Tx, sf
    scales = np.logspace(1, 5, 32) / fs
    out["scales"] = scales
    for wav in ("gmw", "morlet"):
        Wx, sc, dWx = O.cwt(x, wav, scales, fs=float(fs), nv=16, derivative=True)
        rows = [0, 7, 16, 31]
        out[f"cwt_{wav}_rows"] = np.array(rows)
        out[f"cwt_{wav}_Wx_rows"] = Wx[rows]
        out[f"cwt_{wav}_dWx_rows"] = dWx[rows]
        out[f"cwt_{wav}_Wx_abs_sum"] = np.abs(Wx).sum(axis=1)
        Tq, sfq = O.ssq_cwt(x, wav, scales, fs=float(fs), nv=16)
        out[f"ssq_cwt_{wav}_Tx_abs_rowsum"] = np.abs(Tq).sum(axis=1)
        out[f"ssq_cwt_{wav}_freqs"] = sfq
    Tm, sfm = O.ssq_cwt(x, "gmw", None, fs=float(fs), nv=32, maprange="maximal", gamma=1e-6)
    out["ssq_cwt_maximal_Tx_abs_rowsum"] = np.abs(Tm).sum(axis=1)
    out["ssq_cwt_maximal_freqs"] = sfm
    return out


def upstream_adm():
    from ssqueezepy import Wavelet
    from ssqueezepy.utils.cwt_utils import adm_ssq
    out = {}
    for name, cfg, wpk in (("gmw", {"gamma": 3, "beta": 60}, 20.0 ** (1.0 / 3.0)), ("morlet", {"mu": 6}, 6.0)):
        wav = Wavelet((name, dict(cfg, dtype="float64")))
        out[f"{name}_css"] = np.float64(adm_ssq(wav))
        out[f"{name}_w"] = np.array([0.5 * wpk, wpk, 1.3 * wpk])
        out[f"{name}_psih"] = np.asarray(wav.fn(out[f"{name}_w"]), dtype=np.float64)
    return out


def upstream_even512():
    import ssqueezepy as S
    rng = np.random.default_rng(20261020)
    N, n_fft, hop = 2560, 512, 32
    t = np.arange(N)
    x = rng.standard_normal(N) + 3.0 * np.sin(2 * np.pi * (0.01 + 0.00002 * t) * t)
    win = np.hanning(n_fft)
    Tx, Sx, ssqf, Sfs, w, dSx = S.ssq_stft(x[1:], window=win, n_fft=n_fft, hop_len=hop, fs=1.0, modulated=False,
                                          dtype="float64", get_w=True, get_dWx=True)
    j0, j1 = (n_fft // 2) // hop + 1, (N - 1 - (n_fft // 2 + 2)) // hop
    out = dict(x=x, window=win, cols=np.array([j0, j1]), Tx=Tx[:, j0:j1], Sx=Sx[:, j0:j1], dSx=dSx[:, j0:j1],
               ssq_freqs=np.asarray(ssqf))
    Tx_o, sf_o, aux = O.ssq_stft(x, win, n_fft=n_fft, hop_len=hop, fs=1.0, return_aux=True)
    assert np.abs(Tx_o[:, j0:j1] - out["Tx"]).max() <= 1e-12 * np.abs(out["Tx"]).max(), "oracle != upstream"
    assert np.abs(aux["Sx"][:, j0:j1] - out["Sx"]).max() <= 1e-12 * np.abs(out["Sx"]).max()
    return out


def upstream_components():
    from ssqueezepy._ssq_cwt import _invert_components
    rng = np.random.default_rng(20261021)
    n_rows, n = 40, 300
    Tx = rng.standard_normal((n_rows, n)) + 1j * rng.standard_normal((n_rows, n))
    cc = np.stack([10 + np.round(3 * np.sin(np.arange(n) / 20)).astype(int), np.full(n, 28),
                   rng.integers(-1, n_rows + 3, n)], axis=1)
    cc[50:60, 0] = -1
    cw = np.stack([np.full(n, 3), np.full(n, 5), rng.integers(0, 4, n)], axis=1)
    x = _invert_components(Tx, cc.astype("int32"), cw.astype("int32"))
    assert np.abs(O.invert_components(Tx, cc, cw) - x).max() < 1e-12, "oracle != upstream"
    return dict(Tx=Tx, cc=cc.astype(np.int32), cw=cw.astype(np.int32), x=x)


def upstream_cwt():
    import ssqueezepy as S
    from ssqueezepy import Wavelet
    out = {}
    rng = np.random.default_rng(20261019)
    N = 600
    x = rng.standard_normal(N) + np.sin(2 * np.pi * 0.07 * np.arange(N))
    scales = 2.0 ** np.linspace(1, 6.5, 12)
    out["x"], out["scales"] = x, scales
    for name, cfg, wpk in (("gmw", {"gamma": 3, "beta": 60}, 20.0 ** (1.0 / 3.0)), ("morlet", {"mu": 6}, 6.0)):
        wav = Wavelet((name, dict(cfg, dtype="float64")))
        Wx, _, dWx = S.cwt(x, wav, scales=scales, fs=1.0, l1_norm=True, derivative=True, padtype="reflect")
        ratio = float(wav.fn(np.array([wpk]))[0] / O.generate_wavelet_fourier(np.array([wpk]), 1.0, name)[0].real)
        out[f"{name}_Wx"], out[f"{name}_dWx"], out[f"{name}_ratio"] = Wx, dWx, np.float64(ratio)
        gamma = 1e-3 * np.abs(Wx).max()
        out[f"{name}_gamma"] = np.float64(gamma)
        out[f"{name}_w"] = S.phase_cwt(Wx, dWx, difftype="trig", gamma=gamma)
        Wo, _, dWo = O.cwt(x, name, scales, fs=1.0, derivative=True)
        err = np.abs(Wx - ratio * Wo).max(axis=1) / np.abs(Wx).max()
        out[f"{name}_rows_ok"] = np.nonzero(err < 1e-7)[0]
        print(name, "rows agreeing with the oracle:", out[f"{name}_rows_ok"], "max err there", err[err < 1e-7].max())
    return out


def upstream_compat():
    """Upstream's own stft / ssq_stft / istft / issq_stft (old/ssqueezepy/_stft.py, _ssq_stft.py) with their default
    conventions -- even and odd n_fft, every frame including the padded ones, modulated True and False, array / string
    / default (dpss) windows: the pin of the upstream-compatible mode (ssqueeze_rs_b200/compat.py)."""
    import ssqueezepy as S
    rng = np.random.default_rng(20261020)
    out = {}
    cases = []
    specs = [(400, 128, 4, True, "array", 1.0), (400, 128, 4, False, "array", 1.0), (300, 120, 1, True, "hann", 1.0),
             (301, 121, 3, True, None, 1.0), (300, 64, 1, True, "array", 250.0), (1000, 512, 32, True, "array", 1.0),
             (257, 60, 2, False, "hamming", 1.0)]
    for ci, (N, n_fft, hop, mod, wkind, fs) in enumerate(specs):
        x = rng.standard_normal(N)
        window = np.hanning(n_fft + 2)[1:-1].copy() if wkind == "array" else wkind
        kw = dict(window=window, n_fft=n_fft, hop_len=hop, modulated=mod)
        Tx, Sx, ssqf, Sfs, w, dSx = S.ssq_stft(x, fs=fs, dtype="float64", get_w=True, get_dWx=True, **kw)
        xr = S.istft(Sx, N=N, **kw)
        p = f"k{ci}_"
        out[p + "x"] = x
        out[p + "Tx"], out[p + "Sx"], out[p + "dSx"], out[p + "w"] = (np.asarray(a) for a in (Tx, Sx, dSx, w))
        out[p + "ssq_freqs"], out[p + "Sfs"] = np.asarray(ssqf), np.asarray(Sfs)
        out[p + "istft"] = np.asarray(xr)
        out[p + "window_fit"] = np.asarray(S._stft.get_window(window, n_fft if not isinstance(window, np.ndarray) else len(window), n_fft, dtype="float64"))
        if hop == 1 and mod:
            out[p + "issq"] = np.asarray(S.issq_stft(Tx, window=window, n_fft=n_fft))
        cases.append((N, n_fft, hop, int(mod), {"array": 0, "hann": 1, None: 2, "hamming": 3}[wkind], fs))
    out["cases"] = np.array(cases, dtype=np.float64)
    return out


def upstream_wavelets():
    """Upstream's wavelets at the points where the Rust definitions (rust/src/wavelets/{morlet,gmw}.rs) coincide."""
    from ssqueezepy import _gmw
    from ssqueezepy.wavelets import Wavelet
    out = {}
    w = np.concatenate([np.linspace(-2.0, 12.0, 57), [0.0, 1e-3, 2.7144176165949063, 6.0]])
    out["w"] = w
    for mu in (6.0, 13.4, 5.0):
        out[f"morlet_mu{mu}"] = np.asarray(Wavelet(("morlet", {"mu": mu, "dtype": "float64"}))(w, nohalf=True)).astype(np.complex128).ravel()
    for (g, b) in ((3.0, 60.0), (3.0, 20.0), (2.0, 7.5)):
        for norm in ("bandpass", "energy"):
            f = _gmw.gmw(g, b, norm, 0, centered_scale=False, dtype="float64")
            wp = np.where(w > 0, w, 1.0)
            out[f"gmw_{g}_{b}_{norm}"] = np.where(w > 0, np.asarray(f(wp)), 0.0).astype(np.complex128)
        wm, we = _gmw.morsefreq(g, b, n_out=2)
        out[f"wc_peak_{g}_{b}"] = np.array(wm)
        out[f"wc_energy_{g}_{b}"] = np.array(we)
    # time domain: compute_gmw(time=True) follows the same recipe as gmw_time (gmw.rs:284-327)
    for n in (64, 101):
        X, x = _gmw.compute_gmw(n, 1.5, 3.0, 60.0, time=True, norm="bandpass", order=0, dtype="float64")
        out[f"gmw_time_{n}"] = np.asarray(x).astype(np.complex128)
        out[f"gmw_freq_{n}"] = np.asarray(X).astype(np.complex128)
    return out


def upstream_ridges():
    from ssqueezepy.ridge_extraction import extract_ridges
    rng = np.random.default_rng(20261019)
    out = {}
    cases = []
    T = 300
    t = np.arange(T)
    for ci, (F, transform, n_ridges, bw, penalty, cdt) in enumerate([
            (48, "cwt", 2, 4, 2.0, np.complex128), (48, "cwt", 1, 15, 20.0, np.complex64),
            (65, "stft", 3, 3, 0.5, np.complex128), (65, "stft", 2, 2, 5.0, np.complex64)]):
        f = np.arange(F)[:, None]
        c1 = 8 + 25 * t / T
        c2 = F - 10 - 12 * np.sin(2 * np.pi * t / T)
        Tf = (np.exp(-0.5 * ((f - c1[None]) / 1.2) ** 2) + 0.6 * np.exp(-0.5 * ((f - c2[None]) / 1.5) ** 2)
              + 0.05 * rng.standard_normal((F, T))) * np.exp(2j * np.pi * rng.random((F, T)))
        Tf = Tf.astype(cdt)
        scales = (2.0 ** np.linspace(1, 6, F)) if transform == "cwt" else np.linspace(0, 0.5, F)
        if transform == "stft":
            scales = scales + 1e-3  # log() is not taken for 'stft'; keep a generic grid
        idx, rf, re_ = extract_ridges(Tf, scales, penalty=penalty, n_ridges=n_ridges, bw=bw, transform=transform,
                                      get_params=True, parallel=False)
        p = f"r{ci}_"
        out[p + "Tf"] = Tf
        out[p + "scales"] = scales
        out[p + "idx"] = idx
        out[p + "ridge_f"] = rf
        out[p + "ridge_e"] = re_
        cases.append((F, n_ridges, bw, penalty, 0 if transform == "cwt" else 1))
    out["cases"] = np.array(cases, dtype=np.float64)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "compat":
        np.savez_compressed(os.path.join(HERE, "upstream_compat.npz"), **upstream_compat())
        print("upstream_compat.npz", os.path.getsize(os.path.join(HERE, "upstream_compat.npz")), "bytes")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "wavelets":
        np.savez_compressed(os.path.join(HERE, "upstream_wavelets.npz"), **upstream_wavelets())
        print("upstream_wavelets.npz", os.path.getsize(os.path.join(HERE, "upstream_wavelets.npz")), "bytes")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ridges":  # regenerate one file only
        np.savez_compressed(os.path.join(HERE, "upstream_ridges.npz"), **upstream_ridges())
        print("upstream_ridges.npz", os.path.getsize(os.path.join(HERE, "upstream_ridges.npz")), "bytes")
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, "upstream_ridges.npz"), **upstream_ridges())
    np.savez_compressed(os.path.join(HERE, "upstream_wavelets.npz"), **upstream_wavelets())
    np.savez_compressed(os.path.join(HERE, "upstream_compat.npz"), **upstream_compat())
    np.savez_compressed(os.path.join(HERE, "upstream_components.npz"), **upstream_components())
    np.savez_compressed(os.path.join(HERE, "upstream_even512.npz"), **upstream_even512())
    np.savez_compressed(os.path.join(HERE, "upstream_cwt.npz"), **upstream_cwt())
    np.savez_compressed(os.path.join(HERE, "upstream_adm.npz"), **upstream_adm())
    np.savez_compressed(os.path.join(HERE, "upstream_odd.npz"), **upstream_cases())
    np.savez_compressed(os.path.join(HERE, "readme_cases.npz"), **readme_cases())
    for f in ("upstream_components.npz", "upstream_even512.npz", "upstream_cwt.npz", "upstream_adm.npz", "upstream_odd.npz", "readme_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
