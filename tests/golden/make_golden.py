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


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "upstream_adm.npz"), **upstream_adm())
    np.savez_compressed(os.path.join(HERE, "upstream_odd.npz"), **upstream_cases())
    np.savez_compressed(os.path.join(HERE, "readme_cases.npz"), **readme_cases())
    for f in ("upstream_adm.npz", "upstream_odd.npz", "readme_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
