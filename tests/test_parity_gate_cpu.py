"""The parity gate itself (oracle/parity.py), checked without a GPU: a NumPy float32 restatement of the kernel's
arithmetic against the float64 oracle must leave no unexplained bin, the fp32 error must stay inside the derived
tolerance, and kernels that apply a WRONG reassignment rule must be rejected."""
import numpy as np
import pytest

from oracle import parity as P
from oracle import ssq_oracle as O


def _cases():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(12000) * 30
    t = np.arange(12000) / 30000.0
    yield "noise 512/32", x, np.hanning(512), 512, 32, 30000.0, {}
    yield "tone+noise 512/32", 50 * np.sin(2 * np.pi * 1000 * t) + 1e-3 * rng.standard_normal(12000), np.hanning(512), 512, 32, 30000.0, {}
    yield "README sine 256/64", np.sin(2 * np.pi * 100 * np.arange(1000) / 1000.0), np.hanning(256), 256, 64, 1000.0, {}
    yield "noise 1024/256", x, np.hanning(1024), 1024, 256, 30000.0, {}
    yield "gamma bites", x[:3000], np.hanning(128), 128, 16, 250.0, dict(gamma=5.0)
    yield "random window, odd hop", x[:8000], 0.2 + rng.random(512), 512, 33, 1000.0, {}
    yield "odd n_fft", x[:3000], np.hanning(257)[1:-1].copy(), 255, 7, 1.0, {}


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_fp32_restatement_passes_the_gate(case):
    _, x, win, n_fft, hop, fs, kw = case
    _, _, ao = O.ssq_stft(x, win, n_fft=n_fft, hop_len=hop, fs=fs, return_aux=True, **kw)
    kb, binf = P.emulate_fp32_kernel(x, win, n_fft, hop, fs, **kw)
    dw = 0.5 * fs / (n_fft // 2)
    rep = P.classify_stft_bins(kb, ao, n_fft, fs, kw.get("gamma"), w_dev=np.where(kb >= 0, binf * dw, np.inf))
    assert rep["unexplained"] == 0, P.public(rep)
    assert rep["max_err_over_tol"] < 0.6, P.public(rep)  # the derived bound holds with room, and is not vacuous
    # the device's Tx would equal the oracle's on every column without a flip
    assert rep["mismatch_total"] <= (rep["within_edge"] + rep["ill_conditioned"] + rep["gate_edge"]
                                     + rep["below_energy_gate"])


@pytest.mark.parametrize("rule", ["floor", "drop_out_of_range"])
def test_wrong_rules_are_rejected(rule):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(6000) * 3
    win = 0.2 + rng.random(256)
    _, _, ao = O.ssq_stft(x, win, n_fft=256, hop_len=16, fs=100.0, return_aux=True)
    kb, _ = P.emulate_fp32_kernel(x, win, 256, 16, 100.0, rule=rule)
    rep = P.classify_stft_bins(kb, ao, 256, 100.0)
    assert rep["unexplained"] > 0, (rule, P.public(rep))


def test_off_by_one_and_scatter_errors_are_rejected():
    rng = np.random.default_rng(2)
    x = rng.standard_normal(4000)
    win = np.hanning(128)
    _, _, ao = O.ssq_stft(x, win, n_fft=128, hop_len=8, fs=1.0, return_aux=True)
    kb, _ = P.emulate_fp32_kernel(x, win, 128, 8, 1.0)
    assert P.classify_stft_bins(kb, ao, 128, 1.0)["unexplained"] == 0
    bad = kb.copy()
    bad[10, 5] = min(bad[10, 5] + 1, 64) if bad[10, 5] != 64 else 63  # ONE bin off by one
    assert P.classify_stft_bins(bad, ao, 128, 1.0)["unexplained"] >= 1
    # reaccumulate: Tx follows the bins it is given
    T = P.reaccumulate(ao["Sx"], ao["k"], 1.0)
    To, _ = O.ssq_stft(x, win, n_fft=128, hop_len=8, fs=1.0)
    assert np.abs(T - To).max() < 1e-12 * np.abs(To).max()
