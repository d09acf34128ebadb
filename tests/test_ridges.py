"""Ridge extraction (spec old/ssqueezepy/ridge_extraction.py:11-232).

CPU: the NumPy restatement (oracle/ridge_oracle.py) against upstream's own outputs (tests/golden/upstream_ridges.npz),
index for index.  GPU: the device's dynamic programme equals the oracle's bit for bit when both start from the same
E = -log(energy / max + eps) (every step is an IEEE add / multiply / min / compare), E itself agrees to rounding (one
libm `log`), and the end-to-end indices agree with upstream's on the golden maps."""
import os

import numpy as np
import pytest

from oracle import ridge_oracle as R

G = os.path.join(os.path.dirname(__file__), "golden")


def _cases():
    z = np.load(os.path.join(G, "upstream_ridges.npz"))
    for ci, (F, nr, bw, pen, tr) in enumerate(z["cases"]):
        p = f"r{ci}_"
        yield dict(Tf=z[p + "Tf"], scales=z[p + "scales"], idx=z[p + "idx"], ridge_f=z[p + "ridge_f"], ridge_e=z[p + "ridge_e"],
                   n_ridges=int(nr), bw=int(bw), penalty=float(pen), transform="cwt" if tr == 0 else "stft")


def test_oracle_equals_upstream_index_for_index():
    n = 0
    for c in _cases():
        idx, rf, re_ = R.extract_ridges(c["Tf"], c["scales"], penalty=c["penalty"], n_ridges=c["n_ridges"], bw=c["bw"],
                                        transform=c["transform"], get_params=True)
        assert np.array_equal(idx, c["idx"])
        assert np.array_equal(rf, c["ridge_f"]) and np.array_equal(re_, c["ridge_e"])
        n += 1
    assert n == 4


def test_oracle_python_slice_semantics_of_the_band_removal():
    """`energy[int(ridx - bw):int(ridx + bw), t] = 0`: a negative start counts from the end (usually an empty slice)."""
    e = np.ones((10, 3))
    R.zero_band(e, np.array([1, 5, 9]), 3)
    assert e[:, 0].sum() == 10          # [-2:4] -> [8:4]: empty
    assert np.array_equal(np.nonzero(e[:, 1] == 0)[0], np.arange(2, 8))
    assert np.array_equal(np.nonzero(e[:, 2] == 0)[0], np.arange(6, 10))


def _device_vs_oracle(Tf, scales, **kw):
    from ssqueeze_rs_b200 import _rs
    idx, rf, re_, E = _rs.extract_ridges(Tf, scales, get_params=True, return_E=True, **kw)
    dtype = np.float64 if Tf.dtype == np.complex128 else np.float32
    assert idx.shape == (Tf.shape[1], kw.get("n_ridges", 1)) and rf.dtype == dtype and E.dtype == dtype
    s_dev = R.device_coords(scales, kw.get("transform", "cwt"), dtype)
    # (1) same E, same coordinates -> the dynamic programme is bit-exact
    io, fo, eo = R.extract_ridges(Tf, scales, get_params=True, E_override=list(E), s_override=s_dev, **kw)
    assert np.array_equal(idx, io), int((idx != io).sum())
    assert np.array_equal(rf, fo)
    assert np.allclose(re_, eo, rtol=4 * np.finfo(dtype).eps)  # |Tf|^2: one hypot
    # (2) E itself: one log of a quotient
    st = R.stages(Tf, scales, **{k: v for k, v in kw.items()})
    for i, (Eo, _, _, _) in enumerate(st):
        fin = np.isfinite(Eo) & np.isfinite(E[i])
        assert np.array_equal(np.isfinite(Eo), np.isfinite(E[i])) or fin.mean() > 0.999
        assert np.abs(E[i][fin] - Eo[fin]).max() <= 16 * np.finfo(dtype).eps * max(1.0, np.abs(Eo[fin]).max())
        if i == 0:
            break  # later ridges depend on the removed bands, which depend on the indices
    return idx


@pytest.mark.gpu
def test_device_ridges_on_the_golden_maps():
    for c in _cases():
        kw = dict(penalty=c["penalty"], n_ridges=c["n_ridges"], bw=c["bw"], transform=c["transform"])
        idx = _device_vs_oracle(c["Tf"], c["scales"], **kw)
        # end to end against upstream's own indices (E differs by a rounding of log: the eps test of the backward pass
        # can flip on a few columns)
        agree = (idx == c["idx"]).mean()
        assert agree > 0.97, (agree, kw)


@pytest.mark.gpu
def test_device_ridges_readme_sine_and_chirp():
    """BASELINE configs[0] (1 s 100 Hz sine, ssq_stft 256/64) and a chirp through ssq_cwt: the ridge follows the
    component; device == oracle on the device's own Tx."""
    from ssqueeze_rs_b200 import _rs
    fs = 1000
    t = np.linspace(0, 1, fs, endpoint=False)
    x = np.sin(2 * np.pi * 100 * t)
    Tx, sf = _rs.ssq_stft(x, np.hanning(256), n_fft=256, hop_len=64, fs=float(fs))
    for cdt in (np.complex128, np.complex64):
        idx = _device_vs_oracle(Tx.astype(cdt), sf + 1e-9, penalty=2.0, n_ridges=1, bw=4, transform="stft")
        assert np.all(np.abs(sf[idx[:, 0]] - 100.0) <= 2 * (sf[1] - sf[0])), sf[idx[:, 0]]  # (the first frame is half padding)
    N = 4096
    tt = np.arange(N) / 1000.0
    xc = np.cos(2 * np.pi * (30 * tt + 0.5 * 60 * tt ** 2))
    Tc, sfc, = _rs.ssq_cwt(xc, "gmw", None, fs=1000.0, nv=16, maprange="maximal")
    sc = _rs.cwt(xc, "gmw", None, fs=1000.0, nv=16)[1]
    for cdt in (np.complex128, np.complex64):
        idx = _device_vs_oracle(Tc.astype(cdt), sc, penalty=2.0, n_ridges=2, bw=6, transform="cwt")
        assert idx.shape == (N, 2)
    # the first ridge climbs with the chirp (Tx rows are flipped: high frequencies first)
    r = idx[N // 8: -N // 8, 0].astype(float)
    assert np.corrcoef(r, np.arange(len(r)))[0, 1] < -0.9


@pytest.mark.gpu
def test_device_ridges_batched_tx_stays_on_device():
    """Engine: ssq_stft writes Tx into HBM, extract_ridges consumes it there; every channel equals the single-map call."""
    import torch
    from ssqueeze_rs_b200 import _rs
    from ssqueeze_rs_b200.batch import Engine
    eng = Engine(0)
    rng = np.random.default_rng(5)
    n, fs = 20000, 30000.0
    tt = np.arange(n) / fs
    x = np.stack([np.sin(2 * np.pi * (500 + 300 * c) * tt) + 0.3 * rng.standard_normal(n) for c in range(5)]).astype(np.float32)
    win = np.hanning(512)
    # modulated=True: the reference's unmodulated phase transform scatters a tone as soon as there is noise under it
    # (reference behaviour, restated faithfully); the modulated one keeps it on its bin
    Tx = eng.ssq_stft(torch.from_numpy(x).cuda(), win, 512, 32, fs, modulated=True)
    sf = np.arange(257) * 0.5 / 256 + 1e-9  # normalised frequencies, as upstream's stft `scales`
    idx, rf, re_ = eng.extract_ridges(Tx, sf, penalty=0.5, n_ridges=2, bw=3, transform="stft", get_params=True)
    torch.cuda.synchronize()
    assert idx.shape == (5, Tx.shape[2], 2) and idx.dtype == torch.int32
    Th = Tx.cpu().numpy()
    for c in (0, 4):
        i1, f1, e1 = _rs.extract_ridges(Th[c], sf, penalty=0.5, n_ridges=2, bw=3, transform="stft", get_params=True)
        assert np.array_equal(idx[c].cpu().numpy(), i1)
        assert np.array_equal(rf[c].cpu().numpy(), f1) and np.array_equal(re_[c].cpu().numpy(), e1)
        med = np.median(sf[i1[:, 0]]) * fs
        assert abs(med - (500 + 300 * c)) < 2 * (sf[1] - sf[0]) * fs, (c, med)
