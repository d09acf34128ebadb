"""GPU parity of the CWT path (`cwt`, `cwt_simd`, `ssq_cwt`) against the float64
oracle (cwt.rs / cwt_simd.rs / ssq_cwt.rs restatement).  Tolerance: rtol 1e-4 of
the array maximum (fp32 arithmetic); reassignment flips at bin edges are counted
and bounded."""
import os

import numpy as np
import pytest

from oracle import parity as P
from oracle import ssq_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4


def _rs():
    from ssqueeze_rs_b200 import _rs
    return _rs


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def sine():
    fs = 1000
    t = np.linspace(0, 1, fs, endpoint=False)
    return np.sin(2 * np.pi * 100 * t), fs


def test_cwt_readme_shapes_and_values():
    """tests/cwt_test.py:19-57 and tests/cwt_simd_test.py:19-57 re-hosted."""
    rs = _rs()
    x, fs = sine()
    scales = np.logspace(1, 5, 32) / fs
    z = np.load(os.path.join(G, "readme_cases.npz"))
    for fn in (rs.cwt, rs.cwt_simd):
        out = fn(x, wavelet="gmw", scales=scales, fs=fs, nv=16, derivative=True)
        assert len(out) == 3  # always a 3-tuple (cwt.rs:60,143)
        Wx, sc, dWx = out
        assert Wx.shape == (32, 1000) and dWx.shape == (32, 1000) and Wx.dtype == np.complex128
        assert np.array_equal(sc, scales)
        Wo, _, dWo = O.cwt(x, "gmw", scales, fs=float(fs), nv=16, derivative=True)
        assert rel(Wx, Wo) < RTOL and rel(dWx, dWo) < RTOL
        rows = z["cwt_gmw_rows"]
        assert rel(Wx[rows], z["cwt_gmw_Wx_rows"]) < RTOL
    Wx, sc, dWx = rs.cwt(x, "morlet", scales, fs=fs)
    assert dWx is None
    Wo, _, _ = O.cwt(x, "morlet", scales, fs=float(fs))
    assert rel(Wx, Wo) < RTOL
    assert rel(Wx[z["cwt_morlet_rows"]], z["cwt_morlet_Wx_rows"]) < RTOL


@pytest.mark.parametrize("wavelet", ["gmw", "morlet", "unknown-falls-back-to-gmw"])
@pytest.mark.parametrize("N", [777, 1000, 3000, 20000])
def test_cwt_noise_default_scales(wavelet, N):
    rs = _rs()
    x = np.random.default_rng(N).standard_normal(N)
    nv = 8 if N < 5000 else 2
    Wx, sc, dWx = rs.cwt(x, wavelet, None, fs=250.0, nv=nv, derivative=True)
    Wo, so, dWo = O.cwt(x, wavelet, None, fs=250.0, nv=nv, derivative=True)
    assert np.allclose(sc, so, rtol=1e-14)
    assert Wx.shape == Wo.shape
    assert rel(Wx, Wo) < RTOL, rel(Wx, Wo)
    assert rel(dWx, dWo) < RTOL, rel(dWx, dWo)
    # cwt_simd: exp(p ln2) scale generator (cwt_simd.rs:489-527)
    W2, s2, _ = rs.cwt_simd(x, wavelet, None, fs=250.0, nv=nv)
    assert np.allclose(s2, O.generate_log_scales(N, nv, simd=True), rtol=1e-14)
    assert rel(W2, Wo) < RTOL


def test_cwt_options():
    rs = _rs()
    x = np.random.default_rng(1).standard_normal(1500)
    sc = np.array([2.0, 3.5, 8.0, 20.0, 77.0])
    for kw in (dict(l1_norm=False), dict(padtype="zero"), dict(rpadded=True), dict(t=np.arange(1500) * 0.004),
               dict(vectorized=False), dict(padtype="nonsense"), dict(rpadded=True, l1_norm=False, derivative=True)):
        Wx, s, dWx = rs.cwt(x, "gmw", sc, **kw)
        Wo, so, dWo = O.cwt(x, "gmw", sc, **kw)
        assert Wx.shape == Wo.shape, kw
        assert rel(Wx, Wo) < RTOL, kw
        if dWo is not None:
            assert rel(dWx, dWo) < RTOL, kw
    with pytest.raises(ValueError):
        rs.cwt(x, t=np.array([0.0]))
    with pytest.raises(TypeError):
        rs.cwt(x.astype(np.float32))


def check_ssq_cwt(x, wavelet, scales, fs, nv, fused=None, **kw):
    """Full parity report of one ssq_cwt call: the Tx row of every (scale, column) comes from the reassignment itself
    (`kb`); every difference from the oracle's rule (ssq_cwt.rs:165-209: round half away from zero, drop outside the
    grid, flipud) must be explained by the fp32 gate of oracle/parity.py, and Tx must be the reference's accumulation
    over exactly those rows.  fused: None = whatever the library picks, True/False = force."""
    rs = _rs()
    from ssqueeze_rs_b200 import _lib
    ctx = _lib.default_context()
    if fused is not None:
        ctx.set_option("no_cwt_fused", 0 if fused else 1)
    try:
        Tx, sf, aux = rs.ssq_cwt(x, wavelet, scales, fs=fs, nv=nv, return_aux=True, **kw)
        name = ctx.last_kernel_name()
    finally:
        ctx.set_option("no_cwt_fused", 0)
    To, sfo, ao = O.ssq_cwt(x, wavelet, scales, fs=fs, nv=nv, return_aux=True, **kw)
    assert Tx.shape == To.shape and np.allclose(sf, sfo, rtol=1e-13)
    wav = "morlet" if wavelet == "morlet" else "gmw"
    rep = P.classify_cwt_bins(aux["kb"], ao, x, wav, 1.0 / fs, kw.get("padtype", "reflect"), sfo,
                              flipud=kw.get("flipud", True), gamma=kw.get("gamma"), w_dev=aux["w"])
    pub = P.public(rep)
    pub["kernel"] = name
    assert rep["unexplained"] == 0, pub
    assert rep["max_err_over_tol"] <= 1.0, pub
    Tref = P.reaccumulate_cwt(ao["Wx"], aux["kb"], Tx.shape[0], kw.get("squeezing", "sum"))
    assert rel(Tx, Tref) < RTOL, ("Tx vs oracle Wx accumulated over the device's rows", rel(Tx, Tref), pub)
    good = ~(aux["kb"] != ao["k"]).any(axis=0)
    if good.any():
        assert rel(Tx[:, good], To[:, good]) < RTOL, pub
    pub["good_cols"] = int(good.sum())
    return pub


def test_ssq_cwt_readme_cases():
    """tests/ssq_cwt_test.py:19-57 (explicit scales, peak) and :410-419 (defaults, maximal)."""
    rs = _rs()
    x, fs = sine()
    scales = np.logspace(1, 5, 32) / fs
    z = np.load(os.path.join(G, "readme_cases.npz"))
    for wav in ("gmw", "morlet"):
        Tx, sf = rs.ssq_cwt(x, wavelet=wav, scales=scales, fs=fs, nv=16, padtype="reflect", squeezing="sum",
                            maprange="peak")
        assert Tx.shape == (32, 1000) and sf.shape == (32,)
        assert np.allclose(sf, z[f"ssq_cwt_{wav}_freqs"], rtol=1e-13)
        check_ssq_cwt(x, wav, scales, float(fs), 16)
    Tx, sf = rs.ssq_cwt(x, wavelet="gmw", scales=None, fs=fs, nv=32, squeezing="sum", maprange="maximal", gamma=1e-6)
    check_ssq_cwt(x, "gmw", None, float(fs), 32, maprange="maximal", gamma=1e-6)
    assert np.allclose(np.abs(Tx).sum(axis=1), z["ssq_cwt_maximal_Tx_abs_rowsum"], rtol=5e-3,
                       atol=2e-3 * z["ssq_cwt_maximal_Tx_abs_rowsum"].max())


@pytest.mark.parametrize("kw", [dict(), dict(maprange="maximal"), dict(ssq_freqs="linear", maprange="maximal"),
                                dict(squeezing="lebesgue", maprange="maximal"), dict(flipud=False, maprange="maximal"),
                                dict(padtype="zero", maprange="maximal"), dict(gamma=0.5, maprange="maximal"),
                                dict(wavelet="morlet", maprange="maximal")])
def test_ssq_cwt_chirp_noise(kw):
    rs = _rs()
    N = 4096
    t = np.arange(N) / 1000.0
    rng = np.random.default_rng(3)
    x = np.sin(2 * np.pi * (20 * t + 0.5 * 40 * t ** 2)) + 0.5 * rng.standard_normal(N)
    kw = dict(kw)
    wav = kw.pop("wavelet", "gmw")
    for fused in (True, False):  # L = 2^13: the stand-alone reassignment either way; larger sizes below
        check_ssq_cwt(x, wav, None, 1000.0, 8, fused=fused, **kw)


def test_ssq_cwt_multipass_fft_and_batch():
    """pad_len 2^15 (three FFT passes) and the batched device entry point."""
    import torch
    from ssqueeze_rs_b200 import _lib
    from ssqueeze_rs_b200.batch import Engine
    rs = _rs()
    N = 20000
    t = np.arange(N) / 1000.0
    rng = np.random.default_rng(4)
    x = np.sin(2 * np.pi * (5 * t + 0.5 * 20 * t ** 2)) + 0.5 * rng.standard_normal(N)
    sc = 2.0 ** np.linspace(1, 9, 24)
    check_ssq_cwt(x, "gmw", sc, 1000.0, 32, maprange="maximal")
    To, sfo = O.ssq_cwt(x, "gmw", sc, fs=1000.0, maprange="maximal")
    eng = Engine(0)
    xb = np.stack([x, x[::-1].copy(), 2 * x]).astype(np.float32)
    out = eng.ssq_cwt(torch.from_numpy(xb).cuda(), "gmw", sc, fs=1000.0, maprange="maximal")
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert out.shape == (3, 24, N)
    bad = np.abs(out[0].astype(np.complex128) - To) > RTOL * np.abs(To).max()
    assert bad.mean() < 5e-3  # (every flip is classified by check_ssq_cwt above: same kernels, float64 entry point)
    assert np.allclose(out[2], 2 * out[0], rtol=1e-5, atol=1e-5 * np.abs(out[0]).max())
    W = eng.cwt(torch.from_numpy(xb).cuda(), "gmw", sc, fs=1000.0)
    torch.cuda.synchronize()
    Wo, _, _ = O.cwt(xb[1].astype(np.float64), "gmw", sc, fs=1000.0)
    assert rel(W[1].cpu().numpy(), Wo) < RTOL


@pytest.mark.parametrize("N,nsc", [(9000, 40), (100000, 14), (250000, 9), (700000, 6)])
def test_ssq_cwt_fused_tail(N, nsc):
    """The fused last pass (W and dW rows of a scale in one CTA, phase transform and reassignment from registers,
    Wx / dWx never written): pad_len 2^14 (passes 7,7), 2^18 (7,4,7), 2^19 (7,5,7), 2^21 (7,7,7 -- BASELINE config 3's
    length), scales that skip zero, one and two broadcast passes, every option; each bin classified against the
    oracle, and the fused and stand-alone paths agree."""
    import torch
    from ssqueeze_rs_b200 import _lib
    from ssqueeze_rs_b200.batch import Engine
    rng = np.random.default_rng(N)
    fs = 1000.0
    t = np.arange(N) / fs
    x = np.sin(2 * np.pi * (3 * t + 0.5 * (200.0 / t[-1]) * t ** 2)) + 0.5 * rng.standard_normal(N)
    sc = 2.0 ** np.linspace(1, np.log2(N / 4), nsc)
    rep = check_ssq_cwt(x, "gmw", sc, fs, 32, fused=True, maprange="maximal")
    assert "fused" in rep["kernel"], rep
    if N <= 100000:
        for kw in (dict(), dict(maprange="maximal", ssq_freqs="linear"), dict(maprange="maximal", squeezing="lebesgue"),
                   dict(maprange="maximal", flipud=False), dict(maprange="maximal", padtype="zero"),
                   dict(maprange="maximal", gamma=0.5), dict(maprange="maximal", gamma=-1.0)):
            check_ssq_cwt(x, "gmw", sc, fs, 32, fused=True, **kw)
        check_ssq_cwt(x, "morlet", sc, fs, 32, fused=True, maprange="maximal")
        rep2 = check_ssq_cwt(x, "gmw", sc, fs, 32, fused=False, maprange="maximal")
        assert "reassign" in rep2["kernel"], rep2
    # batched device entry point: channels are independent, fused == stand-alone up to the order of additions
    eng = Engine(0)
    xb = torch.from_numpy(np.stack([x, x[::-1].copy(), 0.5 * x]).astype(np.float32)).cuda()
    a, sf, aux = eng.ssq_cwt(xb, "gmw", sc, fs=fs, maprange="maximal", return_aux=True)
    assert "fused" in eng.last_kernel_name()
    eng.ctx.set_option("no_cwt_fused", 1)
    try:
        b, _, aux_b = eng.ssq_cwt(xb, "gmw", sc, fs=fs, maprange="maximal", return_aux=True)
        assert "reassign" in eng.last_kernel_name()
    finally:
        eng.ctx.set_option("no_cwt_fused", 0)
    torch.cuda.synchronize()
    assert torch.equal(aux["kb"], aux_b["kb"])  # same W, dW bits -> same rows
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    assert float((a[2] - 0.5 * a[0]).abs().max()) <= 1e-5 * float(a.abs().max())


@pytest.mark.parametrize("N", [9000, 40000, 100000, 700000])
def test_cwt_fft_plans(N):
    """pad_len 2^14 (passes 7,7), 2^16 (7,5,4), 2^18 (7,4,7), 2^21 (7,7,7 -- BASELINE config 3's length)."""
    rs = _rs()
    rng = np.random.default_rng(N)
    t = np.arange(N) / 1000.0
    x = np.sin(2 * np.pi * (3 * t + 0.5 * 0.02 * t ** 2)) + 0.3 * rng.standard_normal(N)
    sc = np.array([2.0, 11.0, 97.0, 1500.0, 30000.0])  # the last three skip one or two broadcast passes
    for wav in ("gmw", "morlet"):
        Wx, _, dWx = rs.cwt(x, wav, sc, fs=1000.0, derivative=True)
        Wo, _, dWo = O.cwt(x, wav, sc, fs=1000.0, derivative=True)
        assert rel(Wx, Wo) < RTOL, (wav, rel(Wx, Wo))
        assert rel(dWx, dWo) < RTOL, (wav, rel(dWx, dWo))
    Wx, _, _ = rs.cwt(x, "gmw", sc, fs=1000.0, rpadded=True, padtype="zero")
    Wo, _, _ = O.cwt(x, "gmw", sc, fs=1000.0, rpadded=True, padtype="zero")
    assert rel(Wx, Wo) < RTOL


def test_icwt_one_integral():
    """cwt.rs:548-627 (one-integral branch) against the oracle, both norms, x_len, x_mean."""
    rs = _rs()
    from ssqueeze_rs_b200 import PanicException, SsqError
    rng = np.random.default_rng(8)
    x = rng.standard_normal(3000)
    sc = 2.0 ** np.linspace(1, 8, 40)
    for wav in ("gmw", "morlet"):
        for l1 in (True, False):
            Wx, _, _ = rs.cwt(x, wav, sc, fs=1.0, l1_norm=l1)
            xr = rs.icwt(Wx, wav, sc, l1_norm=l1, x_mean=0.25)
            xo = O.icwt(Wx, wav, sc, l1_norm=l1, x_mean=0.25)
            assert xr.shape == (3000,) and xr.dtype == np.float64
            assert np.abs(xr - xo).max() < RTOL * np.abs(xo).max()
    Wx, _, _ = rs.cwt(x, "gmw", sc, fs=1.0)
    xo = O.icwt(Wx, "gmw", sc, x_len=1000)
    xr = rs.icwt(Wx, "gmw", sc, x_len=1000)
    assert len(xr) == 1000 and np.abs(xr - xo).max() < RTOL * np.abs(xo).max()
    with pytest.raises(ValueError):
        rs.icwt(Wx)                                   # "Scales must be provided"
    with pytest.raises(PanicException):
        rs.icwt(Wx, "gmw", sc, x_len=5000)            # index out of bounds in the reference


def test_admissibility_icwt_exact_and_issq_cwt():
    """the true admissibility constant, icwt with it, and issq_cwt (spec
    old/ssqueezepy/_ssq_cwt.py:313-378) against the oracle, plus the reconstruction property."""
    rs = _rs()
    for wav in ("gmw", "morlet"):
        assert np.isclose(rs.adm_ssq(wav), O.adm_ssq(wav), rtol=1e-9), wav
    N = 4096
    t = np.arange(N)
    x = np.cos(2 * np.pi * 0.05 * t) + 0.5 * np.cos(2 * np.pi * 0.11 * t + 1.0)
    for wav in ("gmw", "morlet"):
        Wx, sc, _ = rs.cwt(x, wav, nv=32)
        xr = rs.icwt(Wx, wav, sc, exact_adm=True)
        xo = O.icwt(Wx, wav, sc, exact_adm=True)
        assert np.abs(xr - xo).max() < RTOL * np.abs(xo).max(), wav
        assert np.abs(xr - x).mean() < 5e-3, (wav, np.abs(xr - x).mean())
        Tx, _ = rs.ssq_cwt(x, wav, nv=32, maprange="maximal")
        xs = rs.issq_cwt(Tx, wav, sc)
        assert xs.shape == (N,) and xs.dtype == np.float64
        assert np.abs(xs - O.issq_cwt(Tx, wav, sc)).max() < RTOL * np.abs(x).max(), wav
        assert np.abs(xs - x).mean() < 5e-3, (wav, np.abs(xs - x).mean())
    with pytest.raises(ValueError):
        rs.issq_cwt(Tx)                               # "Scales must be provided"
    with pytest.raises(TypeError):
        rs.issq_cwt(Tx.real, "gmw", sc)


def test_cwt_against_upstream_golden():
    """The CUDA cwt against upstream ssqueezepy's own output (tests/golden/upstream_cwt.npz): Wx and dWx agree up
    to the constant ratio between the two wavelet definitions (all gmw rows; morlet rows with no energy at the
    Nyquist bin)."""
    rs = _rs()
    z = np.load(os.path.join(G, "upstream_cwt.npz"))
    x, sc = z["x"], z["scales"]
    for wav in ("gmw", "morlet"):
        rows = z[f"{wav}_rows_ok"]
        Wx, _, dWx = rs.cwt(x, wav, sc, fs=1.0, derivative=True)
        r = float(z[f"{wav}_ratio"])
        Wu, dWu = z[f"{wav}_Wx"], z[f"{wav}_dWx"]
        assert np.abs(r * Wx[rows] - Wu[rows]).max() < RTOL * np.abs(Wu).max(), wav
        assert np.abs(r * dWx[rows] - dWu[rows]).max() < RTOL * np.abs(dWu).max(), wav


def test_full_size_round_trips_config3():
    """BASELINE configs[2] size (2^20 samples, nv=32, 576 scales) on the device: cwt -> icwt and
    ssq_cwt(maximal) -> issq_cwt give the chirp back (size-independent property; the oracle would need > 80 GB
    in float64 at this size), and the squeezed transform keeps the column sums of Wx."""
    import torch
    from ssqueeze_rs_b200.batch import Engine
    eng = Engine(0)
    n = 1 << 20
    t = torch.arange(n, device="cuda", dtype=torch.float64) / n
    x = torch.sin(2 * np.pi * n * (0.002 * t + 0.5 * 0.2 * t * t)).to(torch.float32).view(1, -1).contiguous()
    sc = eng.default_scales(n, 32)
    assert len(sc) == 576
    Wx = eng.cwt(x, "gmw", sc, fs=1.0)
    xr = eng.icwt(Wx, sc, "gmw", exact_adm=True)
    core = slice(n // 16, n - n // 16)
    err = float((xr - x)[0, core].abs().mean())
    assert err < 2e-2, err
    colsum_W = Wx.sum(dim=1)
    del Wx
    torch.cuda.empty_cache()
    Tx = eng.ssq_cwt(x, "gmw", sc, fs=1.0, maprange="maximal")
    xs = eng.issq_cwt(Tx, sc, "gmw")
    err2 = float((xs - x)[0, core].abs().mean())
    assert err2 < 2e-2, err2
    # ssqueezing only moves coefficients along the scale axis (those whose bin falls outside the grid are dropped)
    colsum_T = Tx.sum(dim=1)
    rel_cs = float((colsum_T - colsum_W)[0, core].abs().mean() / colsum_W[0, core].abs().mean())
    assert rel_cs < 5e-2, rel_cs


def test_issq_cwt_components():
    """Curve-band inversion (old/ssqueezepy/_ssq_cwt.py:380-402) on the device against upstream's own output
    (up to the admissibility / log-step factor) and against the oracle."""
    rs = _rs()
    z = np.load(os.path.join(G, "upstream_components.npz"))
    Tx, cc, cw = z["Tx"], z["cc"], z["cw"]
    sc = 2.0 ** np.linspace(1, 5, Tx.shape[0])
    for wav in ("gmw", "morlet"):
        x = rs.issq_cwt(Tx, wav, sc, cc, cw)
        assert x.shape == (4, 300) and x.dtype == np.float64
        f = (2.0 / rs.adm_ssq(wav)) * np.log(sc[1] / sc[0])
        assert np.abs(x / f - z["x"]).max() < RTOL * np.abs(z["x"]).max(), wav
        assert np.abs(x - O.issq_cwt(Tx, wav, sc, cc, cw)).max() < RTOL * np.abs(x).max(), wav
    one = rs.issq_cwt(Tx, "gmw", sc, cc[:, 0], cw[:, 0])  # 1-D cc / cw: one band
    assert one.shape == (2, 300)
    assert np.abs(one[0] - rs.issq_cwt(Tx, "gmw", sc, cc, cw)[0]).max() == 0.0
    with pytest.raises(ValueError):
        rs.issq_cwt(Tx, "gmw", sc, cc)


def test_cwt_tiny_and_ragged_inputs():
    """Edge cases the reference accepts: N from 2 samples (pad_len 2 .. 64: single-pass FFT plans), odd lengths,
    explicit scales, both wavelets, the squeezed transform with every option -- against the oracle."""
    rs = _rs()
    rng = np.random.default_rng(99)
    sc = np.array([2.0, 4.0, 9.5])
    for N in (2, 3, 4, 5, 8, 16, 33, 127, 1025):
        x = rng.standard_normal(N)
        for wav in ("gmw", "morlet"):
            Wx, s, dWx = rs.cwt(x, wav, sc, fs=2.0, derivative=True)
            Wo, so, dWo = O.cwt(x, wav, sc, fs=2.0, derivative=True)
            assert Wx.shape == Wo.shape == (3, N), (N, wav)
            # (+1e-30: psi-hat below the fp32 range is zero on the device, 1e-60 in the float64 oracle)
            assert np.abs(Wx - Wo).max() <= RTOL * np.abs(Wo).max() + 1e-30, (N, wav)
            assert np.abs(dWx - dWo).max() <= RTOL * np.abs(dWo).max() + 1e-30, (N, wav)
        for kw in (dict(), dict(maprange="maximal"), dict(squeezing="lebesgue", maprange="maximal"),
                   dict(flipud=False, maprange="maximal", ssq_freqs="linear")):
            Tx, sf = rs.ssq_cwt(x, "gmw", sc, fs=2.0, **kw)
            To, sfo = O.ssq_cwt(x, "gmw", sc, fs=2.0, **kw)
            assert Tx.shape == To.shape == (3, N) and np.allclose(sf, sfo, rtol=1e-13), (N, kw)
            bad = np.abs(Tx - To) > RTOL * np.abs(To).max() + 1e-7
            assert bad.mean() <= 0.05, (N, kw, float(bad.mean()))
    if True:  # default scales on short signals: ns = ceil((log2(N/2) - 1) nv) (cwt.rs:461-489); 0 scales for N <= 4
        for N in (5, 8, 33):
            x = rng.standard_normal(N)
            Wx, s, _ = rs.cwt(x, "gmw", None, nv=4)
            Wo, so, _ = O.cwt(x, "gmw", None, nv=4)
            assert Wx.shape == Wo.shape and np.allclose(s, so, rtol=1e-13), N
            assert np.abs(Wx - Wo).max() <= RTOL * np.abs(Wo).max() + 1e-30, N


def test_icwt_two_integral_power_of_two():
    """cwt.rs:629-712 (two-integral branch) on rpadded coefficients (row length 2^k) against the oracle; other
    lengths are refused."""
    rs = _rs()
    from ssqueeze_rs_b200 import SsqError
    rng = np.random.default_rng(21)
    for N, sc in ((3000, 2.0 ** np.linspace(1, 8, 30)), (700, 2.0 ** np.linspace(1, 6, 12)), (70000, 2.0 ** np.linspace(1, 12, 40))):
        x = rng.standard_normal(N)
        for wav in ("gmw", "morlet"):
            for l1 in (True, False):
                Wx, _, _ = rs.cwt(x, wav, sc, fs=1.0, rpadded=True, l1_norm=l1)
                L = Wx.shape[1]
                assert L & (L - 1) == 0
                xr = rs.icwt(Wx, wav, sc, one_int=False, l1_norm=l1, x_mean=-0.5)
                xo = O.icwt(Wx, wav, sc, one_int=False, l1_norm=l1, x_mean=-0.5)
                assert xr.shape == (L,)
                assert np.abs(xr - xo).max() < 2 * RTOL * np.abs(xo + 0.5).max(), (N, wav, l1)
    # any row length (rustfft takes any): 3000, 777 and 4097 columns through Bluestein; x_len shorter than the rows
    for N, sc in ((3000, 2.0 ** np.linspace(1, 8, 30)), (777, 2.0 ** np.linspace(1, 6, 12)), (4097, 2.0 ** np.linspace(1, 9, 20))):
        x = rng.standard_normal(N)
        for wav in ("gmw", "morlet"):
            Wx, _, _ = rs.cwt(x, wav, sc, fs=1.0)
            for xl in (None, N // 3):
                xr = rs.icwt(Wx, wav, sc, one_int=False, x_mean=0.125, x_len=xl)
                from ssqueeze_rs_b200 import _lib
                assert "icwt2" in _lib.default_context().last_kernel_name()
                xo = O.icwt(Wx, wav, sc, one_int=False, x_mean=0.125, x_len=xl)
                assert xr.shape == xo.shape
                assert np.abs(xr - xo).max() < 2 * RTOL * np.abs(xo - 0.125).max(), (N, wav, xl)


def test_rs_ssq_cwt_batch_against_the_per_channel_drop_in():
    """`_rs.ssq_cwt_batch` (all channels in one call, complex64 in pinned host memory or on the device) against the
    scalar `_rs.ssq_cwt` loop: the same kernels; the fused tail adds contributions of different scales in no fixed
    order, so the comparison is to fp32 rounding of the map's maximum, and the column sums to 1e-6."""
    import torch
    from ssqueeze_rs_b200 import _rs
    rng = np.random.default_rng(8)
    n = 3000
    tt = np.arange(n) / 1000.0
    x = np.stack([np.sin(2 * np.pi * (20 + 10 * c) * tt) + 0.2 * rng.standard_normal(n) for c in range(3)])
    Tx, sf = _rs.ssq_cwt_batch(x, fs=1000.0, maprange="maximal", nv=16)
    assert Tx.dtype == np.complex64 and Tx.shape[0] == 3 and Tx.shape[2] == n
    for c in range(3):
        Tc, sfc = _rs.ssq_cwt(x[c], fs=1000.0, maprange="maximal", nv=16)
        assert np.array_equal(sf, sfc) and Tx[c].shape == Tc.shape
        scale = np.abs(Tc).max()
        assert np.abs(Tx[c] - Tc).max() <= 2e-5 * scale
        assert np.abs(Tx[c].sum(0) - Tc.sum(0)).max() <= 1e-5 * np.abs(Tc.sum(0)).max()
    Td, _ = _rs.ssq_cwt_batch(x.astype(np.float32), fs=1000.0, maprange="maximal", nv=16, device_out=True)
    assert isinstance(Td, torch.Tensor) and Td.is_cuda
    assert np.abs(Td.cpu().numpy() - Tx).max() <= 2e-5 * np.abs(Tx).max()
    T2, _ = _rs.ssq_cwt_batch(x, fs=1000.0, maprange="maximal", nv=16, out=Tx)
    assert T2 is Tx
