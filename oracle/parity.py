"""Parity gates of SURVEY.md 8(d) -- TEST INFRASTRUCTURE, NOT PRODUCT (tests/, bench.py's parity leg).

north_star: "reassignment bin indices identical except where w falls within fp32 rounding of a bin edge,
with such cases counted and reported".  This module turns that sentence into a decision per (source bin,
frame): given the destination bins the CUDA kernel used (`kb`, written by the kernel itself) and the float64
oracle's Sx, dSx, w and arg-min bins, every mismatch is classified as

  within_edge        the device bin is what the reference's rule (ssq_stft.rs:276-301: nearest grid point,
                     ties to the lower index, clamp) gives for some w' with |w' - w_f64| <= tol, tol < dw/2;
  ill_conditioned    the same with tol >= dw/2: |Sx| is so far below the frame's spectral level that an fp32
                     transform cannot resolve Im(dSx/Sx) to half a bin (includes SURVEY's energy gate
                     |Sx| < n_fft eps32 max_k|Sx|, counted separately as below_energy_gate);
  gate_edge          one side gated the bin (|Sx| < gamma), the other did not, and |Sx| is within the fp32
                     error of gamma;
  unexplained        everything else.  The tests and bench.py assert unexplained == 0.

tol is DERIVED from the fp32 error model of the kernel, not fitted to the mismatches:
  * one packed complex FFT per frame, Z = FFT(x w + i x dw s): each output carries an absolute error of
    about eps32 * sqrt(log2 n_fft) * R, R = rms_k |Z[k]| over the frame (random-walk growth of a
    radix-2^k FFT; Schatzman 1996) -- C_FFT below is the allowance in units of eps32 * R;
  * 2 Sx = Z[k] + conj Z[N-k], 2 V = (Z[k] - conj Z[N-k]) / i inherit it; q = Im(V / Sx) then moves by
    (e / |Sx|) (1 + |V / Sx|), e = C_FFT eps32 R; in grid units that is times cphase = (n_freqs-1)/(pi s);
  * the closed form |k - q cphase| adds a few roundings of its own magnitude (rcp.approx is 1 ulp):
    C_ARITH eps32 (k + |q| cphase).
`emulate_fp32_kernel` restates the kernel's arithmetic in NumPy float32 so the constants can be checked on a
CPU (tests/test_parity_gate_cpu.py): over the input families of the GPU tests (noise, the neural recipe,
tones, tiny amplitudes, random windows) the worst observed |binf32 - binf64| / tol is 0.35 with C_FFT = 3,
C_ARITH = 8, i.e. the bound holds with a factor ~3 to spare and is not loose by more than that; the GPU tests
report the same ratio for the device's own w (`max_err_over_tol`).
"""
from __future__ import annotations

import math

import numpy as np

from . import ssq_oracle as O

EPS32 = float(np.finfo(np.float32).eps)
C_FFT = 3.0
C_ARITH = 8.0


def s_scale_of(window_fit: np.ndarray) -> float:
    """The kernel packs z = x w + i x dw s with s = sqrt(sum w^2 / sum dw^2) (ssqcuda.cu stft_tables)."""
    dw = O.diff_window(np.asarray(window_fit, dtype=np.float64))
    sd = float((dw * dw).sum())
    sw = float((window_fit * window_fit).sum())
    return math.sqrt(sw / sd) if sd > 0.0 and sw > 0.0 else 1.0


def stft_bin_tolerance(Sx, dSx, n_fft, fs, window_fit):
    """tol [n_freqs, n_frames] in units of the ssq grid step, and the frame levels R [n_frames]."""
    n_freqs = Sx.shape[0]
    s = s_scale_of(window_fit)
    V = dSx * (s / fs)
    p = np.abs(Sx) ** 2 + np.abs(V) ** 2
    full = 2.0 * p.sum(axis=0) - p[0]
    if n_fft % 2 == 0:
        full = full - p[-1]
    R = np.sqrt(np.maximum(full, 0.0) / n_fft)  # rms over the n_fft bins of Z
    cphase = (n_freqs - 1.0) / (math.pi * s)
    aS = np.maximum(np.abs(Sx), 1e-300)
    ratio = np.abs(V) / aS
    with np.errstate(over="ignore", invalid="ignore"):
        tol = C_FFT * EPS32 * cphase * (R[None, :] / aS) * (1.0 + ratio)
        tol = tol + C_ARITH * EPS32 * (np.arange(n_freqs)[:, None] + ratio * cphase)
    return np.nan_to_num(tol, nan=np.inf, posinf=np.inf), R


def classify_stft_bins(kb_dev, aux_o, n_fft, fs, gamma=None, w_dev=None):
    """kb_dev: int [n_freqs, n_frames] from the kernel (-1 gated).  aux_o: the oracle's aux dict
    (Sx, dSx, w, k, window).  w_dev: the kernel's own w (Hz), optional -> max_err_over_tol.
    Returns the report dict."""
    Sx, dSx, w_o, k_o = aux_o["Sx"], aux_o["dSx"], aux_o["w"], aux_o["k"]
    n_freqs, n_frames = Sx.shape
    kb_dev = np.asarray(kb_dev).astype(np.int64)
    assert kb_dev.shape == k_o.shape
    g = 10.0 * O.EPS64 if gamma is None else float(gamma)
    dw = 0.5 * fs / (n_freqs - 1.0)
    tol, R = stft_bin_tolerance(Sx, dSx, n_fft, fs, aux_o["window"])
    mism = kb_dev != k_o
    gated_o, gated_d = k_o < 0, kb_dev < 0
    gate_diff = gated_o ^ gated_d
    # |Sx| against gamma with the fp32 error of Sx itself
    gate_edge = gate_diff & (np.abs(np.abs(Sx) - g) <= C_FFT * EPS32 * R[None, :] + 4 * EPS32 * np.abs(Sx))
    with np.errstate(invalid="ignore", over="ignore"):
        binf = np.where(np.isfinite(w_o), w_o / dw, np.nan)
    # the reference's rule applied to w' in [binf - tol, binf + tol] can give any bin in [lo, hi]:
    # nearest grid point with ties to the lower index = ceil(b - 0.5); clamp to the grid
    with np.errstate(invalid="ignore", over="ignore"):
        lo = np.ceil(np.nan_to_num(binf - tol, nan=0.0, neginf=-1e18) - 0.5)
        hi = np.ceil(np.nan_to_num(binf + tol, nan=0.0, posinf=1e18) - 0.5)
    lo = np.clip(lo, 0, n_freqs - 1)
    hi = np.clip(hi, 0, n_freqs - 1)
    consistent = (kb_dev >= lo) & (kb_dev <= hi) & ~gated_d & ~gated_o
    colmax = np.abs(Sx).max(axis=0, keepdims=True)
    below_gate = np.abs(Sx) < n_fft * EPS32 * np.maximum(colmax, 1e-300)
    well = tol < 0.5
    within_edge = mism & consistent & well
    ill = mism & consistent & ~well
    unexplained = mism & ~(consistent | gate_edge)
    rep = dict(
        bins_total=int(mism.size), mismatch_total=int(mism.sum()), within_edge=int(within_edge.sum()),
        ill_conditioned=int(ill.sum()), below_energy_gate=int((mism & below_gate).sum()),
        gate_edge=int((mism & gate_edge).sum()), unexplained=int(unexplained.sum()),
        mismatch_above_energy_gate=int((mism & ~below_gate).sum()),
        max_tol_at_edge_flips=float(tol[within_edge].max()) if within_edge.any() else 0.0)
    if w_dev is not None:
        both = np.isfinite(w_o) & np.isfinite(w_dev) & np.isfinite(tol) & (tol > 0)
        with np.errstate(invalid="ignore", over="ignore"):
            ratio = np.abs(np.asarray(w_dev, dtype=np.float64) - w_o) / dw / tol
        rep["max_err_over_tol"] = float(ratio[both].max()) if both.any() else 0.0
    rep["_unexplained_mask"] = unexplained
    return rep


def public(rep):
    """The report without the mask (JSON-able)."""
    return {k: v for k, v in rep.items() if not k.startswith("_")}


def reaccumulate(Sx_o, kb_dev, fs, squeezing="sum"):
    """Tx the reference's accumulation (ssq_stft.rs:290-299) gives for the ORACLE's Sx with the DEVICE's bins:
    equal to the device's Tx within rtol iff the kernel put every value where it says it did."""
    n_freqs, n_frames = Sx_o.shape
    dw = 0.5 * fs / (n_freqs - 1.0)
    T = np.zeros_like(Sx_o)
    cols = np.arange(n_frames)
    for i in range(n_freqs):
        ok = kb_dev[i] >= 0
        wgt = Sx_o[i] if squeezing != "lebesgue" else np.full(n_frames, 1.0 / n_freqs + 0j)
        np.add.at(T, (kb_dev[i][ok], cols[ok]), wgt[ok] * dw)
    return T


# ---------------------------------------------------------------------------------------------
# NumPy float32 restatement of the kernel's arithmetic (the same formulas, a different FFT
# factorisation): used on the CPU to check the tolerance model and that the gate has teeth.
# ---------------------------------------------------------------------------------------------
def emulate_fp32_kernel(x, window, n_fft, hop, fs, padtype="reflect", gamma=None, rule="reference"):
    """Returns kb [n_freqs, n_frames] (int, -1 gated) and binf (float32) the way the CUDA kernels compute
    them (stft_h32r.cuh h32r_item / stft_r1024.cuh r1k_item), in float32 throughout.
    rule: 'reference' | 'ties_up' | 'floor' | 'drop_out_of_range' -- the wrong rules exist to show that the
    classification rejects them."""
    x = np.asarray(x, dtype=np.float64)
    wfit = O.fit_window(np.asarray(window, dtype=np.float64), n_fft)
    dwin = O.diff_window(wfit)
    s = s_scale_of(wfit)
    padded = O._pad_stft(x, n_fft, padtype).astype(np.float32)
    fr = O._frames(padded, n_fft, hop)
    n_freqs = n_fft // 2 + 1
    z = (fr * wfit.astype(np.float32)[None, :]).astype(np.float32) + 1j * (
        fr * (dwin * s).astype(np.float32)[None, :]).astype(np.float32)
    Z = np.fft.fft(z.astype(np.complex64), axis=1)
    assert Z.dtype == np.complex64
    k = np.arange(n_freqs)
    Zk = Z[:, k]
    Zn = Z[:, (n_fft - k) % n_fft]
    f32 = np.float32
    c = (Zk.real + Zn.real).astype(f32)
    d = (Zk.imag - Zn.imag).astype(f32)
    a = (Zk.imag + Zn.imag).astype(f32)
    b = (Zn.real - Zk.real).astype(f32)
    den = (c * c + d * d).astype(f32)
    num = (b * c - a * d).astype(f32)
    cphase = f32((n_freqs - 1.0) / (math.pi * s))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        q = (num * (f32(1.0) / den)).astype(f32)
        binf = np.abs(k.astype(f32)[None, :] - q * cphase).astype(f32)
        if rule == "ties_up":
            r = np.floor(binf + f32(0.5))
        elif rule == "floor":
            r = np.floor(binf)
        else:
            r = np.ceil(binf - f32(0.5))
    g = 10.0 * O.EPS64 if gamma is None else float(gamma)
    gate2 = f32(-1.0) if g < 0 else f32(min(4.0 * g * g, 3.0e38))
    kb = np.clip(np.nan_to_num(r, nan=0.0, posinf=1e9, neginf=-1e9), 0, n_freqs - 1).astype(np.int64)
    if rule == "drop_out_of_range":
        kb = np.where(np.nan_to_num(r, nan=0.0) > n_freqs - 1, -1, kb)
    kb = np.where(den < gate2, -1, kb)
    return kb.T.copy(), binf.T.copy()
