"""ctypes loader of oracle/libssq_ref.so (the C restatement; TEST INFRASTRUCTURE:
parity checks and bench.py's CPU baseline only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libssq_ref.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(so)
        _lib.ssq_stft_ref.restype = C.c_int
        _lib.ssq_stft_ref.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ssq_ref_num_threads.restype = C.c_int
        _lib.ssq_ref_set_threads.argtypes = [C.c_int]
    return _lib


def num_threads():
    return load().ssq_ref_num_threads()


def use_all_cores():
    """The Rayon global pool uses every host core; do the same (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().ssq_ref_set_threads(n)
    return num_threads()


def ssq_stft(x, window_fit, n_fft, hop, fs, padtype="reflect", squeezing="sum", gamma=None, mode=0,
             want_Sx=False):
    """mode 0: as-written (ssq_stft.rs:122-303), 1: as-intended."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(window_fit, dtype=np.float64)
    assert len(w) == n_fft
    n_freqs, n_frames = n_fft // 2 + 1, (len(x) - 1) // hop + 1
    Tx = np.empty((n_freqs, n_frames), dtype=np.complex128)
    sf = np.empty(n_freqs, dtype=np.float64)
    Sx = np.empty((n_freqs, n_frames), dtype=np.complex128) if want_Sx else None
    rc = load().ssq_stft_ref(x.ctypes.data, len(x), w.ctypes.data, n_fft, hop, float(fs),
                             1 if padtype == "zero" else 0, 1 if squeezing == "lebesgue" else 0,
                             -1.0 if gamma is None else float(gamma), mode, Tx.ctypes.data, sf.ctypes.data,
                             Sx.ctypes.data if want_Sx else None)
    if rc != 0:
        raise RuntimeError("ssq_stft_ref failed")
    return (Tx, sf, Sx) if want_Sx else (Tx, sf)
