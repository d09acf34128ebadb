"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/ssqcuda.h declares; compute calls fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ssqcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssq_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    from ssqueeze_rs_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in ssqcuda.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_shapes(built_lib):
    from ssqueeze_rs_b200 import _rs
    assert _rs.hello_from_bin() == "Hello from ssqueeze!"  # lib.rs:16-19
    assert "ssqcuda" in _rs.version()
    a, b = ctypes.c_int64(), ctypes.c_int64()
    assert built_lib.ssq_stft_shape(1000, 256, 64, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (129, 16)  # tests/stft_test.py expected shape
    assert built_lib.ssq_stft_shape(1000, 256, 0, ctypes.byref(a), ctypes.byref(b)) != 0  # hop 0 panics upstream
    assert built_lib.ssq_cwt_shape(1000, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (2048, 524)
    assert built_lib.ssq_cwt_shape(1 << 20, ctypes.byref(a), ctypes.byref(b)) == 0
    assert a.value == 1 << 21
    ns = built_lib.ssq_cwt_default_scales(1 << 20, 32, 0, None)
    assert ns == 576  # BASELINE config 3


def test_default_scales_match_oracle(built_lib):
    from oracle import ssq_oracle as O
    for simd in (0, 1):
        ns = built_lib.ssq_cwt_default_scales(1000, 16, simd, None)
        out = np.empty(ns)
        built_lib.ssq_cwt_default_scales(1000, 16, simd, out.ctypes.data)
        assert np.allclose(out, O.generate_log_scales(1000, 16, simd=bool(simd)), rtol=1e-14)


def test_argument_validation_without_gpu(built_lib):
    from ssqueeze_rs_b200 import _rs
    x = np.zeros(100)
    with pytest.raises(TypeError):
        _rs.stft(x.astype(np.float32), 16, 4, np.ones(16), "reflect")  # PyReadonlyArray1<f64>
    with pytest.raises(TypeError):
        _rs.ssq_stft([0.0] * 10, np.ones(4))
    with pytest.raises(ValueError):
        _rs.ssq_stft(x, np.ones(32), n_fft=16)  # ssq_stft.rs:96-101, raised before any device work
    with pytest.raises(ValueError):
        _rs.cwt(x, t=np.array([0.0]))  # cwt.rs:68-70
    from ssqueeze_rs_b200 import PanicException
    with pytest.raises(PanicException):
        _rs.ssq_stft(x, np.ones(0), n_fft=0, win_len=0)  # n_fft - 1 underflows (stft_utils.rs:20); no buffer is sized from it


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device every compute entry point must raise, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ssqueeze_rs_b200 import _rs, SsqError
    with pytest.raises(SsqError):
        _rs.ssq_stft(np.zeros(64), np.ones(16), n_fft=16)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ssqueeze_rs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                s = open(os.path.join(dp, f)).read()
                assert "oracle" not in s.replace("no Python/NumPy fallback", ""), f"{f} mentions the oracle"


def test_host_only_entry_points_without_a_gpu():
    """Entry points that need no device: the admissibility integral (host math, double) against the oracle's
    quadrature, the shape queries and the default scales against the oracle."""
    import ctypes as C
    import numpy as np
    from oracle import ssq_oracle as O
    from ssqueeze_rs_b200 import _lib
    lib = _lib.load()
    for wid, name in ((0, "gmw"), (1, "morlet")):
        out = C.c_double(0.0)
        assert lib.ssq_cwt_admissibility(wid, C.addressof(out)) == 0
        assert np.isclose(out.value, O.adm_ssq(name), rtol=1e-9), (name, out.value)
    nf, nt = C.c_int64(), C.c_int64()
    assert lib.ssq_stft_shape(1000, 256, 64, C.byref(nf), C.byref(nt)) == 0
    assert (nf.value, nt.value) == (129, 16)  # tests/stft_test.py:137-151 of the reference
    for n, nv in ((1000, 32), (1 << 20, 32), (4097, 16)):
        ns = lib.ssq_cwt_default_scales(n, nv, 0, C.c_void_p(0))
        sc = np.empty(ns)
        lib.ssq_cwt_default_scales(n, nv, 0, C.c_void_p(sc.ctypes.data))
        so = O.generate_log_scales(n, nv)
        assert len(so) == ns and np.allclose(sc, so, rtol=1e-13)


def test_ssq_stft_batch_argument_validation_without_gpu(built_lib):
    """The batched drop-in rejects what the scalar call rejects, before anything touches a device."""
    from ssqueeze_rs_b200 import _rs
    w = np.hanning(64)
    with pytest.raises(TypeError):
        _rs.ssq_stft_batch(np.zeros(100), w)                      # 1-D: that is `ssq_stft`
    with pytest.raises(TypeError):
        _rs.ssq_stft_batch(np.zeros((2, 100), dtype=np.int16), w)
    with pytest.raises(ValueError):
        _rs.ssq_stft_batch(np.zeros((2, 100)), w, n_fft=32)       # window longer than n_fft (ssq_stft.rs:96-101)
    with pytest.raises(ValueError):
        _rs.ssq_stft_batch(np.zeros((2, 100)), w, n_fft=64, win_len=32)
    with pytest.raises(OverflowError):
        _rs.ssq_stft_batch(np.zeros((2, 100)), w, hop_len=-1)
    with pytest.raises(TypeError):
        _rs.ssq_cwt_batch(np.zeros(100))
    with pytest.raises(TypeError):
        _rs.stft_batch(np.zeros(100), 64, 16, w, "reflect")
    with pytest.raises(OverflowError):
        _rs.stft_batch(np.zeros((2, 100)), -64, 16, w, "reflect")


def test_parquet_recording_streams_the_table(tmp_path):
    """The parquet reader of the streaming path (host only): forward slices of any size equal the table the reference's
    script would have loaded whole (tests/stft_test.py:374-377), float and int16 columns, a column subset."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from ssqueeze_rs_b200.batch import ParquetRecording
    rng = np.random.default_rng(0)
    data = rng.standard_normal((10_001, 5))
    pq.write_table(pa.table({f"ch{j}": data[:, j] for j in range(5)}), tmp_path / "f.parquet", row_group_size=3000)
    rec = ParquetRecording(tmp_path / "f.parquet", batch_rows=1024)
    assert rec.shape == (10_001, 5) and rec.dtype == np.float32 and rec.ndim == 2
    pos, out = 0, []
    for n in (1, 999, 4096, 17, 10_001):
        blk = rec[pos:pos + n]
        assert blk.flags["C_CONTIGUOUS"] and blk.dtype == np.float32
        out.append(blk)
        pos = min(pos + n, 10_001)
    assert np.array_equal(np.concatenate(out), data.astype(np.float32))
    with pytest.raises(ValueError):
        rec[0:10]
    ints = rng.integers(-3000, 3000, size=(5000, 3), dtype=np.int16)
    pq.write_table(pa.table({f"c{j}": ints[:, j] for j in range(3)}), tmp_path / "i.parquet")
    rec = ParquetRecording(tmp_path / "i.parquet", columns=["c2", "c0"])
    assert rec.dtype == np.int16 and rec.shape == (5000, 2)
    assert np.array_equal(rec[0:5000], ints[:, [2, 0]])
    with pytest.raises(ValueError):
        ParquetRecording(tmp_path / "i.parquet", columns=["nope"])
