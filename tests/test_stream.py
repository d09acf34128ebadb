"""GPU: streaming ssq_stft over interleaved chunks equals the whole-signal
transform bit for bit, for ragged chunk sizes, int16 and float32 input, both kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(n_total, channels, chunks, n_fft, hop, dtype, padtype="reflect", squeezing="sum"):
    import torch
    from ssqueeze_rs_b200.batch import Engine, SsqStftStream
    eng = Engine(0)
    rng = np.random.default_rng(n_total + channels + hop)
    if dtype == "i16":
        rec = rng.integers(-3000, 3000, size=(n_total, channels), dtype=np.int16)
        scale = 0.195  # uV per count, as in extracellular recordings
        x = (rec.astype(np.float32) * np.float32(scale)).T.copy()
    else:
        rec = (rng.standard_normal((n_total, channels)) * 20).astype(np.float32)
        scale = 1.0
        x = rec.T.copy()
    win = np.hanning(n_fft)
    whole = eng.ssq_stft(torch.from_numpy(x).cuda(), win, n_fft, hop, 30000.0, padtype=padtype, squeezing=squeezing)
    st = SsqStftStream(eng, channels, n_total, max(chunks), win, n_fft, hop, 30000.0, padtype=padtype,
                       squeezing=squeezing)
    assert st.total_frames == whole.shape[2]
    parts, pos = [], 0
    ci = 0
    while pos < n_total:
        c = min(chunks[ci % len(chunks)], n_total - pos)
        parts.append(st.push(torch.from_numpy(rec[pos:pos + c].copy()).cuda(), scale))
        pos += c
        ci += 1
    torch.cuda.synchronize()
    got = torch.cat(parts, dim=2)
    assert got.shape == whole.shape
    assert torch.equal(got, whole), float((got - whole).abs().max())
    st.close()
    return eng.last_kernel_name()


@pytest.mark.parametrize("dtype", ["i16", "f32"])
def test_stream_fast_path_ragged_chunks(dtype):
    name = _run(50000, 5, [7001, 333, 12288, 1, 20000], 512, 32, dtype)
    assert "h32r" in name


def test_stream_short_recordings_and_options():
    _run(300, 3, [100], 512, 32, "f32")                 # shorter than n_fft: both paddings in one frame
    _run(2000, 2, [999, 1], 512, 32, "i16", padtype="zero", squeezing="lebesgue")
    _run(33, 1, [5], 512, 32, "f32")


def test_stream_other_kernels_and_hops():
    assert "256" in _run(9000, 4, [2500, 777], 256, 64, "f32")        # four frames per warp
    assert "1024" in _run(20000, 3, [4100, 999], 1024, 256, "i16")
    _run(6000, 2, [1700], 512, 300, "i16")              # hop > n_fft/2: end-reflection reaches before the frame start
    assert "generic" in _run(4000, 2, [1111], 128, 1, "f32")


def test_stream_hop_larger_than_n_fft_small_chunks():
    """hop > n_fft with chunks smaller than the hop: the start of the next frame lies beyond the samples received so
    far, the samples in between belong to no frame (they must be dropped, not written before the carry-over buffer)."""
    _run(30000, 3, [300], 512, 1000, "f32")
    _run(30000, 2, [300, 77, 1500], 512, 512 + 17, "i16")
    assert "256" in _run(9000, 4, [100], 256, 256 + 17, "f32")
    _run(5000, 2, [64], 128, 700, "f32")                # generic kernel


@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_recording_feeder_from_a_memory_mapped_file(tmp_path, dtype):
    """Host half of the streaming path: a (samples, channels) file on disk, memory-mapped as the reference's scripts do
    (tests/stft_ssq_test.py:218-283), fed in ragged chunks through the pinned ring; equals the whole-signal transform
    bit for bit, with more pushes than ring slots (slot reuse) and a last chunk shorter than the others."""
    import torch
    from ssqueeze_rs_b200.batch import Engine, RecordingFeeder
    eng = Engine(0)
    n_total, channels = 70001, 6
    rng = np.random.default_rng(11)
    if dtype == np.int16:
        rec = rng.integers(-3000, 3000, size=(n_total, channels), dtype=np.int16)
        scale = 0.195
    else:
        rec = (rng.standard_normal((n_total, channels)) * 20).astype(np.float32)
        scale = 1.0
    path = tmp_path / "probe.dat"
    rec.tofile(path)
    mm = np.memmap(path, dtype=dtype, mode="r").reshape(-1, channels)
    win = np.hanning(512)
    x = (rec.astype(np.float32) * np.float32(scale)).T.copy()
    whole = eng.ssq_stft(torch.from_numpy(x).cuda(), win, 512, 32, 30000.0)
    with RecordingFeeder(eng, mm, win, 512, 32, 30000.0, chunk=9000, scale=scale, depth=2) as feed:
        assert feed.total_frames == whole.shape[2]
        parts = [t.clone() for t in feed]
    torch.cuda.synchronize()
    assert len(parts) == 8
    got = torch.cat(parts, dim=2)
    assert got.shape == whole.shape
    assert torch.equal(got, whole), float((got - whole).abs().max())


def test_recording_feeder_argument_errors():
    from ssqueeze_rs_b200.batch import Engine, RecordingFeeder
    eng = Engine(0)
    with pytest.raises(ValueError):
        RecordingFeeder(eng, np.zeros(100, np.int16), np.hanning(512))
    with pytest.raises(ValueError):
        RecordingFeeder(eng, np.zeros((100, 2), np.float64), np.hanning(512))


def test_recording_feeder_edge_cases():
    """Fewer pushes than ring slots, a single-sample last chunk, hop > n_fft (samples between frames are dropped), the
    n_fft 1024 kernel behind the same feeder, and a strided (non-contiguous) view of the recording."""
    import torch
    from ssqueeze_rs_b200.batch import Engine, RecordingFeeder
    eng = Engine(0)
    rng = np.random.default_rng(3)
    for n_total, channels, chunk, n_fft, hop, depth in ((5000, 3, 5000, 512, 32, 4), (4097, 2, 4096, 512, 32, 2),
                                                       (9000, 2, 700, 512, 529, 2), (20000, 3, 3000, 1024, 256, 3)):
        rec = (rng.standard_normal((n_total, channels)) * 20).astype(np.float32)
        win = np.hanning(n_fft)
        whole = eng.ssq_stft(torch.from_numpy(rec.T.copy()).cuda(), win, n_fft, hop, 30000.0)
        with RecordingFeeder(eng, rec, win, n_fft, hop, 30000.0, chunk=chunk, depth=depth) as feed:
            got = torch.cat([t.clone() for t in feed], dim=2)
        torch.cuda.synchronize()
        assert torch.equal(got, whole), (n_total, channels, chunk, n_fft, hop)
    wide = (rng.standard_normal((6000, 8)) * 20).astype(np.float32)
    view = wide[:, ::2]  # every other channel: rows are not contiguous
    whole = eng.ssq_stft(torch.from_numpy(np.ascontiguousarray(view.T)).cuda(), np.hanning(512), 512, 32, 30000.0)
    with RecordingFeeder(eng, view, np.hanning(512), 512, 32, 30000.0, chunk=2500) as feed:
        got = torch.cat([t.clone() for t in feed], dim=2)
    assert torch.equal(got, whole)


def test_recording_feeder_from_parquet(tmp_path):
    """The reference's on-disk format end to end: parquet columns -> ParquetRecording -> pinned ring -> ssq_stft
    (n_fft 1024, hop 256 as in tests/stft_test.py:386) equals the transform of the table loaded whole."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    import torch
    from ssqueeze_rs_b200.batch import Engine, ParquetRecording, RecordingFeeder
    eng = Engine(0)
    rng = np.random.default_rng(4)
    data = (rng.standard_normal((50_000, 4)) * 20).astype(np.float32)
    pq.write_table(pa.table({f"ch{j}": data[:, j] for j in range(4)}), tmp_path / "rec.parquet", row_group_size=8192)
    win = np.hanning(1024)
    whole = eng.ssq_stft(torch.from_numpy(data.T.copy()).cuda(), win, 1024, 256, 24414.0625)
    rec = ParquetRecording(tmp_path / "rec.parquet", batch_rows=5000)
    with RecordingFeeder(eng, rec, win, 1024, 256, 24414.0625, chunk=12_000) as feed:
        got = torch.cat([t.clone() for t in feed], dim=2)
    torch.cuda.synchronize()
    assert torch.equal(got, whole)


@pytest.mark.parametrize("n_fft,hop", [(512, 32), (1024, 256), (256, 64), (128, 5)])
def test_stream_stft_mode_and_modulated_ssq(n_fft, hop):
    """The stft caller of the reference (tests/stft_test.py:215-271) as a stream -- frames of Sx equal to the
    whole-signal stft bit for bit on every kernel family (hop 32 packs two frames per FFT) -- and the modulated ssq_stft."""
    import torch
    from ssqueeze_rs_b200.batch import Engine, RecordingFeeder
    eng = Engine(0)
    rng = np.random.default_rng(n_fft + hop)
    rec = (rng.standard_normal((30_001, 3)) * 20).astype(np.float32)
    win = np.hanning(n_fft)
    xd = torch.from_numpy(rec.T.copy()).cuda()
    whole = eng.stft(xd, win, n_fft, hop)
    with RecordingFeeder(eng, rec, win, n_fft, hop, chunk=7777, transform="stft") as feed:
        got = torch.cat([t.clone() for t in feed], dim=2)
    if n_fft in (256, 512, 1024):
        # the register kernels put two frames through one FFT in stft mode (z = x_A w + i x_B w): which of a pair a
        # frame is depends on where the push starts, so the stream equals the whole-signal transform to fp32 rounding
        # instead of bit for bit (the generic kernel, last case, does not pair: bit-equal)
        err = float((got - whole).abs().max() / whole.abs().max())
        assert err < 2e-6, err
    else:
        assert torch.equal(got, whole), float((got - whole).abs().max())
    whole_m = eng.ssq_stft(xd, win, n_fft, hop, 30000.0, modulated=True)
    with RecordingFeeder(eng, rec, win, n_fft, hop, 30000.0, chunk=7777, modulated=True) as feed:
        got = torch.cat([t.clone() for t in feed], dim=2)
    assert torch.equal(got, whole_m)
