"""Wall time of the drop-in scalar calls (`_rs.ssq_stft`, `_rs.stft`: float64 in, complex128 out), one channel of
BASELINE configs[1] per call -- the loop of tests/stft_ssq_test.py:230-251 in the reference."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402

from ssqueeze_rs_b200 import _rs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_800_000
x = np.random.default_rng(0).standard_normal(n)
win = np.hanning(512)
for name, fn in (("ssq_stft", lambda: _rs.ssq_stft(x, win, n_fft=512, hop_len=32, fs=30000.0)),
                 ("stft", lambda: _rs.stft(x, 512, 32, win, "reflect"))):
    fn()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    print(f"{name}: {min(ts) * 1e3:.1f} ms per call (best of 4), {n / min(ts) / 1e6:.1f} Msamples/s, "
          f"output {out[0].nbytes / 1e6:.0f} MB {out[0].dtype}")

# the batched drop-in: every channel of a recording in one call, complex64 out (pinned host memory / on the device)
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
xb = np.random.default_rng(1).standard_normal((ch, n)).astype(np.float32)
out, _ = _rs.ssq_stft_batch(xb, win, n_fft=512, hop_len=32, fs=30000.0)  # allocates the pinned result once
for name, fn in (("ssq_stft_batch -> pinned host", lambda: _rs.ssq_stft_batch(xb, win, n_fft=512, hop_len=32, fs=30000.0, out=out)),
                 ("ssq_stft_batch -> device", lambda: _rs.ssq_stft_batch(xb, win, n_fft=512, hop_len=32, fs=30000.0, device_out=True))):
    fn()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = fn()
        if not isinstance(r[0], np.ndarray):
            import torch
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del r
    print(f"{name}: {ch} channels, {min(ts) * 1e3:.1f} ms per call (best of 3), {ch * n / min(ts) / 1e6:.1f} Msamples/s")
