"""Small driver for ncu / quick timing: runs the batched ssq_stft (or another
op) a few times on synthetic noise.  Not part of the product or the tests."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ssqueeze_rs_b200.batch import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=32)
ap.add_argument("--samples", type=int, default=1_800_000)
ap.add_argument("--n-fft", type=int, default=512)
ap.add_argument("--hop", type=int, default=32)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--op", default="ssq_stft")
ap.add_argument("--signal", default="noise", choices=["noise", "tone", "neural"])
ap.add_argument("--nv", type=int, default=32)
a = ap.parse_args()

eng = Engine(0)
g = torch.Generator(device="cuda")
g.manual_seed(0)
if a.signal == "noise":
    x = torch.randn((a.channels, a.samples), generator=g, device="cuda") * 10
elif a.signal == "tone":
    t = torch.arange(a.samples, device="cuda", dtype=torch.float64) / 30000.0
    x = torch.sin(2 * np.pi * 1000.0 * t).to(torch.float32).repeat(a.channels, 1).contiguous()
else:
    from bench import make_neural
    x = make_neural(torch, a.channels, a.samples, 30000.0, torch.device("cuda", 0), 1)
win = np.hanning(a.n_fft)
nfq, nfr = a.n_fft // 2 + 1, (a.samples - 1) // a.hop + 1
if a.op in ("ssq_cwt", "cwt"):
    import ctypes as C
    from ssqueeze_rs_b200._lib import load
    ns = load().ssq_cwt_default_scales(a.samples, a.nv, 0, C.c_void_p(0))
    out = torch.empty((a.channels, ns, a.samples), dtype=torch.complex64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(a.iters):
        ev0.record()
        if a.op == "ssq_cwt":
            eng.ssq_cwt(x, "gmw", None, fs=30000.0, nv=a.nv, maprange="maximal", out=out)
        else:
            out = eng.cwt(x, "gmw", None, fs=30000.0, nv=a.nv)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)  # Engine binds the legacy default stream: the events bracket the kernels
        ab = a.channels * (4 * a.samples + 8 * ns * a.samples)
        print(f"{a.op} ns={ns} iter {it}: {ms:.3f} ms  {a.channels * a.samples / ms / 1e3:.2f} Msamples/s  "
              f"{ab / ms / 1e6:.1f} GB/s algorithmic")
    sys.exit(0)
if a.op in ("issq_stft", "icwt", "issq_cwt"):
    # column-sum inverses on random coefficients (timing only): rows x samples complex64 per channel
    import ctypes as C
    from ssqueeze_rs_b200._lib import load, raise_status
    rows = nfq if a.op == "issq_stft" else load().ssq_cwt_default_scales(a.samples, a.nv, 0, C.c_void_p(0))
    M = torch.randn((a.channels, rows, a.samples, 2), generator=g, device="cuda")
    y = torch.empty((a.channels, a.samples), dtype=torch.float32, device="cuda")
    sc = np.ascontiguousarray(2.0 ** (1.0 + np.arange(rows) / a.nv))
    eng._bind_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(a.iters):
        ev0.record()
        if a.op == "issq_stft":
            eng.issq_stft_ptr(M.data_ptr(), a.channels, rows, a.samples, win, a.n_fft, 1.0, y.data_ptr())
        elif a.op == "icwt":
            st = load().ssq_icwt_batch_f32(eng.ctx.handle, M.data_ptr(), a.channels, rows, a.samples, 0,
                                           sc.ctypes.data, 1, a.samples, 0.0, 0, y.data_ptr())
            raise_status(st, eng.ctx.handle)
        else:
            st = load().ssq_issq_cwt_batch_f32(eng.ctx.handle, M.data_ptr(), a.channels, rows, a.samples, 0,
                                               sc.ctypes.data, y.data_ptr())
            raise_status(st, eng.ctx.handle)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        ab = a.channels * (4 * a.samples + 8 * rows * a.samples)
        print(f"{a.op} rows={rows} iter {it}: {ms:.3f} ms  {a.channels * a.samples / ms / 1e3:.1f} Msamples/s  "
              f"{ab / ms / 1e6:.1f} GB/s algorithmic")
    ref = M[0, :, :4096, 0].double().sum(0)
    print("check", float((y[0, :4096].double() / ref).std()))
    sys.exit(0)
if a.op == "istft":
    Sx = torch.empty((a.channels, nfq, nfr), dtype=torch.complex64, device="cuda")
    eng.stft(x, win, a.n_fft, a.hop, out=Sx)
    xr = torch.empty((a.channels, a.samples), dtype=torch.float32, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(a.iters):
        ev0.record()
        xr = eng.istft(Sx, win, a.n_fft, a.hop, N=a.samples)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        ab = a.channels * (4 * a.samples + 8 * nfq * nfr)
        err = float((xr - x).abs().max() / x.abs().max())
        print(f"istft iter {it}: {ms:.3f} ms (kernel {eng.ctx.last_kernel_ms():.3f})  {a.channels * a.samples / ms / 1e3:.1f} Msamples/s  "
              f"{ab / ms / 1e6:.1f} GB/s algorithmic  roundtrip max err {err:.2e}")
    sys.exit(0)
out = torch.empty((a.channels, nfq, nfr), dtype=torch.complex64, device="cuda")
torch.cuda.synchronize()
for it in range(a.iters):
    if a.op == "ssq_stft":
        eng.ssq_stft(x, win, a.n_fft, a.hop, 30000.0, out=out)
    elif a.op == "stft":
        eng.stft(x, win, a.n_fft, a.hop, out=out)
    ms = eng.ctx.last_kernel_ms()
    ab = a.channels * (4 * a.samples + 8 * nfq * nfr)
    print(f"{eng.last_kernel_name()} iter {it}: {ms:.3f} ms  {a.channels * a.samples / ms / 1e3:.1f} Msamples/s  "
          f"{ab / ms / 1e6:.1f} GB/s algorithmic")
torch.cuda.synchronize()
