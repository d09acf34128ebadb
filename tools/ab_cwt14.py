"""A/B of the ssq_cwt plans on one channel of BASELINE configs[2] geometry (N = 2^20, nv 32): the last two radix-128
passes in one CTA (fft14_fused_kernel) against the three-pass plan (context option no_cwt_14): destination rows, Tx
and time."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import make_neural  # noqa: E402
from ssqueeze_rs_b200.batch import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
x = make_neural(torch, ch, n, 30000.0, torch.device("cuda", 0), 3)
res = {}
for opt in (1, 0):
    eng.ctx.set_option("no_cwt_14", opt)
    Tx, sf, aux = eng.ssq_cwt(x, "gmw", None, fs=30000.0, nv=32, maprange="maximal", return_aux=True)
    torch.cuda.synchronize()
    out = torch.empty_like(Tx)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.ssq_cwt(x, "gmw", None, fs=30000.0, nv=32, maprange="maximal", out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[opt] = (Tx, aux["kb"], min(ts))
    print(f"no_cwt_14={opt}: {min(ts):.2f} ms for {ch} channel(s) ({min(ts) / ch:.2f} ms per channel), kernel {eng.last_kernel_name()}")
T1, k1, _ = res[1]
T0, k0, _ = res[0]
diff = (k1 != k0)
print("rows:", k1.numel(), "differing destination rows:", int(diff.sum()), f"({float(diff.float().mean()):.2e})",
      "max |row difference|:", int((k1 - k0).abs().max()))
scale = float(T1.abs().max())
print("Tx max abs difference / max |Tx|:", float((T1 - T0).abs().max()) / scale,
      " column sums rel diff:", float((T1.sum(1) - T0.sum(1)).abs().max() / T1.sum(1).abs().max()))
