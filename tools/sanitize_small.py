"""Small invocations of every shared-memory kernel for compute-sanitizer (racecheck / memcheck)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ssqueeze_rs_b200.batch import Engine, SsqStftStream  # noqa: E402

eng = Engine(0)
g = torch.Generator(device="cuda")
g.manual_seed(0)
x = torch.randn((3, 6000), generator=g, device="cuda") * 10
win = np.hanning(512)
Tx = eng.ssq_stft(x, win, 512, 32, 30000.0)             # h32r
Sx = eng.stft(x, win, 512, 32)                           # h32r<stft>
xr = eng.istft(Sx, win, 512, 32, N=6000)                 # istft512_tile + finalize
T2 = eng.ssq_stft(x, win, 512, 17, 30000.0)              # tile kernel, other hop
T3 = eng.ssq_stft(x[:, :3000], np.hanning(256), 256, 64, 30000.0)  # generic
tone = torch.sin(torch.arange(6000, device="cuda") * 0.3).repeat(2, 1).contiguous()
T4 = eng.ssq_stft(tone, win, 512, 32, 30000.0)           # collision paths
W = eng.ssq_cwt(x[:, :5000], "gmw", None, fs=1000.0, nv=4, maprange="maximal")  # pad 2^13: fft128 + generic passes
st = SsqStftStream(eng, 3, 6000, 2500, win, 512, 32, 30000.0)
rec = x.t().contiguous()
parts = [st.push(rec[i:i + 2500].contiguous()) for i in range(0, 6000, 2500)]
torch.cuda.synchronize()
assert torch.equal(torch.cat(parts, 2), Tx)
print("ok", float(Tx.abs().sum()), float((xr - x).abs().max()))
