"""Host model of the shared-memory banks in the reassignment of the n_fft 512 kernel: the kernel's step schedule
(lane, step -> source bin) applied to the oracle's destination bins of the bench signal; wavefronts per 32-bit tag
access and per 64-bit accumulator access for a candidate index map.  Reproduces ncu's counts of the shipped layout
(3.14 / 4.75, profiles/r2/r2m_*) -- layouts can be searched here instead of on the GPU.  Test infrastructure: uses
oracle/."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import bench
from oracle import ssq_oracle as O
x = bench.make_neural_cpu(1, 120000, bench.FS, 0x5351+7)
w = np.hanning(512)
ref, _, ao = O.ssq_stft(x[0], w, n_fft=512, hop_len=32, fs=bench.FS, return_aux=True)
K = ao["k"]
src = np.zeros((8,32), int)
for l in range(32):
    rho = l & 7
    for r in range(8):
        if l == 0: src[r,l] = 64*r if r < 4 else 32 + 64*(r-4)
        else:
            ka = l + 64*((r+rho)&7); src[r,l] = min(ka, 512-ka)
def wf32(idx):  # 32-bit words
    tot=0
    banks={}
    for i in set(idx.tolist()):
        banks.setdefault(i%32,set()).add(i)
    return max(len(v) for v in banks.values())
def wf64(idx):
    t=0
    for h in (idx[:16], idx[16:]):
        banks={}
        for i in set(h.tolist()):
            banks.setdefault(i%16,set()).add(i)
        t+=max(len(v) for v in banks.values())
    return t
def evaluate(name, tagmap, accmap, srcs=src):
    a=b=n=0
    for f in range(0, K.shape[1], 11):
        for r in range(8):
            kb = K[srcs[r], f]
            a += wf32(tagmap(kb)); b += wf64(accmap(kb)); n+=1
    print(f"{name}: tag {a/n:.2f} wavefronts/access, acc {b/n:.2f}")
ident=lambda k:k
evaluate("current (tag k, acc k+(k>>6))", ident, lambda k:k+(k>>6))
evaluate("acc k", ident, ident)
