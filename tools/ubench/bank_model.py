import sys, numpy as np
sys.path.insert(0, '/root/repo')
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
