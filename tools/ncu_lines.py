"""Per-source-line instruction table of one kernel in an .ncu-rep (source page).
usage: python tools/ncu_lines.py report.ncu-rep <kernel-id-spec e.g. ::regex:fft128:2> elements_in_launch [top]"""
import collections
import csv
import subprocess
import sys

rep, kid, nel = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
# (source text with embedded quotes is not escaped by ncu: address the numeric columns from the right)
iI, iT, iS = (hdr.index(n) - len(hdr) for n in ("Instructions Executed", "Thread Instructions Executed", "# Samples"))
agg = collections.OrderedDict()
for r in rows:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    agg[int(r[0])] = (r[1], f(r[iI]), f(r[iT]), f(r[iS]))
tot_i = sum(v[1] for v in agg.values())
tot_s = sum(v[3] for v in agg.values())
print(f"total warp-instructions {tot_i:.0f} = {32 * tot_i / nel:.1f} per element (32 lanes); samples {tot_s:.0f}")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:5d} inst/el {32 * v[1] / nel:6.2f} ({100 * v[1] / tot_i:4.1f}%) samp {100 * v[3] / max(tot_s, 1):4.1f}% | {v[0].strip()[:110]}")
