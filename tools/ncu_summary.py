"""Summarise an .ncu-rep: key raw metrics + per-source-line instruction / stall table.
usage: python tools/ncu_summary.py report.ncu-rep frames_in_launch [out_prefix]"""
import collections
import csv
import subprocess
import sys

rep, nfr = sys.argv[1], float(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
lines = ['metric,unit,value']
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        lines.append(f'"{h}","{u}","{v}"')
txt = "\n".join(lines)
print(txt)
if out:
    open(out + "_full.csv", "w").write(txt + "\n")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and r[0] != '' and r[2] == '-':
        agg[(cur, int(r[0]))] = (r[1][:80], int(r[7]), int(r[4]) if r[4].isdigit() else 0,
                                 r[hdr.index('L1 Wavefronts Shared')], r[hdr.index('L1 Wavefronts Shared Ideal')])
tots = max(1, sum(v[2] for v in agg.values()))
lines = ["# per source line: warp-instructions per frame, % of stall samples, shared wavefronts per frame (actual / ideal)"]
for k, v in sorted(agg.items()):
    if v[1] / nfr >= 2 or v[2] / tots > 0.005:
        lines.append(f"{k[0]:22s}{k[1]:4d} inst/frame {v[1]/nfr:7.1f} samp% {100*v[2]/tots:5.1f} wf/frame "
                     f"{float(v[3] or 0)/nfr:6.1f} / {float(v[4] or 0)/nfr:6.1f} | {v[0]}")
txt = "\n".join(lines)
print(txt)
if out:
    open(out + "_lines.txt", "w").write(txt + "\n")
