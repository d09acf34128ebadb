"""Record the DRAM traffic of the benchmarked kernel for bench.py's `roofline.traffic`.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:ssq_stft512 --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 3 \
        --no-cpu-baseline --no-side-configs
    python tools/record_dram.py gpurun_out/dram.csv

writes profiles/bench_dram.json: bytes of the LAST captured launch (read + write), the commit and a fingerprint of
the kernel sources at capture time (bench.py prints both fingerprints so a stale record is visible)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
by = {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    by.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
full = [v for v in by.values() if "ssq_stft512" in v["name"]]
big = max(full, key=lambda v: v.get("dram__bytes_write.sum", 0))
rec = {"kernel": big["name"], "grid": big["grid"],
       "dram_bytes_read": big["dram__bytes_read.sum"], "dram_bytes_write": big["dram__bytes_write.sum"],
       "dram_bytes_per_launch": big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"],
       "gpu_time_ns_under_ncu": big.get("gpu__time_duration.sum"),
       "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
       "kernel_sources_sha": bench.kernel_sources_sha(), "source_csv": os.path.basename(sys.argv[1])}
json.dump(rec, open(os.path.join(ROOT, "profiles", "bench_dram.json"), "w"), indent=1)
print(rec)
