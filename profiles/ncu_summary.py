#!/usr/bin/env python3
"""One row per profiled launch from an ncu report: python profiles/ncu_summary.py report.ncu-rep [out.csv]
(runs `ncu -i ... --page raw --csv` and keeps the columns the roofline needs; achieved GB/s = DRAM bytes / duration)."""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
out = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
out.writerow([hdr[i] for i in idx] + ["dram GB/s (read+write)/duration"])
out.writerow([units[i] for i in idx] + ["GB/s"])
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
to_s = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        t = float(r[hdr.index("gpu__time_duration.sum")]) * to_s.get(units[hdr.index("gpu__time_duration.sum")], 1e-6)
        b = sum(float(r[hdr.index(k)]) * to_bytes.get(units[hdr.index(k)], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        gbs = f"{b / t / 1e9:.1f}"
    except Exception:
        gbs = ""
    out.writerow([r[i] for i in idx] + [gbs])
