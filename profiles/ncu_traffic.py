#!/usr/bin/env python3
"""Per-kernel totals from an ncu --csv launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
Usage: python profiles/ncu_traffic.py launches.csv frames_per_step [out.json]
Prints one row per kernel name (launches, total ms, share of device time, DRAM GB per launch, achieved DRAM GB/s) and writes
the per-launch DRAM traffic of the dominant kernel (conv_i16_tc2_kernel<3, .>) for bench.py's roofline.traffic."""
import csv
import json
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launch = {}
for r in rows[1:]:
    d = launch.setdefault(r[iid], {"k": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("y2::<unnamed>::", "").replace("(int)", "")})
    d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
agg = {}
for d in launch.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'DRAM MB/launch':>15s} {'DRAM GB/s':>10s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {a[0]:8d} {a[1]:10.3f} {a[1] / tot:7.3f} {a[2] / a[0] / 1e6:15.2f} {a[2] / (a[1] * 1e-3) / 1e9 if a[1] else 0:10.1f}")
dom = [(k, a) for k, a in agg.items() if re.match(r"conv_i16_tc2_kernel<3, \d+, 0>", k)]      # 3x3, 128-channel items (not the HALF instantiation)
if dom and len(sys.argv) > 3:
    n = sum(a[0] for _, a in dom)
    out = {"kernel_prefix": "conv_i16_tc2_kernel<3", "frames_per_step": int(sys.argv[2]), "launches": n,
           "dram_bytes_per_launch": sum(a[2] for _, a in dom) / n, "ms_per_launch_under_ncu": sum(a[1] for _, a in dom) / n,
           "share_of_device_time_under_ncu": sum(a[1] for _, a in dom) / tot,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                     "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check` (profiles/r2_bench_launches_ncu.csv)"}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out))
