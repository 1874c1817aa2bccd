#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a markdown table (time, DRAM bytes, DRAM %, tensor-pipe %, occupancy, ...).
    python tools/summarize_ncu.py gpurun_out/r01_kernels_final.ncu-rep profiles/r01_ncu_kernels.md
"""
import csv
import io
import subprocess
import sys

rep, dst = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}


def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")


def mb(r, k):
    u, v = units[col[k]], val(r, k)
    return v * (1000 if u.startswith("G") else (1 if u.startswith("M") else (1e-3 if u.startswith("K") else 1e-6)))


lines = ["# ncu --set full captures (B=512, N=45 sizes: 1 036 800 edge rows; tools/profile_one.py), one launch each, --clock-control none",
         "| kernel | time us | DRAM read MB | DRAM write MB | DRAM % of peak | tensor pipe % | SM % | warps active % | regs | L2 hit % | issue active % |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    name = r[col["Kernel Name"]][:70]
    tu = units[col["gpu__time_duration.sum"]]
    t = val(r, "gpu__time_duration.sum")
    t = t * 1000 if tu.startswith("ms") else (t / 1000 if tu.startswith("ns") else t)
    lines.append(f"| `{name}` | {t:.1f} | {mb(r, 'dram__bytes_read.sum'):.0f} | {mb(r, 'dram__bytes_write.sum'):.0f} | "
                 f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {r[col['launch__registers_per_thread']]} | "
                 f"{val(r, 'lts__t_sector_hit_rate.pct'):.1f} | {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} |")
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
