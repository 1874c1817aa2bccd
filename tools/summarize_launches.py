#!/usr/bin/env python
"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel:
launch count, total time and share.  Under ncu the per-launch times are cold-cache and serialised, so only the SHARES
are comparable with the CUDA-event numbers of bench.py.

    python tools/summarize_launches.py gpurun_out/launches_r01b.csv profiles/r01_launches_summary.csv "header note"
"""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    ns = v * 1e3 if unit.startswith("us") else (v * 1e6 if unit.startswith("ms") else v)
    short = re.sub(r"\(.*", "", r[ki])[:80]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
    tot += ns
out = [f"# ncu launch list summary: {note}", "# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes",
       f"# launches {sum(a[0] for a in agg.values())}, total {tot / 1e6:.1f} ms", "kernel,launches,total_ms,share"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f'"{k}",{n},{ns / 1e6:.3f},{ns / tot:.4f}')
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
