#!/usr/bin/env python
"""Aggregate the per-SASS-instruction stall samples of an ncu report (--page source --csv) by CUDA source line.

    python tools/ncu_lines.py gpurun_out/chain.ncu-rep <launch index> druggen_b200/csrc/build/fused_mlp.o <mangled-kernel-substring> [top]

The .ncu-rep source page is SASS-level; the address -> file:line map comes from `nvdisasm -g` on the same object."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, idx, obj, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
main_file = os.path.basename(obj).replace(".o", ".cu")
line_of, stack, fresh, infn = {}, [], True, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if fresh:
            stack, fresh = [], False
        stack.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            stack.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and stack:
        fresh = True
        outer = [f for f in stack if f[0] == main_file]
        line_of[int(m.group(1), 16)] = ((outer[-1] if outer else stack[-1]), stack[0], m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg, total = {}, 0
base = None
for r in rows[2:]:
    try:
        addr = int(r[col["Address"]], 16)
    except Exception:
        continue
    if base is None:
        base = addr
    n = int(r[col["# Samples"]] or 0)
    total += n
    info = line_of.get(addr - base)
    key = (info[0][1], "" if info[1] == info[0] else f"{info[1][0]}:{info[1][1]}") if info else (0, "?")
    a = agg.setdefault(key, {"n": 0, "st": {}, "instr": 0})
    a["n"] += n
    a["instr"] += int(float(r[col["Instructions Executed"]] or 0))
    for s in stalls:
        v = int(r[col[s]] or 0)
        if v:
            a["st"][s] = a["st"].get(s, 0) + v
print(f"kernel launch {idx}: {total} samples")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:4])
    src = ""
    try:
        src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(obj))), main_file)).read().splitlines()[key[0] - 1].strip()[:90]
    except Exception:
        pass
    print(f"{100 * a['n'] / max(total, 1):5.1f}%  L{key[0]:<4d} {key[1]:<22s} instr={a['instr']:<9d} [{st}]  | {src}")
