#!/usr/bin/env python3
"""Per-line listing (in line order) of one source file from an .ncu-rep, as warp-instr per item.
  python tools/ncu_lines_by_file.py rep file.cuh lo hi n_items
"""
import csv, io, subprocess, sys
rep, target, lo, hi, n_items = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = "?"; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ii = r.index("Instructions Executed"); ti = r.index("Thread Instructions Executed"); si = r.index("# Samples"); continue
    if hdr and len(r) > ii and r[0].strip().isdigit() and r[ii].isdigit():
        out.append((int(r[ii]), int(r[ti]) if r[ti].isdigit() else 0, int(r[si]) if r[si].isdigit() else 0, fname, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out)
tots = sum(o[2] for o in out)
acc = 0
for o in sorted(out, key=lambda x: x[4]):
    if o[3] == target and lo <= o[4] <= hi and o[0] > 0:
        acc += o[0]
        print(f"{o[0]/n_items:6.2f} wipc {100*o[2]/tots:4.1f}%smp lanes {o[1]/max(o[0],1):4.1f}  {o[4]:4d}  {o[5]}")
print(f"sum {acc/n_items:.1f} wipc of {tot/n_items:.1f}")
