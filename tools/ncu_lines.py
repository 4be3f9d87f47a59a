#!/usr/bin/env python3
"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on (or its exported source page, .csv / .csv.gz).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if rep.endswith(".csv.gz"):  # exported on the GPU box: ncu -i rep --page source --csv --print-source cuda,sass | gzip
    import gzip
    raw = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = "?"; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ii = r.index("Instructions Executed"); ti = r.index("Thread Instructions Executed"); si = r.index("# Samples"); continue
    if hdr and len(r) > ii and r[0].strip().isdigit() and r[ii].isdigit():
        out.append((int(r[ii]), int(r[ti]) if r[ti].isdigit() else 0, int(r[si]) if r[si].isdigit() else 0, fname, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out); tots = sum(o[2] for o in out)
print(f"total warp instructions {tot}, samples {tots}")
byfile = {}
for o in out: byfile[o[3]] = byfile.get(o[3], 0) + o[0]
for f, n in sorted(byfile.items(), key=lambda x: -x[1]): print(f"  {f}: {100*n/tot:.1f}%")
out.sort(key=lambda x: -x[0])
for o in out[:top]:
    print(f"{100*o[0]/tot:5.1f}% inst {100*o[2]/max(tots,1):5.1f}% smpl  lanes {o[1]/max(o[0],1):4.1f}  {o[3]}:{o[4]}  {o[5]}")
# optional: per-region totals, regions given as file:lo-hi=name,...
if len(sys.argv) > 3:
    for spec in sys.argv[3].split(","):
        rng, name = spec.split("=")
        f, lh = rng.split(":"); lo, hi = map(int, lh.split("-"))
        n = sum(o[0] for o in out if o[3] == f and lo <= o[4] <= hi)
        ti = sum(o[1] for o in out if o[3] == f and lo <= o[4] <= hi)
        print(f"REGION {name:28s} {100*n/tot:5.1f}% inst  lanes {ti/max(n,1):4.1f}")
