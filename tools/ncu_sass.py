#!/usr/bin/env python3
"""Walk the SASS page of an ncu capture (exported with `ncu -i rep --page source --csv --print-source sass`,
.csv or .csv.gz): per instruction executed count, lanes, stall samples.  Prints the hottest basic blocks
(runs of instructions with the same executed count) so that phases of a kernel can be told apart even
where inlining folds them onto the same source lines.
  python tools/ncu_sass.py file_sass.csv.gz [min_share_pct]
"""
import csv, gzip, io, sys
f = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = gzip.open(f, "rt").read() if f.endswith(".gz") else open(f).read()
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ins = [(r[1].strip(), int(r[ii]), int(r[ti]), int(r[si])) for r in rows if r and r[0].startswith("0x")]
tot = sum(i[1] for i in ins); tots = sum(i[3] for i in ins)
print(f"{len(ins)} SASS instructions, {tot} warp instructions executed, {tots} samples")
# basic blocks by equal executed count
blocks = []; start = 0
for k in range(1, len(ins) + 1):
    if k == len(ins) or ins[k][1] != ins[start][1] or ins[k - 1][0].split()[0] in ("BRA", "EXIT", "RET", "BSYNC", "WARPSYNC") or ins[k-1][0].startswith("@") and " BRA" in ins[k-1][0]:
        blocks.append((start, k)); start = k
for a, b in blocks:
    n = sum(i[1] for i in ins[a:b]); t = sum(i[2] for i in ins[a:b]); s = sum(i[3] for i in ins[a:b])
    if 100 * n / tot >= thr:
        ops = {}
        for i in ins[a:b]:
            op = i[0].split()[1].split(".")[0] if i[0].startswith("@") else i[0].split()[0].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:8])
        print(f"[{a:5d}-{b:5d}) {b-a:4d} instr x {ins[a][1]:9d}  {100*n/tot:5.1f}% inst {100*s/max(tots,1):5.1f}% smpl lanes {t/max(n,1):4.1f}  {top}")
