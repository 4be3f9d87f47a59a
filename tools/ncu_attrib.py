#!/usr/bin/env python3
"""Dynamic warp-instruction counts of one kernel by PHASE: every SASS instruction of the kernel is attributed to
the pmp_v3.cuh function on its inline chain (`nvdisasm -gi` on the cubin of the shipped .so gives the chain,
the exported ncu source page gives the executed counts per SASS address), which the per-source-line view
cannot do (an inlined quat_rot is one line whoever called it).

  cuobjdump -xelf all plainmp_b200/libplainmp_b200.so          # -> plainmp_b200.sm_100a.cubin
  nvdisasm -gi -c plainmp_b200.sm_100a.cubin > all.dis
  python tools/ncu_attrib.py all.dis '<mangled kernel name>' gpurun_out/<tag>_prof_<workload>_sass.csv.gz <items>

The library must be the build that was profiled (same SASS); the script checks the opcodes line up.
"""
import csv
import gzip
import io
import re
import sys
from pathlib import Path

dis, fn, prof, nitems = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
SRC = Path(__file__).resolve().parent.parent / "plainmp_b200" / "csrc"


def functions(path):  # [(first line, name)] of the device functions of a header
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r"__device__ (?:__forceinline__|__noinline__|inline) .*?(\w+)\(", l)
        if m:
            out.append((i, m.group(1)))
    return out


FUN = {f: functions(SRC / f) for f in ("pmp_v3.cuh", "pmp_cloud.cuh", "pmp_device.cuh")}


def fname(f, ln):
    name = None
    for i, n in FUN.get(f, []):
        if i <= ln:
            name = n
    return name


lines = open(dis).read().split("\n")
i0 = lines.index(".text." + fn + ":")
i1 = next(i for i in range(i0 + 1, len(lines)) if lines[i].startswith("//--------------------- .text."))
chains, cur, pending = {}, [], []
for l in lines[i0:i1]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        if pending:
            cur, pending = pending, []
        chains[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(io.StringIO(gzip.open(prof, "rt").read())))
for i, r in enumerate(rows):
    if "Instructions Executed" in r:
        hdr, start = r, i
        break
ia, ii, ti, isrc, ism = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "Source", "# Samples"))
data = [(int(r[ia], 16), int(r[ii]), int(r[ti]) if r[ti].isdigit() else 0, r[isrc], int(r[ism]) if r[ism].isdigit() else 0)
        for r in rows[start + 1:] if len(r) > ti and r[ii].isdigit()]
base = data[0][0]


def opcode(txt):
    t = txt.split()
    return t[1] if t[0].startswith("@") else t[0]


mism = sum(1 for a, e, t, s, sm in data if (a - base) not in chains or opcode(chains[a - base][1]) != opcode(s))
print(f"{len(data)} SASS instructions, {mism} opcode mismatches between the profile and this build")
PHASES = ("v3_group_run", "v3_unit_test", "drain_self", "v3_cloud_flush")  # everything else is phase 1
tot = sum(d[1] for d in data)
tots = sum(d[4] for d in data)
agg, ph = {}, {}
for a, e, t, s, sm in data:
    ch = chains.get(a - base, ([], ""))[0]
    names = [fname(f, ln) for f, ln in ch if f in FUN]
    names = [n for n in names if n]
    phase = next((n for n in reversed(names) if n in PHASES), "phase 1")
    inner = names[0] if names else "-"
    for key, tab in ((phase, ph), (phase + " / " + inner, agg)):
        g = tab.setdefault(key, [0, 0, 0])
        g[0] += e
        g[1] += t
        g[2] += sm
print(f"{tot / nitems:.1f} warp instructions per item")
for tab, top in ((ph, 99), (agg, 30)):
    for k, v in sorted(tab.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{k:44s} {100 * v[0] / tot:5.1f}% inst {100 * v[2] / max(tots, 1):5.1f}% smpl {v[0] / nitems:7.1f} per item  lanes {v[1] / max(v[0], 1):4.1f}")
    print()
