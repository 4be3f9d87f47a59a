"""Branches of the kernel sources that the emulator suite never takes (gcov -b -c output of
tools/coverage_emul.sh, summed over template instantiations; exception edges ignored)."""
import re
import sys
from collections import defaultdict
from pathlib import Path

d, csrc = Path(sys.argv[1]), Path(__file__).resolve().parent.parent / "plainmp_b200" / "csrc"
for f in ("pmp_v3.cuh", "pmp_kernels.cuh", "pmp_cloud.cuh", "pmp_device.cuh"):
    src = (csrc / f).read_text().splitlines()
    cur, taken = None, defaultdict(lambda: defaultdict(int))
    for line in (d / f"{f}.gcov").read_text(errors="replace").splitlines():
        if line.startswith("branch"):
            m = re.match(r"branch\s+(\d+)\s+(?:taken (\d+)|never executed)", line)
            if m and cur and "(throw)" not in line:
                taken[cur][int(m.group(1))] += int(m.group(2)) if m.group(2) else 0
            continue
        if line.startswith(("call", "function", "-----")):
            continue
        m = re.match(r"\s*[^:]+:\s*(\d+):", line)
        if m:
            cur = int(m.group(1))
    n_br = sum(len(v) for v in taken.values())
    n_zero = sum(1 for v in taken.values() for c in v.values() if c == 0)
    print(f"\n{f}: {n_br - n_zero}/{n_br} branch outcomes taken on {len(taken)} branching lines")
    for ln in sorted(taken):
        if any(c == 0 for c in taken[ln].values()):
            print(f"  {ln}: {dict(taken[ln])}  {src[ln - 1].strip()[:100]}")
