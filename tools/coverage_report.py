"""Summarises the .gcov files tools/coverage_emul.sh leaves behind: per source file of plainmp_b200/csrc/
the executable lines, the share the emulator suite executed, and the uncovered line ranges."""
import sys
from pathlib import Path

d = Path(sys.argv[1])
rows = []
for g in sorted(d.glob("*.gcov")):
    name = g.name[: -len(".gcov")]
    if not (name.startswith("pmp_") or name.startswith("plainmp_b200")):
        continue
    # templates: gcov repeats a line once per instantiation; a line counts as executed when any did
    seen = {}
    for line in g.read_text(errors="replace").splitlines():
        parts = line.split(":", 2)
        if len(parts) < 3:
            continue
        cnt, no = parts[0].strip(), parts[1].strip()
        if not no.isdigit() or int(no) == 0 or cnt == "-":
            continue
        ran = not (cnt.startswith("#####") or cnt.startswith("====="))
        seen[int(no)] = seen.get(int(no), False) or ran
    hit = sum(1 for v in seen.values() if v)
    miss = len(seen) - hit
    missed = sorted(n for n, v in seen.items() if not v)
    ranges, start, prev = [], None, None
    for n in missed:
        if start is None:
            start = prev = n
        elif n <= prev + 2:
            prev = n
        else:
            ranges.append((start, prev))
            start = prev = n
    if start is not None:
        ranges.append((start, prev))
    rows.append((name, hit, miss, ranges))
print("kernel-source line coverage of the emulator suite (tests/test_simt_emul.py), gcov on the emulated build")
for name, hit, miss, ranges in rows:
    tot = hit + miss
    print(f"\n{name}: {hit}/{tot} executable lines = {100.0 * hit / max(tot, 1):.1f} %")
    if ranges:
        print("  not executed: " + ", ".join(f"{a}" if a == b else f"{a}-{b}" for a, b in ranges))
