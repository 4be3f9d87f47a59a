"""TEST INFRASTRUCTURE.  Builds tests/cpu_emul/_build/libpmp_simt_emul.so: the kernel and C-ABI sources
of plainmp_b200/csrc/ compiled with g++ against the SIMT emulator of cuda_emul.h.  The sources are
used as they are, except for two purely syntactic rewrites that g++ needs:
    K<<<grid, block, smem, stream>>>(args)   ->   K % emu::Cfg(grid, block, smem, stream)(args)
    extern __shared__ T sm[];                ->   extern T sm[];
The product never loads this library (plainmp_b200/_capi.py knows nothing about it)."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "plainmp_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libpmp_simt_emul.so"


def _rewrite(text: str) -> str:
    text = text.replace("extern __shared__", "extern")
    text = text.replace("<<<", " % emu::Cfg(").replace(">>>", ")")
    text = text.replace('#include "../../include/plainmp_b200.h"', f'#include "{ROOT / "include" / "plainmp_b200.h"}"')
    return text


def build(force: bool = False, sanitize: str | None = None) -> Path:
    """sanitize="address": the memcheck variant (tools/asan_emul.sh) — device allocations without slack,
    unrequested dynamic shared memory poisoned; load it with libasan preloaded."""
    import os

    sanitize = sanitize or os.environ.get("PMP_EMUL_SANITIZE") or None
    LIB = OUT / (f"libpmp_simt_emul_{sanitize}.so" if sanitize else "libpmp_simt_emul.so")
    srcs = sorted(CSRC.glob("*.cu*")) + [CSRC / "pmp_cloud_build.h"]
    deps = srcs + [HERE / "cuda_emul.h", HERE / "emul_extra.cpp", HERE / "stubs" / "cub" / "device" / "device_radix_sort.cuh", Path(__file__)]
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return LIB
    OUT.mkdir(exist_ok=True)
    for p in srcs:
        name = p.name.replace(".cu", ".cpp") if p.suffix == ".cu" else p.name
        (OUT / name).write_text(_rewrite(p.read_text()))
    cmd = [
        "g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-ffp-contract=off",
        f"-I{HERE}", f"-I{HERE / 'stubs'}", f"-I{OUT}",
        "-o", str(LIB), str(OUT / "plainmp_b200.cpp"), str(OUT / "pmp_cloud_build.cpp"), str(HERE / "emul_extra.cpp"),
    ]
    if sanitize:
        cmd[1:1] = [f"-fsanitize={sanitize}", "-g", "-fno-omit-frame-pointer"] + (["-DPMP_EMUL_ASAN"] if "address" in sanitize else [])
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
