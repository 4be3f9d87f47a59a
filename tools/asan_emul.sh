#!/bin/bash
# AddressSanitizer over the SIMT-emulated kernels: the CPU-tier counterpart of compute-sanitizer's
# memcheck.  Device allocations carry no slack and the dynamic shared memory a launch did not request
# is poisoned, so an out-of-bounds global or shared access of any kernel is reported with the source
# line in plainmp_b200/csrc/.  Runs the whole emulator suite (tests/test_simt_emul.py), fuzz included.
#   tools/asan_emul.sh [pytest args]        (CPU only)
set -e
cd "$(dirname "$0")/.."
PMP_EMUL_SANITIZE=address python -c "import sys; sys.path.insert(0, 'tests'); from cpu_emul import build; print(build.build())"
PMP_EMUL_SANITIZE=address ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0" \
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) python -m pytest tests/test_simt_emul.py -x -q "$@"
# and UndefinedBehaviorSanitizer (shifts, signed overflow, misaligned or null accesses) on the same suite
PMP_EMUL_SANITIZE=undefined python -c "import sys; sys.path.insert(0, 'tests'); from cpu_emul import build; print(build.build())"
PMP_EMUL_SANITIZE=undefined LD_PRELOAD=$(g++ -print-file-name=libubsan.so) python -m pytest tests/test_simt_emul.py -x -q -s "$@" 2>&1 \
    | grep -E "runtime error|passed|failed" | sort | uniq -c
