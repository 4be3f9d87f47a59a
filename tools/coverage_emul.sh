#!/bin/bash
# Line coverage of the KERNEL SOURCES (plainmp_b200/csrc/) by the emulator suite: the sources are compiled
# for the SIMT emulator with gcov instrumentation (line numbers are those of csrc/, the two syntactic
# rewrites stay on their lines), tests/test_simt_emul.py runs, and gcov reports which kernel lines no
# CPU-tier test ever executes.   tools/coverage_emul.sh > profiles/<round>_emul_line_coverage.txt
set -e
cd "$(dirname "$0")/.."
python tests/cpu_emul/build.py > /dev/null
B=tests/cpu_emul/_build
C=$B/cov; rm -rf $C; mkdir -p $C
for f in plainmp_b200 pmp_cloud_build; do
  g++ --coverage -O0 -g -std=c++17 -fPIC -pthread -w -ffp-contract=off -Itests/cpu_emul -Itests/cpu_emul/stubs -I$B -c $B/$f.cpp -o $C/$f.o
done
g++ --coverage -O0 -std=c++17 -fPIC -pthread -w -Itests/cpu_emul -Itests/cpu_emul/stubs -I$B -c tests/cpu_emul/emul_extra.cpp -o $C/emul_extra.o
g++ --coverage -shared -pthread -o $B/libpmp_simt_emul_cov.so $C/plainmp_b200.o $C/pmp_cloud_build.o $C/emul_extra.o
PMP_EMUL_LIB=$PWD/$B/libpmp_simt_emul_cov.so python -m pytest tests/test_simt_emul.py -x -q "$@" 1>&2
mkdir -p $C/out
gcov -b -c -o $C $C/plainmp_b200.o $C/pmp_cloud_build.o > $C/gcov.log 2>&1 || true
mv ./*.gcov $C/out/
python tools/coverage_report.py $C/out
python tools/coverage_branches.py $C/out
