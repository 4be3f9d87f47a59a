#!/bin/bash
# ThreadSanitizer over the SIMT-emulated kernels (tests/cpu_emul/): one OS thread per CUDA thread, warp
# collectives are mutex/condvar rendezvous, so TSan sees exactly the happens-before edges that
# __syncwarp / __ballot_sync / __shfl_sync / __syncthreads give on the device.  A report means two lanes
# touch the same shared / global location without such an edge.
#   tools/tsan_emul.sh        (CPU only; needs the emulated sources built once by tests/cpu_emul/build.py)
set -e
cd "$(dirname "$0")/.."
python tests/cpu_emul/build.py > /dev/null
B=tests/cpu_emul/_build
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -w -ffp-contract=off -fsanitize=thread -Itests/cpu_emul -Itests/cpu_emul/stubs -I$B \
    -o $B/libpmp_simt_emul_tsan.so $B/plainmp_b200.cpp $B/pmp_cloud_build.cpp tests/cpu_emul/emul_extra.cpp
TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0" LD_PRELOAD=$(g++ -print-file-name=libtsan.so) \
    python tests/cpu_emul/tsan_run.py $B/libpmp_simt_emul_tsan.so
