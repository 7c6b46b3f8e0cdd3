#!/bin/bash
# AddressSanitizer + UBSan over the host builds of the code that also runs on the GPU (map generator, step
# kernel) and over the CPU oracle: 340 generator runs (six map configurations and every table-overflow path)
# and 350 steps of side-by-side rollouts.  Usage: tools/sanitize_host_builds.sh   (prints two "ok" lines, no reports)
set -e
cd "$(dirname "$0")/../oracle"
mkdir -p _san
FLAGS="-O1 -g -fPIC -ffp-contract=off -fno-math-errno -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas"
g++ $FLAGS -std=c++17 -shared -x c++ -o _san/libpgd_mapgen_host.so mapgen_host.cpp -lm
g++ $FLAGS -std=c++17 -shared -x c++ -o _san/libpgd_step_host.so step_host.cpp -lm
gcc $FLAGS -shared -o _san/libpgd_oracle.so pgd_oracle.c -lm
cd ..
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) python tools/sanitize_run.py
# the thread pool of pgd_step_host (pgdrive_b200/csrc/pgd_hostpool.h) under ThreadSanitizer and ASan / UBSan
g++ -std=c++17 -O1 -g -fsanitize=thread -o oracle/_san/hostpool_tsan oracle/hostpool_check.cpp -lpthread
oracle/_san/hostpool_tsan
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -o oracle/_san/hostpool_asan oracle/hostpool_check.cpp -lpthread
oracle/_san/hostpool_asan
