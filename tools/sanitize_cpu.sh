#!/bin/bash
# ASAN + UBSAN build of the C++ CPU numerics and their self-test, compiled -Wall -Wextra -Werror.
# No GPU needed.   usage: tools/sanitize_cpu.sh
set -eu
cd "$(dirname "$0")/.."
mkdir -p build/sanitize
g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -Wall -Wextra -Werror -Icpd_b200/csrc \
    cpd_b200/csrc/cpu/cpu_ops.cpp tests/cpp/cpu_ops_selftest.cpp -o build/sanitize/cpu_ops_selftest -lpthread
ASAN_OPTIONS=detect_leaks=1 UBSAN_OPTIONS=print_stacktrace=1 ./build/sanitize/cpu_ops_selftest
