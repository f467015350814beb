#!/bin/bash
# All 2^32 fp32 bit patterns through the compile-time casts and the pretruncate + IEEE identity (see the .cpp). No GPU.
set -eu
cd "$(dirname "$0")/.."
mkdir -p build/sanitize
g++ -O2 -std=c++17 -Wall -Wextra -Werror tests/cpp/exhaustive_cast_check.cpp -o build/sanitize/exhaustive_cast_check -lpthread
./build/sanitize/exhaustive_cast_check
