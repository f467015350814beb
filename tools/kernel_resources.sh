#!/bin/bash
# Regenerates profiles/sass/RESOURCES.md (registers / smem / stack / spills of every kernel instantiation) from ptxas -v.
# No GPU needed.
set -eu
cd "$(dirname "$0")/.."
mkdir -p /tmp/cpd_ptx
for f in cpd_b200/csrc/cuda/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Icpd_b200/csrc -c $f -o /tmp/cpd_ptx/$b.o -Xptxas -v 2> /tmp/cpd_ptx/$b.log
done
python tools/kernel_resources.py /tmp/cpd_ptx profiles/sass/RESOURCES.md
