#!/bin/bash
# compute-sanitizer targets for every single-GPU kernel on small shapes: memcheck,
# racecheck, synccheck, initcheck.  Cross-GPU kernels are covered by the protocol stress test in
# tools/multigpu_check.py (sanitizers cannot see across processes).   usage: tools/sanitize.sh [tool ...]
set -u
cd "$(dirname "$0")/.."
TOOLS=${@:-memcheck racecheck synccheck initcheck}
mkdir -p gpurun_out/sanitize
for t in $TOOLS; do
  echo "== compute-sanitizer --tool $t"
  timeout 900 compute-sanitizer --tool $t --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -x -q \
    -k "quantize_variants or gemm_exact_matches_cpu and 17 or aps_fold_matches_cpu and 2 and 4-3 or lars_matches_cpu or gemm_tc_bf16 and 128-128-64" \
    > gpurun_out/sanitize/$t.log 2>&1
  echo "rc=$? ($(grep -c 'ERROR SUMMARY' gpurun_out/sanitize/$t.log) summaries)"; grep "ERROR SUMMARY" gpurun_out/sanitize/$t.log | sort | uniq -c | head -3
done
