#!/bin/bash
# 1-GPU trip: new GEMM kernels (2-CTA cluster, MX block-scaled), self-loop re-measure, sanitizer pass.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip7
mkdir -p $OUT
echo "== pytest: 2cta"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "2cta" > $OUT/pytest_2cta.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_2cta.log | cut -c1-300
echo "== pytest: mx"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "mx" > $OUT/pytest_mx.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_mx.log | cut -c1-300
echo "== pytest: all gpu"; timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gpu.log | cut -c1-300
echo "== fused self-loop"; timeout 300 python tools/fused_selfloop.py > $OUT/selfloop.jsonl 2> $OUT/selfloop.err; cat $OUT/selfloop.jsonl
echo "== gemm bench"; timeout 900 python tools/bench_gemm.py > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; tail -3 $OUT/bench_gemm.err; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']:
    if (r['M']==r['N'] and r['M']>=4096) or 'cublas_ms' in r: print(r)"
echo "== ncu 2cta"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta_kernel -s 1 -c 2 -o $OUT/prof_gemm_2cta python tools/ncu_targets.py gemm_tc2 > $OUT/ncu_2cta.log 2>&1; echo "rc=$?"
echo "== sanitize (memcheck, racecheck)"; bash tools/sanitize.sh memcheck racecheck 2>&1 | tail -8
cp -r gpurun_out/sanitize $OUT/ 2>/dev/null
ls $OUT
