#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip5
mkdir -p $OUT
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== microbench"; timeout 300 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; python -c "
import json; d=json.load(open('$OUT/microbench.json'))
for k in ('aps_fold','lars'): [print(k, r) for r in d[k]]"
echo "== fused self-loop"; timeout 300 python tools/fused_selfloop.py > $OUT/selfloop.jsonl 2> $OUT/selfloop.err; cat $OUT/selfloop.jsonl; tail -2 $OUT/selfloop.err
echo "== ncu fused self-loop"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:allreduce_fused_kernel -s 2 -c 2 -o $OUT/prof_fused python tools/fused_selfloop.py --quick > $OUT/ncu_fused.log 2>&1; echo "rc=$?"
echo "== gemm bench"; timeout 900 python tools/bench_gemm.py > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; tail -3 $OUT/bench_gemm.err; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']:
    if r['M']==r['N'] and r['M']>=4096 or 'cublas_ms' in r: print(r)"
echo "== bench ours"; timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > $OUT/bench_ours.json 2>$OUT/bench_ours.err; cut -c1-330 $OUT/bench_ours.json
