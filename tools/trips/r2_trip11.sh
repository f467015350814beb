#!/bin/bash
# Round 2, ONE GPU: 7-stage TMA ring, split-K deep conv layers, implicit conv default on - full suite, GEMM + quant + headline benches,
# ncu of the pair kernel at 8192^3.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t11; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "full suite rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python tools/bench_gemm.py --quick > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']: print({k: r[k] for k in r if k in ('M','N','K','acc','k_slice','variant','ms','TFLOPs','cublas_ms','cublas_TFLOPs')})"
timeout 400 python bench.py --model resnet18_quant --steps 20 --warmup 3 > $OUT/bench_quant_ours.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_quant_ours.json').read().strip().splitlines()[-1]); print('quant', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d.get('gpu_launches'))"
timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/bench_ours_1.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_1.json').read().strip().splitlines()[-1]); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta_kernel -s 1 -c 1 -o $OUT/prof_gemm_tcgen05_2cta_kernel python tools/ncu_targets.py gemm_tc2 > $OUT/ncu_gemm2.log 2>&1; echo "ncu gemm2 rc=$?"
python tools/summarize_profiles.py "r2t11" > $OUT/summarize.log 2>&1; grep -E "gpu__time_duration|pipe_tensor_cycles|dram__bytes_read.sum " profiles/ncu/r2t11_prof_gemm_tcgen05_2cta_kernel.md | head -4
