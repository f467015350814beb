#!/bin/bash
# Round 2, ONE GPU: implicit-GEMM convolution - its tests (under a timeout: a wrong descriptor hangs on an mbarrier),
# then the Quant_Conv ResNet-18 with the path off / on, the launch shares with it on.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t9; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_implicit_conv.py -q -x > $OUT/pytest_implicit.log 2>&1; echo "implicit tests rc=$?"; tail -15 $OUT/pytest_implicit.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_staging.py -q -x > $OUT/pytest_gemm.log 2>&1; echo "gemm+staging tests rc=$?"; tail -3 $OUT/pytest_gemm.log
timeout 600 python tools/bench_gemm.py --quick > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']: print({k: r[k] for k in r if k in ('M','N','K','acc','k_slice','variant','ms','TFLOPs','cublas_ms','cublas_TFLOPs')})"
for imp in 0 1; do
  CPD_B200_CONV_IMPLICIT=$imp timeout 400 python bench.py --model resnet18_quant --steps 20 --warmup 3 > $OUT/bench_quant_imp$imp.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_quant_imp$imp.json').read().strip().splitlines()[-1]); print('quant implicit=$imp', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d.get('gpu_launches'))"
done
CPD_B200_CONV_IMPLICIT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_quant.csv \
   python bench.py --model resnet18_quant --steps 2 --warmup 3 --no-graph > $OUT/ncu_quant.log 2>&1; echo "ncu_rc=$?"
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches_quant.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > mv:
        try: agg[r[kn][:100]][0] += 1; agg[r[kn][:100]][1] += float(r[mv].replace(",", ""))
        except ValueError: pass
tot = sum(v[1] for v in agg.values())
print("total ms", tot / 1e6, "launches", sum(v[0] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print("%5.1f%% %5d %9.1f us  %s" % (100 * v[1] / tot, v[0], v[1] / 1e3, k))
PY
