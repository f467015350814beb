#!/bin/bash
# Round 2, ONE GPU: persistent 1-CTA GEMM - full GPU suite, GEMM benchmark (quick), quant benchmark + launch shares.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t8; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "full suite rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python tools/bench_gemm.py --quick > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']: print({k: r[k] for k in r if k in ('M','N','K','acc','k_slice','variant','ms','TFLOPs','cublas_ms','cublas_TFLOPs')})"
timeout 400 python bench.py --model resnet18_quant --steps 20 --warmup 3 > $OUT/bench_quant_ours.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_quant_ours.json').read().strip().splitlines()[-1]); print('quant', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_quant.csv \
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 2 -c 1 -o $OUT/prof_gemm_tcgen05_kernel python tools/ncu_targets.py conv > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
python tools/summarize_profiles.py "r2t8" > $OUT/summarize.log 2>&1; grep -E "gpu__time_duration|pipe_tensor_cycles" profiles/ncu/r2t8_prof_gemm_tcgen05_kernel.md | head -3
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:bn_ --csv --log-file $OUT/bn_times.csv python tools/ncu_targets.py bn > $OUT/ncu_bn.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/bn_times.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mn, mv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
for r in rows[hdr + 1:][-16:]:
    if len(r) > mv: print(r[kn][:40], r[mn], r[mv])
PY
timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/bench_ours_1.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_1.json').read().strip().splitlines()[-1]); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
