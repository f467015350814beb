#!/bin/bash
# Round 2, ONE GPU: vectorised staging loads - tests, quant benchmark, launch shares, ncu of the staging kernel; full suite.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t6; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_staging.py tests/test_gpu_fused_bn.py -q -m gpu > $OUT/pytest_stage.log 2>&1; echo "staging+bn tests rc=$?"; tail -4 $OUT/pytest_stage.log
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "full suite rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 400 python tools/kernel_microbench.py --quick > $OUT/microbench.json 2> $OUT/microbench.err; python -c "
import json; d=json.load(open('$OUT/microbench.json'))
for k in ('gemm_tile_e2e','operand_stage'):
    for r in d[k]: print(k, r)"
for fl in "--model resnet18_quant --steps 15" ""; do tag=$(echo "x$fl" | tr -d ' -'); timeout 400 python bench.py --steps 30 --warmup 5 $fl > $OUT/bench_ours_$tag.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_$tag.json').read().strip().splitlines()[-1]); print('ours [$fl]', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"; done
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
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print("%5.1f%% %5d %9.1f us  %s" % (100 * v[1] / tot, v[0], v[1] / 1e3, k))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:operand_stage_kernel -s 2 -c 2 -o $OUT/prof_operand_stage_kernel python tools/ncu_targets.py conv > $OUT/ncu_stage.log 2>&1; echo "ncu stage rc=$?"
python tools/summarize_profiles.py "r2t6" > $OUT/summarize.log 2>&1; tail -2 $OUT/summarize.log
