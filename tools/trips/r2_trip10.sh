#!/bin/bash
# Round 2, ONE GPU: implicit weight gradient (MN-major operand tiles) - descriptor probe, tests, the Quant_Conv ResNet-18.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t10; mkdir -p $OUT
timeout 120 python tools/probe_wgrad_desc.py 2>&1 | tee $OUT/probe_wgrad.log | tail -8
timeout 300 python -m pytest tests/test_gpu_implicit_conv.py -q > $OUT/pytest_implicit.log 2>&1; echo "implicit tests rc=$?"; tail -12 $OUT/pytest_implicit.log
CPD_B200_CONV_IMPLICIT=1 timeout 300 python -m pytest tests/test_gpu_staging.py -q > $OUT/pytest_staging_implicit.log 2>&1; echo "staging tests with implicit on rc=$?"; tail -3 $OUT/pytest_staging_implicit.log
for imp in 1; do
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
