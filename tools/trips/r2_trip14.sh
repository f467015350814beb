#!/bin/bash
# Round 2, ONE GPU: two-level BatchNorm reduction - BN tests + the late file, BN kernel times, all four benches.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t14; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused_bn.py tests/test_zz_gpu_late.py -q > $OUT/pytest_bn_late.log 2>&1; echo "bn + late tests rc=$?"; tail -6 $OUT/pytest_bn_late.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:bn_ --csv --log-file $OUT/bn_times.csv python tools/ncu_targets.py bn > $OUT/ncu_bn.log 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/bn_times.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mn, mv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
for r in rows[hdr + 1:][-8:]:
    if len(r) > mv: print(r[kn][:40], r[mn], r[mv])
PY
for cfg in resnet18 resnet50 davidnet resnet18_quant; do
  timeout 300 python bench.py --steps 30 --warmup 5 --model $cfg > $OUT/bench_ours_$cfg.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_$cfg.json').read().strip().splitlines()[-1]); print('$cfg', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'].get('reasons'))"
done
