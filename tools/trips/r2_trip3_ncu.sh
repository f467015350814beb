#!/bin/bash
# Round 2, ONE GPU: ncu --set full captures of the kernels touched this round (summaries go to profiles/ncu/).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t3; mkdir -p $OUT
export MASTER_ADDR=127.0.0.1
# fake world W = 8: a late launch of the 64 MB two-shot, of the 1 MB packed one-shot and of a 64 KB raw bucket
timeout 900 ncu --set full --clock-control none --import-source on --target-processes all -k regex:allreduce_fused_kernel -s 150 -c 3 \
   -o $OUT/prof_allreduce_fake8 python tools/ncu_targets.py allreduce_fake > $OUT/ncu_allreduce.log 2>&1; echo "ncu allreduce rc=$?"; tail -3 $OUT/ncu_allreduce.log
for k in fold:aps_fold_local_kernel lars:lars_kernel conv:operand_stage_kernel conv:gemm_tcgen05_kernel conv:splitk_fold_kernel bn:bn_stats_kernel bn:bn_apply_kernel bn:bn_bwd_stats_kernel bn:bn_bwd_apply_kernel; do
  t=${k%%:*}; pat=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 2 -o $OUT/prof_$pat python tools/ncu_targets.py $t > $OUT/ncu_$pat.log 2>&1; echo "ncu $pat rc=$?"
done
echo "== smoke under ncu (driver-style launch list)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $OUT/launches_smoke.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.log 2>&1; echo "ncu_rc=$?"; tail -3 $OUT/ncu_smoke.log; grep -o "cpd::[a-z_0-9]*" $OUT/launches_smoke.csv | sort | uniq -c
echo "== fused BN tests + microbench + benches"
timeout 300 python -m pytest tests/test_gpu_fused_bn.py tests/test_gpu_staging.py -q -m gpu > $OUT/pytest_bn_stage.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_bn_stage.log
timeout 400 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; echo "rc=$?"; python -c "import json; d=json.load(open(\"$OUT/microbench.json\")); print({k: d[k] for k in (\"gemm_tile_e2e\",\"operand_stage\")})"
for fl in "" "--fused-bn" "--fused-bn --steal-grads --pipelined"; do tag=$(echo "x$fl" | tr -d ' -'); timeout 300 python bench.py --steps 30 --warmup 5 $fl > $OUT/bench_ours_$tag.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_$tag.json').read().strip().splitlines()[-1]); print('$fl', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
timeout 300 python bench.py --model resnet18_quant --steps 10 --warmup 3 > $OUT/bench_quant_ours.json 2>> $OUT/bench.err; cut -c1-260 $OUT/bench_quant_ours.json
echo "== launch list of resnet18_quant steps (eager, serialised under ncu: compare SHARES)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_quant.csv \
   python bench.py --model resnet18_quant --steps 2 --warmup 3 --no-graph --fused-bn > $OUT/ncu_quant.log 2>&1; echo "ncu_rc=$?"; tail -2 $OUT/ncu_quant.log | cut -c1-200
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/launches_quant.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > mv:
        try: agg[r[kn][:110]][0] += 1; agg[r[kn][:110]][1] += float(r[mv].replace(",", ""))
        except ValueError: pass
tot = sum(v[1] for v in agg.values())
print("total ms", tot / 1e6, "launches", sum(v[0] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%5.1f%% %5d %9.1f us  %s" % (100 * v[1] / tot, v[0], v[1] / 1e3, k))
PY
python tools/summarize_profiles.py > $OUT/summarize.log 2>&1; tail -5 $OUT/summarize.log
ls -la $OUT | head -30
