#!/bin/bash
# Round 2, ONE GPU: validation of the final tree - full GPU suite, smoke (also under ncu), microbench, headline bench with the
# shipped defaults and with every engine option off, the Quant_Conv benchmark + its launch shares, reference arms.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t4; mkdir -p $OUT
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== fused BN with the 1-bit ReLU mask"; CPD_B200_BN_MASK=1 timeout 400 python -m pytest tests/test_gpu_fused_bn.py -q -m gpu > $OUT/pytest_bn_mask.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_bn_mask.log
CPD_B200_BN_MASK=1 timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/bench_ours_bnmask.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_bnmask.json').read().strip().splitlines()[-1]); print('ours [BN mask]', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/smoke.log
echo "== smoke under ncu"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $OUT/launches_smoke.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.log 2>&1; echo "ncu_rc=$?"; grep -o "cpd::[<>a-z_0-9:]*" $OUT/launches_smoke.csv | sort | uniq -c
echo "== microbench"; timeout 400 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('$OUT/microbench.json'))
for k in ('aps_fold','lars','gemm_tile_e2e','operand_stage'):
    for r in d[k]: print(k, r)"
for fl in "" "--no-fused-bn --no-steal-grads --no-pipelined" "--no-fused-bn" "--no-steal-grads" "--model davidnet" "--model resnet50" "--model resnet18_quant --steps 15"; do
  tag=$(echo "x$fl" | tr -d ' -'); timeout 400 python bench.py --steps 30 --warmup 5 $fl > $OUT/bench_ours_$tag.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_$tag.json').read().strip().splitlines()[-1]); print('ours [$fl]', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks']['reasons'])"; done
for m in resnet18 davidnet resnet50; do timeout 400 python bench.py --impl reference --steps 10 --warmup 3 --model $m > $OUT/bench_ref_$m.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ref_$m.json').read().strip().splitlines()[-1]); print('reference $m', d.get('value'), d.get('ms_per_step'), d.get('unavailable'))"; done
echo "== launch shares of the quant benchmark (eager, under ncu)"
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
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%5.1f%% %5d %9.1f us  %s" % (100 * v[1] / tot, v[0], v[1] / 1e3, k))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:operand_stage_kernel -s 2 -c 2 -o $OUT/prof_operand_stage_kernel python tools/ncu_targets.py conv > $OUT/ncu_stage.log 2>&1; echo "ncu stage rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_stats_kernel -s 2 -c 1 -o $OUT/prof_bn_stats_kernel python tools/ncu_targets.py bn > $OUT/ncu_bn1.log 2>&1; echo "ncu bn_stats rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_stats_kernel -s 2 -c 1 -o $OUT/prof_bn_bwd_stats_kernel python tools/ncu_targets.py bn > $OUT/ncu_bn2.log 2>&1; echo "ncu bn_bwd_stats rc=$?"
python tools/summarize_profiles.py "r2t4" > $OUT/summarize.log 2>&1; tail -3 $OUT/summarize.log
