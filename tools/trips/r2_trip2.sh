#!/bin/bash
# Round 2, N GPUs (2 to develop, 8 to confirm): correctness (+specials, fault injection), stress on the shipped kernel,
# sweep incl. in-switch modes, calibration, training benchmarks with and without backward overlap.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/trips/r2_trip2.sh 2'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
MODE=${2:-full}          # full | quick (8 GPUs: keep the box time short)
OUT=gpurun_out/r2t2_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29950
STRESS=${STRESS:-10000}
MAXB=${MAXB:-1073741824}
echo "== correct + specials + stress($STRESS) + fault + sweep (default protocol)"
timeout 900 $TR --master-port $port tools/multigpu_check.py --specials --sections topo,correct,stress,sweep,fault --stress-iters $STRESS \
   --max-bytes $MAXB --out $OUT/check_default.json > $OUT/check_default.log 2>&1; echo "rc=$?"
grep -E "^\[correct|^\[stress|^\[fault|^\[topo|FAIL|Error" $OUT/check_default.log | head -12 | cut -c1-400
if [ "$MODE" = full ]; then
for v in NO_EXTRA_FENCE WEAK_LOADS BOTH; do
  port=$((port + 1))
  case $v in PACK_HW_OFF) E="CPD_B200_PACK_HW=0";; BOTH) E="CPD_B200_WEAK_LOADS=1 CPD_B200_NO_EXTRA_FENCE=1";; *) E="CPD_B200_$v=1";; esac
  echo "== variant $v: check + sweep to 64 MB"
  env $E timeout 420 $TR --master-port $port tools/multigpu_check.py --sections correct,sweep --max-bytes $((1<<26)) \
     --out $OUT/check_$v.json > $OUT/check_$v.log 2>&1; echo "rc=$?"
  grep -E "^\[correct|FAIL|Error" $OUT/check_$v.log | head -4 | cut -c1-300
done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/check_*.json")):
    v = f.split("check_")[-1][:-5]
    try:
        d = json.load(open(f))
    except Exception as e:
        print(v, "no result", e); continue
    print(v, "correct", d.get("correct_all_pass"), "stress", d.get("stress_all_pass"), "fault", d.get("fault_all_pass"), "error" in d)
    for r in d.get("sweep", []):
        print(" ", v, r["fp32_bytes"], {k[6:-3]: x for k, x in r.items() if k.startswith("fused_") and k.endswith("_us") and "phases" not in k},
              "nccl", r.get("nccl_fp32_allreduce_us"), "ref", r.get("reference_algo_us"))
        if v == "default" and r["fp32_bytes"] in (16384, 1048576, 67108864):
            print("    raw phases", r.get("fused_raw_phases_us"))
            print("    nvls phases", {k: x for k, x in (r.get("fused_nvls_phases_us") or r.get("fused_oneshot_phases_us") or {}).items() if k != "C_first_granule"})
PY
echo "== calibration"; port=$((port + 1))
timeout 200 $TR --master-port $port tools/calibrate_run.py > $OUT/calibrate.log 2>&1; echo "rc=$?"; grep -E "calibrate|thresholds" $OUT/calibrate.log | tail -8 | cut -c1-300
if [ "$MODE" = full ]; then
  CFGS=("resnet18:" "resnet18:--pipelined" "resnet18:--overlap" "resnet18:--overlap --comm-ctas 148 --bucket-mb 24" "resnet18:--overlap --comm-ctas 74 --bucket-mb 24" "davidnet:" "resnet50:" "resnet18_quant:")
else
  CFGS=("resnet18:" "resnet18:--overlap --comm-ctas 148 --bucket-mb 24" "resnet50:" "davidnet:" "resnet18_quant:")
fi
for cfg in "${CFGS[@]}"; do
  m=${cfg%%:*}; fl=${cfg#*:}; tag=$(echo "$m$fl" | tr -d ' -')
  port=$((port + 1))
  echo "== bench ours $m $fl N=$N"
  timeout 400 $TR --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --model $m $fl > $OUT/bench_ours_$tag.json 2> $OUT/bench_ours_$tag.err; echo "rc=$?"
  python -c "
import json,sys
try:
    d=json.loads(open('$OUT/bench_ours_$tag.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['impl_detail']['buckets'], d['clocks'])
except Exception as e: print('no result', e)"; tail -2 $OUT/bench_ours_$tag.err | cut -c1-300
done
if [ "$MODE" = full ]; then
for m in resnet18 resnet18_quant; do
  port=$((port + 1))
  echo "== bench reference $m N=$N"
  timeout 600 $TR --master-port $port bench.py --gpus $N --steps 8 --warmup 3 --model $m --impl reference > $OUT/bench_ref_$m.json 2> $OUT/bench_ref_$m.err; echo "rc=$?"; cut -c1-250 $OUT/bench_ref_$m.json
done
fi
if [ "${SWEEPBENCH:-0}" = 1 ]; then
echo "== bench.py --allreduce-sweep"; port=$((port + 1))
timeout 400 $TR --master-port $port bench.py --gpus $N --allreduce-sweep --sweep-max-bytes $((1<<28)) > $OUT/bench_sweep.json 2> $OUT/bench_sweep.err; echo "rc=$?"; cut -c1-300 $OUT/bench_sweep.json
fi
