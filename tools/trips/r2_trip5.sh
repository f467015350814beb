#!/bin/bash
# Round 2, N GPUs, short: correctness + sweep + headline bench after the peer-load fix (and W = 4 coverage).
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/trips/r2_trip5.sh 8'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
OUT=gpurun_out/r2t5_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MAXB=${MAXB:-268435456}
timeout 500 $TR --master-port 29961 tools/multigpu_check.py --specials --sections topo,correct,stress,sweep,fault --stress-iters ${STRESS:-3000} \
   --max-bytes $MAXB --out $OUT/check_default.json > $OUT/check_default.log 2>&1; echo "rc=$?"
grep -E "^\[correct|^\[stress|^\[fault|FAIL|Error" $OUT/check_default.log | head -8 | cut -c1-300
python - <<PY
import json
d = json.load(open("$OUT/check_default.json"))
print("correct", d.get("correct_all_pass"), "stress", d.get("stress_all_pass"), "fault", d.get("fault_all_pass"))
for r in d.get("sweep", []):
    print(" ", r["fp32_bytes"], {k[6:-3]: x for k, x in r.items() if k.startswith("fused_") and k.endswith("_us") and "phases" not in k}, "nccl", r.get("nccl_fp32_allreduce_us"), "ref", r.get("reference_algo_us"))
    if r["fp32_bytes"] in (16384, 1048576, 67108864):
        print("    raw phases", {k: v for k, v in (r.get("fused_raw_phases_us") or {}).items() if v is not None})
        print("    nvls phases", {k: x for k, x in (r.get("fused_nvls_phases_us") or r.get("fused_oneshot_phases_us") or {}).items() if k != "C_first_granule"}, (r.get("fused_nvls_phases_us") or {}).get("C_first_granule"))
PY
for cfg in "resnet18:" ${EXTRA_CFGS:-}; do
  m=${cfg%%:*}; fl=${cfg#*:}; tag=$(echo "$m$fl" | tr -d ' -')
  timeout 400 $TR --master-port 29962 bench.py --gpus $N --steps 40 --warmup 5 --model $m $fl > $OUT/bench_ours_$tag.json 2> $OUT/bench_ours_$tag.err; echo "rc=$?"
  python -c "
import json
d=json.loads(open('$OUT/bench_ours_$tag.json').read().strip().splitlines()[-1]); print('$m $fl', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])"
done
