#!/bin/bash
# Round 2, final N-GPU run on the shipped tree: correctness + 1 KB..1 GB sweep, then every bench config.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/trips/r2_trip13.sh 8'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
OUT=gpurun_out/r2t13_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29971 tools/multigpu_check.py --specials --sections topo,correct,sweep --max-bytes 1073741824 \
   --out $OUT/check_default.json > $OUT/check_default.log 2>&1; echo "rc=$?"
grep -E "^\[correct|FAIL|Error" $OUT/check_default.log | head -6 | cut -c1-300
python - <<PY
import json
d = json.load(open("$OUT/check_default.json"))
print("correct", d.get("correct_all_pass"))
for r in d.get("sweep", []):
    print(" ", r["fp32_bytes"], {k[6:-3]: x for k, x in r.items() if k.startswith("fused_") and k.endswith("_us") and "phases" not in k}, "nccl", r.get("nccl_fp32_allreduce_us"), "ref", r.get("reference_algo_us"))
PY
for cfg in resnet18 resnet50 davidnet resnet18_quant; do
  timeout 300 $TR --master-port 29972 bench.py --gpus $N --steps 30 --warmup 5 --model $cfg > $OUT/bench_ours_$cfg.json 2> $OUT/bench_ours_$cfg.err; echo "rc=$?"
  python -c "
import json
d=json.loads(open('$OUT/bench_ours_$cfg.json').read().strip().splitlines()[-1]); print('$cfg', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])"
done
