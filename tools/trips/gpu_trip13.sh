#!/bin/bash
# Final N-GPU measurement: correctness + stress, sweep, bench ours + reference (ResNet-18).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
OUT=gpurun_out/trip13_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check"; timeout 600 $TR --master-port 29921 tools/multigpu_check.py --sections topo,correct,stress --stress-iters 500 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[topo|^\[correct|^\[stress|FAIL|ERROR" $OUT/check.log | head -12 | cut -c1-500
echo "== sweep"; timeout 600 $TR --master-port 29922 tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-1073741824} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/sweep.json"))
for r in d["sweep"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")})
print("phases at largest:", d["sweep"][-1].get("fused_twoshot_phases_us"), d["sweep"][-1].get("fused_nvls_phases_us"))
PY
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29923 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-1500 $OUT/bench_ours.json
echo "== bench reference N=$N"; timeout 900 $TR --master-port 29924 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-400 $OUT/bench_ref.json
