#!/bin/bash
# trip 14: validate the rolled-loop all-reduce (correctness + stress + sweep with phase stamps) at N GPUs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/trip14_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check"; timeout 300 $TR --master-port 29931 tools/multigpu_check.py --sections correct,stress --stress-iters 300 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[correct|^\[stress|FAIL|ERROR" $OUT/check.log | head -12 | cut -c1-400
echo "== sweep"; timeout 300 $TR --master-port 29932 tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-67108864} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/sweep.json"))
for r in d["sweep"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")})
    print("    oneshot phases", r.get("fused_oneshot_phases_us"))
PY
