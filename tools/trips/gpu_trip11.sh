#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/trip11_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check"; timeout 600 $TR --master-port 29911 tools/multigpu_check.py --sections correct,stress --stress-iters 600 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[correct|^\[stress|FAIL|ERROR" $OUT/check.log | head -12 | cut -c1-400
echo "== sweep"; timeout 600 $TR --master-port 29912 tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-67108864} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/sweep.json"))
for r in d["sweep"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")})
    print("    raw phases", r.get("fused_raw_phases_us"), "oneshot phases", r.get("fused_oneshot_phases_us"))
PY
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29913 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-330 $OUT/bench_ours.json
echo "== bench ours N=$N multi-bucket overlap"; timeout 400 $TR --master-port 29914 bench.py --gpus $N --steps 30 --warmup 5 --overlap --bucket-mb 4 > $OUT/bench_ours_overlap.json 2>> $OUT/bench_ours.err; cut -c1-330 $OUT/bench_ours_overlap.json
echo "== pytest gpu (1 GPU)"; timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
