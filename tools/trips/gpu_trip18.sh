#!/bin/bash
# N-GPU run after the wide-guard fold + fresh-buffer sweep: correctness, sweep (clean data), bench ours.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
OUT=gpurun_out/trip18_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check+sweep"; timeout 420 $TR --master-port 29941 tools/multigpu_check.py --sections correct,sweep --max-bytes ${MAXB:-1073741824} --out $OUT/check_sweep.json > $OUT/check_sweep.log 2>&1; echo "rc=$?"; grep -E "^\[correct|FAIL|ERROR|Error" $OUT/check_sweep.log | head -8 | cut -c1-400
python - <<PY
import json
d=json.load(open("$OUT/check_sweep.json"))
for r in d["sweep"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")}, [k for k,v in r.items() if k.endswith("finite_after") and not v])
    ph = r.get("fused_oneshot_phases_us") or r.get("fused_nvls_phases_us") or r.get("fused_twoshot_phases_us")
    print("    phases:", ph)
PY
echo "== bench ours N=$N"; timeout 300 $TR --master-port 29943 bench.py --gpus $N --steps 40 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-1200 $OUT/bench_ours.json
