#!/bin/bash
# Round-2, N GPUs: correctness + sweep for the default kernel and the two opt-in variants, then the bench.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/trips/round2_step2.sh 2'      (then 8)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/r2_step2_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29950
for v in default PACK_HW POLL_CTA0 WEAK_LOADS NO_EXTRA_FENCE; do
  port=$((port + 1))
  echo "== variant $v: check + sweep"
  env $( [ $v = default ] && echo X=1 || echo CPD_B200_$v=1 ) timeout 420 $TR --master-port $port tools/multigpu_check.py --specials --sections correct,sweep \
    --max-bytes ${MAXB:-268435456} --out $OUT/check_sweep_$v.json > $OUT/check_sweep_$v.log 2>&1; echo "rc=$?"
  grep -E "^\[correct|FAIL|Error" $OUT/check_sweep_$v.log | head -4 | cut -c1-300
done
python - <<PY
import json
for v in ("default", "PACK_HW", "POLL_CTA0", "WEAK_LOADS", "NO_EXTRA_FENCE"):
    try:
        d = json.load(open("$OUT/check_sweep_%s.json" % v))
    except Exception as e:
        print(v, "no result", e); continue
    for r in d["sweep"]:
        print(v, r["fp32_bytes"], {k[6:-3]: x for k, x in r.items() if k.startswith("fused_") and k.endswith("_us") and "phases" not in k}, "nccl", r.get("nccl_fp32_allreduce_us"))
        if r["fp32_bytes"] in (16384, 1048576):
            print("    raw phases", r.get("fused_raw_phases_us"), "\n    oneshot phases", {k: x for k, x in (r.get("fused_oneshot_phases_us") or {}).items() if k != "C_first_granule"})
PY
echo "== bench ours N=$N"; timeout 300 $TR --master-port 29970 bench.py --gpus $N --steps 40 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-400 $OUT/bench_ours.json
echo "== bench ours N=$N, pipelined e2e"; timeout 300 $TR --master-port 29971 bench.py --gpus $N --steps 40 --warmup 5 --pipelined > $OUT/bench_ours_pipelined.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-400 $OUT/bench_ours_pipelined.json
