#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/trip3_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29700
echo "== topo_single"; timeout 120 python tools/debug_mg.py topo_single > $OUT/topo_single.log 2>&1; echo "rc=$?"; tail -4 $OUT/topo_single.log | cut -c1-600
for st in nccl_only dist_init symm_vmm symm_torch symm_ipc ar_vmm ar_torch ar_ipc; do
  P=$((P+1))
  echo "== $st"; timeout 180 $TR --master-port $P tools/debug_mg.py $st > $OUT/$st.log 2>&1; echo "rc=$?"; grep -E "ok|rank|done|Error|error|Fatal|File|line" $OUT/$st.log | grep -v "error_file\|Setting OMP" | head -14 | cut -c1-400
done
echo "== full check"
timeout 900 $TR --master-port 29750 tools/multigpu_check.py --sections topo,symm,correct,stress --stress-iters 1500 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[|FAIL|ERROR" $OUT/check.log | head -30 | cut -c1-500
echo "== sweep"
timeout 900 $TR --master-port 29751 tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-268435456} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; grep fp32_bytes $OUT/sweep.log | tail -12 | cut -c1-700
echo "== bench ours N=$N"
timeout 600 $TR --master-port 29752 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cat $OUT/bench_ours.json | cut -c1-1800; grep -E "Error|error|File" $OUT/bench_ours.err | head -8
echo "== bench ours N=$N nccl-gather mode"
timeout 600 $TR --master-port 29753 bench.py --gpus $N --steps 20 --warmup 3 --algo nccl > $OUT/bench_ours_nccl.json 2>> $OUT/bench_ours.err; cat $OUT/bench_ours_nccl.json | cut -c1-600
