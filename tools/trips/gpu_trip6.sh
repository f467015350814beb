#!/bin/bash
# N-GPU trip (N=8 or 4): fused all-reduce correctness + stress + sweep, bench both arms.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-8}
OUT=gpurun_out/trip6_n$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check"; NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR --master-port 29801 tools/multigpu_check.py --sections topo,symm,correct,stress --stress-iters 600 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[topo|^\[symm|^\[correct|^\[stress|FAIL|ERROR" $OUT/check.log | head -20 | cut -c1-600; grep -m3 -E "NVLS|nvls" $OUT/check.log | cut -c1-200
echo "== sweep"; timeout 600 $TR --master-port 29802 tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-1073741824} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; grep fp32_bytes $OUT/sweep.log | cut -c1-900
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29803 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-1300 $OUT/bench_ours.json; grep -E "Error|error" $OUT/bench_ours.err | head -5
echo "== bench reference N=$N"; timeout 900 $TR --master-port 29804 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-500 $OUT/bench_ref.json
echo "== bench ours N=$N overlap+buckets"; timeout 400 $TR --master-port 29805 bench.py --gpus $N --steps 20 --warmup 3 --overlap --bucket-mb 8 > $OUT/bench_ours_overlap.json 2>> $OUT/bench_ours.err; cut -c1-400 $OUT/bench_ours_overlap.json
for m in resnet50 davidnet; do
echo "== bench $m ours N=$N"; timeout 400 $TR --master-port 29806 bench.py --model $m --gpus $N --steps 20 --warmup 3 > $OUT/bench_${m}_ours.json 2>> $OUT/bench_ours.err; cut -c1-330 $OUT/bench_${m}_ours.json
echo "== bench $m reference N=$N"; timeout 600 $TR --master-port 29807 bench.py --model $m --impl reference --gpus $N --steps 6 --warmup 3 > $OUT/bench_${m}_ref.json 2>> $OUT/bench_ref.err; cut -c1-330 $OUT/bench_${m}_ref.json
done
ls $OUT
