#!/bin/bash
# Multi-GPU trip: fused all-reduce bring-up, correctness, stress, sweep, N-GPU bench (both arms).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/trip2_n$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
PORT=29511
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== tcgen05 gemm tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_tc or tile_mode" > $OUT/pytest_tc.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_tc.log
echo "== multigpu check: topo,symm,correct,stress"
NCCL_DEBUG=WARN timeout 900 $TR --master-port $PORT tools/multigpu_check.py --sections topo,symm,correct,stress --stress-iters 1500 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[|FAIL|ERROR|Error|error" $OUT/check.log | head -40; tail -5 $OUT/check.log
echo "== sweep"
timeout 900 $TR --master-port $((PORT+1)) tools/multigpu_check.py --sections sweep --max-bytes ${MAXB:-268435456} --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; grep fp32_bytes $OUT/sweep.log | tail -12; tail -3 $OUT/sweep.log
echo "== bench ours N=$N"
timeout 600 $TR --master-port $((PORT+2)) bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cat $OUT/bench_ours.json; tail -5 $OUT/bench_ours.err
echo "== bench ours N=$N nccl-gather mode"
timeout 600 $TR --master-port $((PORT+3)) bench.py --gpus $N --steps 20 --warmup 3 --algo nccl > $OUT/bench_ours_nccl.json 2>> $OUT/bench_ours.err; cat $OUT/bench_ours_nccl.json
echo "== bench reference N=$N"
timeout 900 $TR --master-port $((PORT+4)) bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
ls $OUT
