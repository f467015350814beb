#!/bin/bash
# 1-GPU trip: re-verify changed kernels, microbench, GEMM bench vs reference/cuBLAS, ncu captures, other models.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip4
mkdir -p $OUT
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== microbench"; timeout 300 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; python -c "
import json; d=json.load(open('$OUT/microbench.json'))
for k in ('aps_fold','lars'): [print(k, r) for r in d[k]]"
echo "== gemm bench"; timeout 900 python tools/bench_gemm.py > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; tail -3 $OUT/bench_gemm.err; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for k in ('quantize','gemm_exact','gemm_tc'): [print(k, r) for r in d[k]]"
echo "== ncu"; for k in gemm_tc fold lars; do
  pat="gemm_tcgen05_kernel"; [ $k = fold ] && pat="aps_fold_local_kernel"; [ $k = lars ] && pat="lars_kernel"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 3 -o $OUT/prof_$k python tools/ncu_targets.py $k > $OUT/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
for m in resnet50 davidnet; do
  echo "== bench $m ours"; timeout 600 python bench.py --model $m --gpus 1 --steps 20 --warmup 5 > $OUT/bench_${m}_ours.json 2> $OUT/bench_${m}_ours.err; cat $OUT/bench_${m}_ours.json | cut -c1-400; tail -2 $OUT/bench_${m}_ours.err
  echo "== bench $m reference"; timeout 900 python bench.py --model $m --impl reference --gpus 1 --steps 8 --warmup 3 > $OUT/bench_${m}_ref.json 2> $OUT/bench_${m}_ref.err; cat $OUT/bench_${m}_ref.json | cut -c1-400; grep -E "Error|error" $OUT/bench_${m}_ref.err | head -3
done
echo "== bench resnet18 ours (emulate_node 8, overlap, lars+kahan variants)"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --emulate-node 8 > $OUT/bench_r18_emul8.json 2>/dev/null; cut -c1-300 $OUT/bench_r18_emul8.json
ls $OUT | head -40
