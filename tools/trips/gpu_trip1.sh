#!/bin/bash
# First GPU trip: build check, smoke, GPU unit tests, 1-GPU bench (both arms), launch list + ncu of the top kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip1
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
echo "== build+smoke" ; timeout 900 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cat $OUT/bench_ours.json; tail -3 $OUT/bench_ours.err
echo "== bench ours (no graph, nchw variants)"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-graph > $OUT/bench_ours_nograph.json 2>> $OUT/bench_ours.err; cat $OUT/bench_ours_nograph.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --nchw > $OUT/bench_ours_nchw.json 2>> $OUT/bench_ours.err; cat $OUT/bench_ours_nchw.json
echo "== bench reference"; timeout 1500 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file $OUT/launches.csv python bench.py --gpus 1 --steps 3 --warmup 3 --no-graph > $OUT/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full on our kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aps_fold_local_kernel|lars_kernel|quantize_kernel" -c 6 -o $OUT/prof_ours python tools/kernel_microbench.py --quick > $OUT/ncu_full.log 2>&1; echo "rc=$?"
echo "== microbench"; timeout 600 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; cat $OUT/microbench.json | head -40
ls -la $OUT
