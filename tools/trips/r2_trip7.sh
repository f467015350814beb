#!/bin/bash
# Round 2, ONE GPU, short: BatchNorm statistics kernels with round-robin row dealing - tests, ncu, headline bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t7; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_fused_bn.py -q -m gpu > $OUT/pytest_bn.log 2>&1; echo "bn tests rc=$?"; tail -3 $OUT/pytest_bn.log
for i in 1 2; do timeout 300 python bench.py --steps 40 --warmup 5 > $OUT/bench_ours_$i.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_$i.json').read().strip().splitlines()[-1]); print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['reasons'])"; done
timeout 300 python bench.py --steps 30 --warmup 5 --model resnet50 > $OUT/bench_ours_r50.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_r50.json').read().strip().splitlines()[-1]); print('r50', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 300 python bench.py --steps 30 --warmup 5 --model davidnet > $OUT/bench_ours_dn.json 2>> $OUT/bench.err; python -c "
import json; d=json.loads(open('$OUT/bench_ours_dn.json').read().strip().splitlines()[-1]); print('davidnet', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
for k in bn_stats_kernel bn_bwd_stats_kernel; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $OUT/prof_$k python tools/ncu_targets.py bn > $OUT/ncu_$k.log 2>&1; echo "ncu $k rc=$?"; done
python tools/summarize_profiles.py "r2t7" > $OUT/summarize.log 2>&1
for k in bn_stats_kernel bn_bwd_stats_kernel; do grep -E "gpu__time_duration|dram__bytes_read|dram_throughput" profiles/ncu/r2t7_prof_$k.md | head -3; done
