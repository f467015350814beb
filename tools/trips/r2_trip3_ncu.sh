#!/bin/bash
# Round 2, ONE GPU: ncu --set full captures of the kernels touched this round (summaries go to profiles/ncu/).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t3; mkdir -p $OUT
export MASTER_ADDR=127.0.0.1
# fake world W = 8: a late launch of the 64 MB two-shot, of the 1 MB packed one-shot and of a 64 KB raw bucket
timeout 900 ncu --set full --clock-control none --import-source on --target-processes all -k regex:allreduce_fused_kernel -s 150 -c 3 \
   -o $OUT/prof_allreduce_fake8 python tools/ncu_targets.py allreduce_fake > $OUT/ncu_allreduce.log 2>&1; echo "ncu allreduce rc=$?"; tail -3 $OUT/ncu_allreduce.log
for k in fold:aps_fold_local_kernel lars:lars_kernel conv:operand_stage_kernel conv:gemm_tcgen05_kernel conv:splitk_fold_kernel; do
  t=${k%%:*}; pat=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 2 -o $OUT/prof_$pat python tools/ncu_targets.py $t > $OUT/ncu_$pat.log 2>&1; echo "ncu $pat rc=$?"
done
python tools/summarize_profiles.py > $OUT/summarize.log 2>&1; tail -5 $OUT/summarize.log
ls -la $OUT | head -30
