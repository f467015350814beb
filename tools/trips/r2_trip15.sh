#!/bin/bash
# Round 2, ONE GPU, final validation of the shipped tree: full GPU suite, smoke (+ ncu launch list as the driver takes it),
# default bench line, convolution kernels vs cuDNN bf16, memcheck of the new kernels, FCN trainer log.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t15; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "full suite rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/launches_smoke.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
kn = rows[hdr].index("Kernel Name")
names = sorted({r[kn].split("(")[0][-60:] for r in rows[hdr + 1:] if len(r) > kn and "cpd" in r[kn]})
print(len(names), "distinct repo kernels under ncu:"); print("\n".join("  " + n for n in names))
PY
timeout 300 python bench.py > $OUT/bench_default.json 2>> $OUT/bench.err; tail -1 $OUT/bench_default.json | cut -c1-400
timeout 400 python tools/conv_bench.py 64 512 > $OUT/conv_bench.json 2> $OUT/conv_bench.err; echo "conv_bench rc=$?"; python -c "
import json; d=json.load(open('$OUT/conv_bench.json'))
for r in d['conv']: print(r)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_implicit_conv.py tests/test_gpu_fused_bn.py -x -q \
   -k "bitwise and g1 or wgrad and g0 or split_k or fused_bn_act_matches_stock_ops and shape0 or quant_conv_implicit and cfg1" > $OUT/memcheck_new.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY" $OUT/memcheck_new.log | sort | uniq -c | head -3; tail -2 $OUT/memcheck_new.log
timeout 300 python examples/FCN/train_fcn.py --arch fcn_r50_d8 --crop 513 --batch-size 2 --max-iters 20 --dataset-len 64 --val-len 4 --val-freq 20 --print-freq 5 --workers 2 --use_APS > $OUT/fcn_train.log 2>&1; echo "fcn rc=$?"; tail -4 $OUT/fcn_train.log
