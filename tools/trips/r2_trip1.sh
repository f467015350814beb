#!/bin/bash
# Round 2, trip 1 (ONE GPU): new tests, fake-world A/B of the protocol switches, ncu over smoke() as the driver runs it, bench N=1.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t1; mkdir -p $OUT
echo "== late GPU tests"; timeout 900 python -m pytest tests/test_zz_gpu_late.py -q -m gpu -x > $OUT/pytest_late.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_late.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -5 $OUT/smoke.log
echo "== ncu over smoke (driver-style launch list)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $OUT/launches_smoke.csv \
   python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.log 2>&1; echo "ncu_rc=$?"; tail -25 $OUT/ncu_smoke.log
grep -c "cpd::" $OUT/launches_smoke.csv; grep -o "cpd::[a-z_0-9]*" $OUT/launches_smoke.csv | sort | uniq -c
for v in default PACK_HW_OFF POLL_CTA0 WEAK_LOADS NO_EXTRA_FENCE; do
  echo "== fake world W=8, variant $v"
  case $v in default) E="X=1";; PACK_HW_OFF) E="CPD_B200_PACK_HW=0";; *) E="CPD_B200_$v=1";; esac
  env $E timeout 300 python tools/fake_world.py --max-bytes $((1<<26)) --out $OUT/fake8_$v.json > $OUT/fake8_$v.log 2>&1; echo "rc=$?"
  grep "fake-world check" $OUT/fake8_$v.log | cut -c1-400
done
python - <<PY
import json
for v in ("default", "PACK_HW_OFF", "POLL_CTA0", "WEAK_LOADS", "NO_EXTRA_FENCE"):
    try:
        d = json.load(open("$OUT/fake8_%s.json" % v))
    except Exception as e:
        print(v, "no result", e); continue
    for r in d["fake_world"]:
        ph = r.get("oneshot_phases_us", {})
        print(v, r["fp32_bytes"], "raw", r.get("raw_us"), "oneshot", r.get("oneshot_us"), "twoshot", r.get("twoshot_us"),
              "| A", ph.get("A_amax+gridbar"), "B", ph.get("B_pack+gridbar"), "flags", ph.get("exp_exchange"), ph.get("data_flags"), "C", ph.get("C_fold"))
PY
echo "== bench ours N=1"; timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-600 $OUT/bench_ours.json; tail -3 $OUT/bench_ours.err
echo "== bench ours N=1 steal"; timeout 300 python bench.py --steps 30 --warmup 5 --steal-grads > $OUT/bench_ours_steal.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ours_steal.json
echo "== bench ours N=1 overlap+steal+pipelined"; timeout 300 python bench.py --steps 30 --warmup 5 --overlap --steal-grads --pipelined > $OUT/bench_ours_all.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ours_all.json; python -c "import json; d=json.loads(open(\"$OUT/bench_ours_all.json\").read().strip().splitlines()[-1]); print(d[\"e2e\"])"
echo "== fused BN tests"; timeout 400 python -m pytest tests/test_gpu_fused_bn.py -q -m gpu > $OUT/pytest_bn.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_bn.log
echo "== bench ours N=1 fused-bn"; timeout 300 python bench.py --steps 30 --warmup 5 --fused-bn > $OUT/bench_ours_fusedbn.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ours_fusedbn.json
echo "== bench ours N=1 fused-bn+steal+pipelined"; timeout 300 python bench.py --steps 30 --warmup 5 --fused-bn --steal-grads --pipelined > $OUT/bench_ours_fusedbn_all.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-200 $OUT/bench_ours_fusedbn_all.json
echo "== bench ours N=1 overlap"; timeout 300 python bench.py --steps 30 --warmup 5 --overlap > $OUT/bench_ours_overlap.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ours_overlap.json
echo "== staging tests"; timeout 600 python -m pytest tests/test_gpu_staging.py -q -m gpu > $OUT/pytest_staging.log 2>&1; echo "rc=$?"; tail -30 $OUT/pytest_staging.log
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench resnet18_quant ours"; timeout 300 python bench.py --model resnet18_quant --steps 10 --warmup 3 > $OUT/bench_quant_ours.json 2> $OUT/bench_quant_ours.err; echo "rc=$?"; cut -c1-500 $OUT/bench_quant_ours.json; tail -5 $OUT/bench_quant_ours.err
echo "== bench resnet18_quant reference"; timeout 600 python bench.py --model resnet18_quant --impl reference --steps 5 --warmup 3 > $OUT/bench_quant_ref.json 2> $OUT/bench_quant_ref.err; echo "rc=$?"; cut -c1-400 $OUT/bench_quant_ref.json; tail -5 $OUT/bench_quant_ref.err
echo "== microbench"; timeout 400 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; echo "rc=$?"; python -c "import json; d=json.load(open(\"$OUT/microbench.json\")); print({k: d[k] for k in (\"aps_fold\",\"lars\",\"gemm_tile_e2e\",\"operand_stage\")})"; tail -3 $OUT/microbench.err
echo "== trainer logs (kept under profiles/trainers/)"
mkdir -p $OUT/trainers
timeout 300 python examples/ResNet18/mix.py --grad_exp 4 --grad_man 3 --use_APS --max-iter 40 --dataset_len 8192 --val-len 1024 --val_freq 20 --print_freq 10 --workers 2 --save_path $OUT/trainers/r18 > $OUT/trainers/resnet18_mix.log 2>&1; echo "mix rc=$?"; tail -4 $OUT/trainers/resnet18_mix.log | cut -c1-200
timeout 300 python examples/DavidNet/dawn.py --dist 0 --epoch 2 --dataset_len 8192 --grad_exp 5 --grad_man 2 --use_APS --use_kahan --use_lars --warm_up_epoch 1 > $OUT/trainers/davidnet_dawn.log 2>&1; echo "dawn rc=$?"; tail -5 $OUT/trainers/davidnet_dawn.log | cut -c1-200
timeout 400 python examples/ResNet50/main.py --grad_exp 5 --grad_man 2 --use-APS --max-iter 20 --dataset-len 640 --val-len 64 --epochs 1 --workers 2 --log-dir $OUT/trainers/r50_logs --checkpoint-format $OUT/trainers/r50_ck/checkpoint-{epoch}.pth.tar > $OUT/trainers/resnet50_main.log 2>&1; echo "main rc=$?"; tail -4 $OUT/trainers/resnet50_main.log | cut -c1-200
rm -rf $OUT/trainers/r18/*.pth $OUT/trainers/r50_ck
