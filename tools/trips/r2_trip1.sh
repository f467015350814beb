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
echo "== bench ours N=1 overlap"; timeout 300 python bench.py --steps 30 --warmup 5 --overlap > $OUT/bench_ours_overlap.json 2>> $OUT/bench_ours.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ours_overlap.json
echo "== staging tests"; timeout 600 python -m pytest tests/test_gpu_staging.py -q -m gpu > $OUT/pytest_staging.log 2>&1; echo "rc=$?"; tail -30 $OUT/pytest_staging.log
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_gpu.log
echo "== bench resnet18_quant ours"; timeout 300 python bench.py --model resnet18_quant --steps 10 --warmup 3 > $OUT/bench_quant_ours.json 2> $OUT/bench_quant_ours.err; echo "rc=$?"; cut -c1-500 $OUT/bench_quant_ours.json; tail -5 $OUT/bench_quant_ours.err
echo "== bench resnet18_quant reference"; timeout 600 python bench.py --model resnet18_quant --impl reference --steps 5 --warmup 3 > $OUT/bench_quant_ref.json 2> $OUT/bench_quant_ref.err; echo "rc=$?"; cut -c1-400 $OUT/bench_quant_ref.json; tail -5 $OUT/bench_quant_ref.err
echo "== microbench"; timeout 400 python tools/kernel_microbench.py > $OUT/microbench.json 2> $OUT/microbench.err; echo "rc=$?"; python -c "import json; d=json.load(open(\"$OUT/microbench.json\")); print({k: d[k] for k in (\"aps_fold\",\"lars\",\"gemm_tile_e2e\",\"operand_stage\")})"; tail -3 $OUT/microbench.err
