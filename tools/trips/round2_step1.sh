#!/bin/bash
# Round-2 opener, ONE GPU (~4 min): everything that was added after round 1's GPU budget ran out.
#   gpurun --timeout 900 -- 'bash tools/trips/round2_step1.sh'
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2_step1; mkdir -p $OUT
echo "== new GPU tests"; timeout 600 python -m pytest tests/test_zz_gpu_late.py tests/test_gpu_kernels.py -q -m gpu \
  -k "fake_world or more_than_2_31 or overlapped or aps_fold or emulate" > $OUT/pytest_new.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_new.log
for v in default PACK_HW POLL_CTA0 WEAK_LOADS NO_EXTRA_FENCE; do
  echo "== fake world W=8, variant $v (check + timing)"
  env $( [ $v = default ] && echo X=1 || echo CPD_B200_$v=1 ) timeout 300 python tools/fake_world.py --max-bytes $((1<<26)) --out $OUT/fake8_$v.json > $OUT/fake8_$v.log 2>&1; echo "rc=$?"
  grep "fake-world check" $OUT/fake8_$v.log | cut -c1-300
done
python - <<PY
import json
for v in ("default", "PACK_HW", "POLL_CTA0", "WEAK_LOADS", "NO_EXTRA_FENCE"):
    try:
        d = json.load(open("$OUT/fake8_%s.json" % v))
    except Exception as e:
        print(v, "no result", e); continue
    for r in d["fake_world"]:
        ph = r.get("oneshot_phases_us", {})
        print(v, r["fp32_bytes"], "raw", r.get("raw_us"), "oneshot", r.get("oneshot_us"), "twoshot", r.get("twoshot_us"),
              "| B", ph.get("B_pack+gridbar"), "flags", ph.get("exp_exchange"), ph.get("data_flags"), "C", ph.get("C_fold"))
PY
echo "== full GPU suite"; timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== library swap: the reference's ResNet-18 recipe on its own CPDtorch and on our facade"
for lib in reference ours; do timeout 300 python tools/compat_run.py --lib $lib --model resnet18 --steps 6 2>$OUT/compat_$lib.err | tail -1 | tee $OUT/compat_$lib.json | cut -c1-300; done
