#!/bin/bash
# 1-GPU validation of the wide-guard fold: GPU kernel tests (CUDA fold == CPU fold bitwise) + fake-world W=8 check/timing
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip16; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fold or fused or allreduce or emulate" 2>&1 | tail -3
timeout 240 python tools/fake_world.py --out $OUT/fake8.json > $OUT/fake8.log 2>&1; echo "rc=$?"
grep "fake-world check" $OUT/fake8.log | cut -c1-600
python - <<PY
import json
d=json.load(open("$OUT/fake8.json"))
for r in d["fake_world"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")}, "C:", {k[:-10]: r[k]["C_fold"] for k in r if k.endswith("phases_us")})
PY
