#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip17; mkdir -p $OUT

timeout 240 python tools/fake_world.py --max-bytes $((1<<26)) --out $OUT/fake8.json > $OUT/fake8.log 2>&1; echo "rc=$?"
grep "fake-world check" $OUT/fake8.log | cut -c1-3000
python - <<PY
import json
d=json.load(open("$OUT/fake8.json"))
for r in d["fake_world"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")})
    for k in r:
        if k.endswith("phases_us") and "raw" not in k: print("    ", k[:-10], "C", r[k]["C_fold"], r[k].get("C_first_granule"))
PY
