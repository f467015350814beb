#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip12
mkdir -p $OUT
for k in gemm_exact:gemm_exact_kernel mx:gemm_tcgen05_mx_kernel mx:mx_quantize_kernel quantize:quantize_kernel gemm_tc2:gemm_tcgen05_2cta_kernel; do
  t=${k%%:*}; pat=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o $OUT/prof_$pat python tools/ncu_targets.py $t > $OUT/ncu_$pat.log 2>&1; echo "ncu $pat rc=$?"
done
echo "== gemm bench"; timeout 900 python tools/bench_gemm.py > $OUT/bench_gemm.json 2> $OUT/bench_gemm.err; python -c "
import json; d=json.load(open('$OUT/bench_gemm.json'))
for r in d['gemm_tc']:
    if r['M']==8192 and r.get('variant','').startswith('2cta') or 'cublas_ms' in r and r['M']==8192: print(r)"
echo "== mx gemm bench"; timeout 300 python - <<'PY'
import torch, sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tools')
from kernel_microbench import time_ms
from cpd_b200.ops import native
res=[]
for n in (4096, 8192):
    a, bt = torch.randn(n, n, device='cuda'), torch.randn(n, n, device='cuda')
    ms_q = time_ms(lambda: native.mx_quantize(a), iters=5, warm=2)
    ac, asf = native.mx_quantize(a); bc, bsf = native.mx_quantize(bt)
    for (e, m, ks) in ((8, 23, 0), (5, 10, 1024), (5, 10, 256), (4, 3, 128)):
        ms = time_ms(lambda: native.gemm_tc_mx(ac, asf, bc, bsf, e, m, k_slice=ks), iters=5, warm=2, flush=False)
        res.append({"n": n, "acc": [e, m], "k_slice": ks or n, "ms": round(ms, 4), "TFLOPs": round(2 * n ** 3 / ms / 1e9, 1), "mx_quantize_ms": round(ms_q, 4),
                    "mx_quantize_GBps": round(n * n * 5.03 / ms_q / 1e6, 1)})
        print(res[-1])
json.dump(res, open('gpurun_out/trip12/bench_mx.json', 'w'), indent=1)
PY
