#!/bin/bash
# Dumps per-kernel SASS of the built library and writes compact evidence files under profiles/sass/:
#   <kernel>.sass.txt   full SASS of the (4,3)/bf16 instantiation of each hand-written kernel
#   SUMMARY.md          mnemonic histograms (UTC*MMA / LDTM / UTMALDG / LDG.E.128 / multimem ...) + ptxas resources
set -euo pipefail
cd "$(dirname "$0")/.."
SO=$(ls cpd_b200/_C*.so | head -1)
OUT=profiles/sass
mkdir -p $OUT
cuobjdump -sass "$SO" > /tmp/all.sass
python - "$OUT" <<'PY'
import re, sys, collections, os, glob
out = sys.argv[1]
txt = open('/tmp/all.sass').read()
funcs = re.split(r'\n\s*Function : ', txt)[1:]
pick = {
 'quantize_kernel': r'quantize_kernel.*FmtTILi4ELi3',
 'amax_kernel': r'amax_kernel',
 'aps_fold_local_kernel': r'aps_fold_local_kernel.*FmtTILi4ELi3EEELb0',
 'lars_kernel': r'lars_kernel',
 'gemm_exact_kernel': r'gemm_exact_kernel.*FmtTILi4ELi3EEELb1',
 'gemm_tcgen05_kernel_bf16': r'gemm_tcgen05_kernelILi0ENS_8FmtHwF16ELb0',
 'gemm_tcgen05_kernel_fp8': r'gemm_tcgen05_kernelILi2ENS_9FmtHwE4M3ELb0',
 'gemm_tcgen05_2cta_kernel_bf16_f16acc': r'gemm_tcgen05_2cta_kernelILi0ENS_3tc26AccF16ELb0',
 'gemm_tcgen05_mx_kernel': r'gemm_tcgen05_mx_kernelINS_8FmtHwF16ELb0',
 'mx_quantize_kernel': r'mx_quantize_kernelILb0',
 'operand_stage_kernel_bf16': r'operand_stage_kernelILi0E',
 'bn_stats_kernel': r'bn_stats_kernel',
 'bn_apply_kernel_relu_res': r'bn_apply_kernelILb1ELb1E',
 'bn_bwd_stats_kernel_relu': r'bn_bwd_stats_kernelILb1E',
 'bn_bwd_apply_kernel_relu_res': r'bn_bwd_apply_kernelILb1ELb1E',
 'allreduce_fused_kernel_e4m3': r'allreduce_fused_kernel.*FmtTILi4ELi3EEELi8ELb0',
 'allreduce_fused_kernel_e5m2_kahan': r'allreduce_fused_kernel.*FmtTILi5ELi2EEELi8ELb1',
}
interesting = re.compile(r'\b(UTC[A-Z]*MMA[\w.]*|UTCBAR[\w.]*|UTCCP[\w.]*|F2FP[\w.]*|F2F\.[\w.]*|UCGABAR[\w.]*|UTCATOMSWS[\w.]*|LDTM[\w.]*|STTM[\w.]*|UTMALDG[\w.]*|UBLKCP[\w.]*|SYNCS[\w.]*|LDG\.E[\w.]*|STG\.E[\w.]*|LD\.E[\w.]*|ST\.E[\w.]*|MULTIMEM[\w.]*|LDGMC[\w.]*|STGMC[\w.]*|REDG?[\w.]*|ATOM[GS]?[\w.]*|MEMBAR[\w.]*|FENCE[\w.]*|CCTL[\w.]*|NANOSLEEP[\w.]*|BAR\.[\w.]*|LOP3[\w.]*|HMMA[\w.]*)')
lines = ["# SASS evidence (cuobjdump -sass of cpd_b200/_C*.so, sm_100a)\n"]
for name, pat in pick.items():
    f = [x for x in funcs if re.search(pat, x.split('\n', 1)[0])]
    if not f:
        lines.append(f"## {name}\n(not found: {pat})\n"); continue
    body = f[0]
    mangled = body.split('\n', 1)[0].strip()
    open(os.path.join(out, name + '.sass.txt'), 'w').write('Function : ' + body)
    ops = collections.Counter(m.group(1) for m in interesting.finditer(body))
    n_inst = len(re.findall(r'/\*[0-9a-f]{4}\*/', body))
    lines.append(f"## {name}\n`{mangled}` — {n_inst} instructions\n")
    for k, v in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0]))[:28]:
        lines.append(f"- `{k}` x{v}")
    lines.append("")
lines.append("## ptxas resource usage (from build/*.ptxas.txt)\n")
for p in sorted(glob.glob('build/cpd_b200/*.ptxas.txt')):
    t = open(p).read()
    regs = sorted(set(int(x) for x in re.findall(r'Used (\d+) registers', t)))
    spills = sorted(set(int(x) for x in re.findall(r'(\d+) bytes spill stores', t)))
    lines.append(f"- {os.path.basename(p).replace('.ptxas.txt','')}: registers {regs}, spill-store bytes {spills}")
open(os.path.join(out, 'SUMMARY.md'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines)[:3000])
PY
