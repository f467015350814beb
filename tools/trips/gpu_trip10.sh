#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
OUT=gpurun_out/trip10_n$N
mkdir -p $OUT /tmp/work
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check"; timeout 600 $TR --master-port 29901 tools/multigpu_check.py --sections correct,stress --stress-iters 800 --out $OUT/check.json > $OUT/check.log 2>&1; echo "rc=$?"; grep -E "^\[correct|^\[stress|FAIL|ERROR" $OUT/check.log | head -12 | cut -c1-400
echo "== sweep"; timeout 600 $TR --master-port 29902 tools/multigpu_check.py --sections sweep --max-bytes 268435456 --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?"; python - <<PY
import json
d=json.load(open("$OUT/sweep.json"))
for r in d["sweep"]:
    print(r["fp32_bytes"], {k:v for k,v in r.items() if k.endswith("_us")}, r.get("fused_twoshot_phases_us"))
PY
echo "== pytest multigpu"; timeout 900 python -m pytest tests/test_multigpu.py -x -q > $OUT/pytest_mg.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_mg.log | cut -c1-300
echo "== bench ours N=$N"; timeout 400 $TR --master-port 29903 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; cut -c1-330 $OUT/bench_ours.json
echo "== trainers (1 GPU)"
(cd /tmp/work && timeout 300 python $OLDPWD/examples/ResNet18/mix.py --grad_exp 4 --grad_man 3 --use_APS --use_lars --emulate_node 2 --max-iter 60 > mix.log 2>&1; echo "mix rc=$?"; timeout 300 python $OLDPWD/examples/ResNet18/mix.py --grad_exp 4 --grad_man 3 --use_APS --use_lars --emulate_node 2 --max-iter 110 --load-path checkpoints/res18_cifar/ckpt_50.pth > mix_resume.log 2>&1; echo "resume rc=$?"; grep -E "Iter|resumed|Error" mix_resume.log | tail -3 | cut -c1-200; timeout 400 python $OLDPWD/examples/ResNet50/main.py --grad_exp 5 --grad_man 2 --use-APS --dataset-len 512 --max-iter 12 --log-dir logs > r50.log 2>&1; echo "r50 rc=$?"; grep -E "Iter|All Loss|Error" r50.log | tail -3 | cut -c1-200)
echo "== dist trainer ($N GPUs)"; (cd /tmp/work && timeout 300 $TR --master-port 29904 $OLDPWD/examples/ResNet18/mix.py --dist --grad_exp 4 --grad_man 3 --use_APS --max-iter 60 > mix_dist.log 2>&1; echo "rc=$?"; grep -E "Iter|All Loss|Error" mix_dist.log | tail -3 | cut -c1-200)
echo "== facade vs plain on GPU"; timeout 200 python - <<'PY' 2>&1 | tail -4
import torch
from CPDtorch.utils.dist_util import sum_gradients
def run(mode):
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(16 * 16 * 16, 10)).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=0.01)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(64, 3, 16, 16, generator=g).cuda(), torch.randint(0, 10, (64,), generator=g).cuda()
    out = []
    for i in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        if mode == 'fp32': sum_gradients(m, use_APS=False, grad_exp=8, grad_man=23)
        elif mode == 'e4m3': sum_gradients(m, use_APS=True, grad_exp=4, grad_man=3)
        opt.step()
        out.append(round(loss.item(), 4))
    return out
for mode in ['none', 'fp32', 'e4m3']: print(mode, run(mode))
PY
