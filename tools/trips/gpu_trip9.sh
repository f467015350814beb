#!/bin/bash
# 1-GPU: run the three example trainers for a few iterations (synthetic data), checkpoint/resume, NVTX flag.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/trip9
mkdir -p $OUT
echo "== ResNet18 mix.py (emulate_node 2, LARS)"; (cd $OUT && timeout 300 python ../../examples/ResNet18/mix.py --grad_exp 4 --grad_man 3 --use_APS --use_lars --emulate_node 2 --max-iter 60 > mix.log 2>&1; echo "rc=$?"; grep -E "Iter|All Loss|resumed|Error" mix.log | tail -6 | cut -c1-200)
echo "== ResNet18 mix.py resume"; (cd $OUT && timeout 300 python ../../examples/ResNet18/mix.py --grad_exp 4 --grad_man 3 --use_APS --use_lars --emulate_node 2 --max-iter 110 --load-path checkpoints/res18_cifar/ckpt_50.pth > mix_resume.log 2>&1; echo "rc=$?"; grep -E "Iter|All Loss|resumed|Error" mix_resume.log | tail -5 | cut -c1-200)
echo "== DavidNet dawn.py (1 epoch of 4096 synthetic images)"; (cd $OUT && timeout 300 python ../../examples/DavidNet/dawn.py --dist 0 --epoch 2 --dataset_len 8192 --grad_exp 5 --grad_man 2 --use_APS --use_kahan > dawn.log 2>&1; echo "rc=$?"; tail -6 dawn.log | cut -c1-200)
echo "== ResNet50 main.py"; (cd $OUT && CPD_B200_NVTX=1 timeout 400 python ../../examples/ResNet50/main.py --grad_exp 5 --grad_man 2 --use-APS --dataset-len 512 --max-iter 12 --log-dir logs > r50.log 2>&1; echo "rc=$?"; grep -E "Iter|All Loss|Error" r50.log | tail -4 | cut -c1-200)
echo "== DDP-style API on a stock model: sum_gradients facade"; timeout 200 python - <<'PY' 2>&1 | tail -5
import torch, time
from CPDtorch.utils.dist_util import sum_gradients
from CPDtorch.quant import float_quantize
m = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(64 * 32 * 32, 10)).cuda()
opt = torch.optim.SGD(m.parameters(), lr=0.01)
x, y = torch.randn(64, 3, 32, 32, device="cuda"), torch.randint(0, 10, (64,), device="cuda")
for i in range(5):
    opt.zero_grad()
    loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()
    sum_gradients(m, use_APS=True, grad_exp=4, grad_man=3)
    opt.step()
    print("step", i, "loss", round(loss.item(), 4))
PY
