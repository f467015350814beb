"""Reference arm of the benchmark: runs the UNMODIFIED drcut/CPD installed in baseline/_ref
through its own public API and stock code path (CPDtorch.utils.dist_util.{dist_init,
sum_gradients}, CPDtorch.quant.float_quantize, its own example models) — none of
cpd_b200's models, kernels or engine are imported here.

The reference's trainers cannot be launched verbatim on this box (they need SLURM, the
CIFAR/ImageNet files on disk, tensorboardX, PyYAML<6, numpy<2 — SURVEY §0.1), so this file
re-states their per-step recipe with the reference's own calls, line by line:
  ResNet-18  example/ResNet18/tools/mix.py:201-316  (--dist --grad_exp 4 --grad_man 3 --use_APS)
  ResNet-50  example/ResNet50/main.py:141-212        (--grad_exp 5 --grad_man 2 --use-APS)
  DavidNet   example/DavidNet/utils.py:328-378       (--grad_exp 5 --grad_man 2 --use_APS)
Environment-only shims (no reference file is edited):
  * SLURM_* variables synthesised from torchrun's RANK/WORLD_SIZE so dist_init() runs;
    TORCHELASTIC_USE_AGENT_STORE is dropped because dist_init overrides MASTER_PORT=12345
  * CXX=baseline/cxx_shim.sh supplies -DAT_CHECK=TORCH_CHECK for the JIT build
  * model.zero_grad(set_to_none=False): the trainers were written for torch<2 where
    zero_grad() kept the .grad tensors that sum_gradients then dereferences
  * synthetic data of the named shape from pinned host memory; CUDA-event timing
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    if not os.path.isdir(os.path.join(REF, "CPDtorch")):
        return False, "baseline/_ref not installed (run baseline/install_reference.sh)"
    return True, ""


def setup_env(rank: int, world: int, lib: str = "reference"):
    """lib="reference": CPDtorch resolves to the unmodified install in baseline/_ref (the benchmark arm).
    lib="ours": the SAME recipes (and the reference's own example models) run on this repo's import-compatible
    CPDtorch facade instead - the "user switches the library, keeps their trainer" scenario (tools/compat_run.py)."""
    if lib == "ours":
        sys.path.insert(0, os.path.dirname(HERE))                # repo root: CPDtorch/ facade + cpd_b200
    else:
        sys.path.insert(0, REF)
        from prebuild_ref_ext import ref_env
        ref_env()
    os.environ["SLURM_NODELIST"] = "hostname127-0-0-1"          # dist_util.py:109: node_list[8:].replace('-', '.')
    os.environ["SLURM_PROCID"] = str(rank)
    os.environ["SLURM_NTASKS"] = str(world)
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)


class RefRunner:
    """One model + the reference's training step."""

    def __init__(self, model_name: str, batch: int, grad_exp: int, grad_man: int, use_aps: bool, use_kahan: bool,
                 use_lars: bool, gemm_exp: int = 5, gemm_man: int = 10):
        import torch
        import torch.nn as nn
        from CPDtorch.utils.dist_util import dist_init, sum_gradients, DistModule
        from CPDtorch.quant import float_quantize  # noqa: F401  (triggers the reference's JIT build)
        self.torch = torch
        self.sum_gradients = sum_gradients
        self.rank, self.world = dist_init()
        self.model_name = model_name
        self.grad_exp, self.grad_man, self.use_aps, self.use_kahan, self.use_lars = grad_exp, grad_man, use_aps, use_kahan, use_lars
        import torch.distributed as dist
        self.dist = dist
        if model_name in ("resnet18", "resnet18_quant"):
            sys.path.insert(0, os.path.join(REF, "example", "ResNet18"))
            import models as ref_models                                        # example/ResNet18/models
            self.model = ref_models.__dict__["res_cifar"]()                      # mix.py:81 (arch: res_cifar)
            if model_name == "resnet18_quant":
                # the reference ships Quant_Conv as a library module (CPDtorch/quant/quant_module.py:88-139) and no model
                # that uses it: swap it in for every nn.Conv2d of the reference's own ResNet-18 (README.md:9 "customized-
                # precision floating point as accumulator while GEMM"); F.unfold + tvm_gemm underneath, (gemm_exp, gemm_man).
                # The 10-class head stays nn.Linear: tvm_gemm's M-remainder branch (rows beyond the last multiple of 16)
                # reads an uninitialised accumulator (float_kernel.cu:221-339).
                from CPDtorch.quant import Quant_Conv as RefQuantConv

                def swap(mod):
                    for name, child in list(mod.named_children()):
                        if isinstance(child, nn.Conv2d):
                            q = RefQuantConv(child.in_channels, child.out_channels, child.kernel_size[0], stride=child.stride[0],
                                             padding=child.padding[0], bias=child.bias is not None, exp=gemm_exp, man=gemm_man)
                            setattr(mod, name, q)
                        else:
                            swap(child)
                swap(self.model)
            self.model.cuda()
            for p in self.model.parameters():                                  # mix.py:86-88
                dist.broadcast(p, 0)
            self.model_params = [p for p in self.model.parameters() if p.requires_grad]          # mix.py:53-63
            self.master_params = [p.clone().float().detach() for p in self.model_params]
            for p in self.master_params:
                p.requires_grad = True
            self.criterion = nn.CrossEntropyLoss().cuda()
            self.momentum, self.weight_decay = 0.9, 1e-4                       # res18_cifar.yaml
            self.optimizer = torch.optim.SGD(self.master_params, 0.1, momentum=self.momentum, weight_decay=self.weight_decay)
            self.momentum_buffer = [torch.zeros_like(p) for p in self.master_params]
            self.shape, self.classes = (3, 32, 32), 10
        elif model_name == "resnet50":
            import torchvision.models as tvm
            torch.backends.cudnn.benchmark = True                               # main.py:64
            self.model = tvm.resnet50().cuda()                                  # main.py:67
            self.model = DistModule(self.model)                                 # main.py:133
            wd = 1e-4
            params = [{"params": [p for n, p in self.model.named_parameters() if "bn" not in n], "weight_decay": wd},
                      {"params": [p for n, p in self.model.named_parameters() if "bn" in n], "weight_decay": 0}]
            self.optimizer = torch.optim.SGD(params, lr=0.1, momentum=0.9, nesterov=True)        # main.py:123-131
            self.criterion = nn.CrossEntropyLoss().cuda()
            self.shape, self.classes = (3, 224, 224), 1000
        elif model_name == "davidnet":
            sys.path.insert(0, os.path.join(REF, "example", "DavidNet"))
            import numpy as np
            if not hasattr(np, "float"):
                np.float = float                                                # utils.py:55,388 use the removed alias
            import davidnet as ref_dn                                           # example/DavidNet/davidnet.py
            torch.backends.cudnn.benchmark = True                               # dawn.py:32
            self.model = ref_dn.TorchGraph(ref_dn.union(ref_dn.net(), ref_dn.losses)).cuda()
            self.model = DistModule(self.model)                                 # dawn.py:71-72
            self.optimizer = torch.optim.SGD(self.model.parameters(), lr=0.0, momentum=0.9, weight_decay=5e-4 * batch,
                                             nesterov=True)                     # dawn.py:73-79
            self.shape, self.classes = (3, 32, 32), 10
        else:
            raise ValueError(model_name)
        self.batch = batch
        g = torch.Generator().manual_seed(1234 + self.rank)
        self.pool = [(torch.randn(batch, *self.shape, generator=g).pin_memory(),
                      torch.randint(0, self.classes, (batch,), generator=g).pin_memory()) for _ in range(4)]
        self.i = 0
        self.lr = 0.1

    # ---------------------------------------------------------------- steps --
    def step(self) -> float:
        x, y = self.pool[self.i % len(self.pool)]
        self.i += 1
        return getattr(self, "_step_" + self.model_name.replace("_quant", ""))(x, y)

    def _step_resnet18(self, input, target) -> float:
        torch, dist = self.torch, self.dist
        model, world_size, emulate_node = self.model, self.world, 1
        target = target.cuda()                                                  # mix.py:233-234
        input_var = input.cuda()
        output = model(input_var, self.rank)                                    # mix.py:238
        loss = self.criterion(output, target) / (world_size * emulate_node)
        reduced_loss = loss.data.clone()
        dist.all_reduce(reduced_loss)                                           # mix.py:240-243
        loss_val = float(reduced_loss.item())
        model.zero_grad(set_to_none=False)
        loss.backward()
        grad_buffer = [[] for _ in model.parameters()]
        for idx, param in enumerate(model.parameters()):                        # mix.py:246-249
            if param.grad is not None:
                grad_buffer[idx].append(param.grad.detach().clone().data)
        model.zero_grad(set_to_none=False)
        for idx, param in enumerate(model.parameters()):                        # mix.py:253-258 (emulate_node == 1)
            if param.grad is not None:
                param.grad.data.copy_(grad_buffer[idx][0])
        self.sum_gradients(model, use_APS=self.use_aps, use_kahan=self.use_kahan, grad_exp=self.grad_exp,
                           grad_man=self.grad_man)                              # mix.py:286-291
        for model_p, master_p in zip(self.model_params, self.master_params):    # mix.py:292-294
            if model_p.grad is not None:
                master_p.backward(model_p.grad.float())
        if self.use_lars:                                                       # mix.py:297-310
            for idx, master_p in enumerate(self.master_params):
                if master_p.grad is not None:
                    local_lr = master_p.norm(2) / (master_p.grad.data.norm(2) + self.weight_decay * master_p.norm(2))
                    local_lr = local_lr * 0.001
                    self.momentum_buffer[idx] = self.momentum * self.momentum_buffer[idx].data \
                        + self.lr * local_lr * (master_p.grad.data + self.weight_decay * master_p.data)
                    master_p.data.copy_(master_p - self.momentum_buffer[idx])
        else:
            self.optimizer.step()                                               # mix.py:312
        for model_p, master_p in zip(self.model_params, self.master_params):    # mix.py:313-314
            model_p.data.copy_(master_p.data)
        self.optimizer.zero_grad(set_to_none=False)
        return loss_val

    def _step_resnet50(self, data, target) -> float:
        torch, dist = self.torch, self.dist
        import math
        import torch.nn.functional as F
        from CPDtorch.quant import float_quantize
        model, world_size, emulate_node = self.model, self.world, 1
        data, target = data.cuda(), target.cuda()                               # main.py:153-154
        self.optimizer.zero_grad(set_to_none=False)
        grad_buffer = [[] for _ in model.parameters()]
        for i in range(0, len(data), self.batch):                               # main.py:160-177
            data_batch, target_batch = data[i:i + self.batch], target[i:i + self.batch]
            output = model(data_batch)
            pred = output.max(1, keepdim=True)[1]                               # accuracy(): main.py:255-258
            pred.eq(target_batch.view_as(pred)).cpu().float().mean()
            loss = F.cross_entropy(output, target_batch) / world_size
            reduced_loss = loss.data.clone()
            dist.all_reduce(reduced_loss)
            loss_val = float(reduced_loss.item())
            loss.div_(math.ceil(float(len(data)) / self.batch))
            loss.backward()
            for idx, param in enumerate(model.parameters()):
                if param.grad is not None:
                    grad_buffer[idx].append(param.grad.detach().clone().data)
            model.zero_grad(set_to_none=False)
        for idx, param in enumerate(model.parameters()):                        # main.py:178-202 (runs even for emulate_node == 1)
            if param.grad is not None:
                max_exp = -100
                for val in grad_buffer[idx]:
                    t_exp = torch.log2(torch.abs(val * emulate_node).max()).ceil().detach().cpu().numpy()
                    if t_exp > max_exp:
                        max_exp = t_exp
                upper_bound = 2 ** (self.grad_exp - 1) - 1
                shift_factor = upper_bound - max_exp
                if max_exp == -100 or not self.use_aps:
                    shift_factor = 0
                for grad in grad_buffer[idx]:
                    grad.data.copy_(float_quantize(grad * (2 ** shift_factor), self.grad_exp, self.grad_man))
                res = torch.zeros_like(grad_buffer[idx][0])
                for val in grad_buffer[idx]:
                    res = float_quantize(res + val, self.grad_exp, self.grad_man)
                param.grad.data.copy_(res.data / (2 ** shift_factor))
        self.sum_gradients(model, use_APS=self.use_aps, grad_exp=self.grad_exp, grad_man=self.grad_man)   # main.py:203-204
        self.optimizer.step()                                                   # main.py:207
        return loss_val

    def _step_davidnet(self, x, y) -> float:
        torch, dist = self.torch, self.dist
        model = self.model
        batch = {"input": x.cuda(), "target": y.cuda()}                         # utils.py:152-178
        out = model(batch)                                                      # utils.py:368
        tmp = out["loss"].clone().detach()                                      # collect(): utils.py:347-356
        dist.all_reduce(tmp)
        loss_val = float(tmp.cpu().numpy())
        c = out["correct"].clone().detach().int()
        dist.all_reduce(c)
        c.cpu().numpy()
        self.optimizer.param_groups[0]["lr"] = 0.4 / self.batch                 # set_opt_params: utils.py:317-325, lr/batch :407-414
        model.zero_grad(set_to_none=False)                                      # update(): utils.py:328-344
        self.optimizer.zero_grad(set_to_none=False)
        model.module.cache["loss"] = model.module.cache["loss"] * 1
        model.module.cache["loss"].backward()
        self.sum_gradients(model, use_APS=self.use_aps, grad_exp=self.grad_exp, grad_man=self.grad_man)
        self.optimizer.step()
        return loss_val
