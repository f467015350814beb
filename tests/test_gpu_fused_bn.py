"""BatchNormAct2d (csrc/cuda/bn_act_sm100.cu) against the stock ops it replaces: nn.BatchNorm2d -> (+ residual) -> ReLU,
forward, backward (dx, dresidual, dgamma, dbeta) and running statistics, on channels_last fp32 CUDA tensors."""
import pytest
import torch
import torch.nn.functional as F

from cpd_b200.nn import BatchNormAct2d

pytestmark = pytest.mark.gpu


def _mostly_close(a, b, tol, max_bad=4):
    """Elementwise agreement, except for the handful of elements whose pre-activation is within round-off of zero: there
    the two implementations may take different sides of the ReLU and the masked gradient differs by a whole dy."""
    bad = (a - b).abs() > tol
    return int(bad.sum()) <= max_bad + int(2e-6 * a.numel())


@pytest.mark.parametrize("shape", [(8, 64, 16, 16), (4, 128, 8, 8), (3, 512, 4, 4), (2, 2048, 7, 7), (32, 64, 32, 32), (5, 12, 3, 5)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_fused_bn_act_matches_stock_ops(shape, relu, res):
    torch.manual_seed(shape[1] + int(relu) * 2 + int(res))
    B, C, H, W = shape
    fused = BatchNormAct2d(C, relu=relu).cuda()
    stock = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        fused.weight.uniform_(0.5, 1.5); fused.bias.normal_()
        fused.running_mean.normal_(); fused.running_var.uniform_(0.5, 2.0)
    stock.load_state_dict(fused.state_dict())
    cl = torch.channels_last
    # a mean far from zero relative to the spread: the shifted-data sums must stay accurate
    x = (torch.randn(shape, device="cuda") * 0.5 + torch.linspace(-3, 3, C, device="cuda").view(1, C, 1, 1)).contiguous(memory_format=cl)
    r = torch.randn(shape, device="cuda").contiguous(memory_format=cl) if res else None
    xs, rs = x.clone().requires_grad_(True), (r.clone().requires_grad_(True) if res else None)
    xf, rf = x.clone().requires_grad_(True), (r.clone().requires_grad_(True) if res else None)
    for step in range(2):                                      # second step: running stats as the shift, re-armed counters
        y = fused(xf, rf)
        ys = stock(xs)
        if res:
            ys = ys + rs
        if relu:
            ys = F.relu(ys)
        assert y.is_contiguous(memory_format=cl)
        assert torch.allclose(y, ys, rtol=1e-4, atol=1e-4), float((y - ys).abs().max())
        dy = torch.randn_like(ys)
        for t in (xf, xs, fused.weight, fused.bias, stock.weight, stock.bias) + ((rf, rs) if res else ()):
            t.grad = None
        y.backward(dy)
        ys.backward(dy)
        scale = float(xs.grad.abs().max()) + 1e-6
        assert _mostly_close(xf.grad, xs.grad, 2e-4 * scale + 1e-5)
        if res:
            assert _mostly_close(rf.grad, rs.grad, 0.0)
        # (a boundary flip moves one channel's sums by one dy: allow a few channels off)
        assert _mostly_close(fused.weight.grad, stock.weight.grad, 1e-3 * float(stock.weight.grad.abs().max()) + 1e-5)
        assert _mostly_close(fused.bias.grad, stock.bias.grad, 1e-3 * float(stock.bias.grad.abs().max()) + 1e-5)
        assert torch.allclose(fused.running_mean, stock.running_mean, rtol=1e-5, atol=1e-5)
        assert torch.allclose(fused.running_var, stock.running_var, rtol=1e-4, atol=1e-5)
        assert int(fused.num_batches_tracked) == int(stock.num_batches_tracked) == step + 1


def test_fused_bn_resnet_equals_stock_resnet_and_captures_in_a_graph():
    from cpd_b200 import PrecisionConfig
    from cpd_b200.models import create_model
    from cpd_b200.train import TrainEngine
    xb = torch.randn(32, 3, 32, 32).pin_memory()
    yb = torch.randint(0, 10, (32,)).pin_memory()
    torch.backends.cudnn.deterministic = True
    runs = []
    for fused_bn in (False, True):
        torch.manual_seed(0)
        eng = TrainEngine(create_model("resnet18_cifar", fused_bn=fused_bn), PrecisionConfig(grad_exp=8, grad_man=23), batch_size=32,
                          lr=0.02, weight_decay=1e-4, cuda_graph=True, warmup_steps=1)
        runs.append([eng.train_step(xb, yb) for _ in range(4)])
    assert abs(runs[0][0] - runs[1][0]) <= 1e-4 * abs(runs[0][0])          # same weights, same batch: same first loss
    assert all(abs(a - b) <= 0.05 * abs(a) + 1e-3 for a, b in zip(runs[0], runs[1])), runs
    assert runs[1][-1] < runs[1][0]
