"""Host-side distributed logic on CPU with the Gloo backend, world_size 2 and 3
(SURVEY §4 'Distributed, CPU multi-process'): launcher detection, DistModule broadcast, the
gather-mode reducer == single-process oracle bitwise, every rank bit-identical."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from cpd_b200.parallel import dist_init
    r, w = dist_init(backend="gloo")
    assert (r, w) == (rank, world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world):
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.ReLU(), torch.nn.Linear(33, 7), torch.nn.BatchNorm1d(7))


def _grads_for(rank, model):
    g = torch.Generator().manual_seed(100 + rank)
    return [torch.randn(p.shape, generator=g) * (10.0 ** -(i + 1)) for i, p in enumerate(model.parameters())]


def _case_sum_gradients(rank, world):
    from cpd_b200 import sum_gradients, DistModule
    out = {}
    for name, kw in {"aps": dict(use_APS=True, grad_exp=4, grad_man=3),
                     "aps_kahan": dict(use_APS=True, grad_exp=5, grad_man=2, use_kahan=True),
                     "plain": dict(use_APS=False, grad_exp=5, grad_man=2),
                     "fp32": dict(use_APS=False, grad_exp=8, grad_man=23)}.items():
        model = DistModule(_model(rank))              # different init per rank -> broadcast must fix it
        for p, g in zip(model.parameters(), _grads_for(rank, model)):
            p.grad = g.clone()
        sum_gradients(model, **kw)
        out[name] = [p.grad.clone() for p in model.parameters()]
        out[name + "_w0"] = next(model.parameters()).detach().clone()
    return out


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sum_gradients_gloo_matches_oracle(world):
    from oracle import aps_fold_oracle
    res = _run("_case_sum_gradients", world)
    model = _model(0)
    per_rank = [_grads_for(r, model) for r in range(world)]
    for name, (aps, kahan, e, m) in {"aps": (True, False, 4, 3), "aps_kahan": (True, True, 5, 2), "plain": (False, False, 5, 2),
                                      "fp32": (False, False, 8, 23)}.items():
        for t in range(len(per_rank[0])):
            want, _ = aps_fold_oracle([per_rank[r][t].reshape(-1).numpy() for r in range(world)], e, m, aps, kahan, float(world))
            for r in range(world):
                got = res[r][name][t].reshape(-1).numpy()
                if name == "fp32":      # plain collective all_reduce (as the reference, dist_util.py:55-59): order is the backend's
                    assert np.allclose(got, want, rtol=1e-5, atol=1e-7), (name, t, r)
                    assert np.array_equal(got, res[0][name][t].reshape(-1).numpy())
                else:
                    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, t, r)
        # DistModule broadcast: every rank ended up with rank 0's weights
        for r in range(1, world):
            assert torch.equal(res[r][name + "_w0"], res[0][name + "_w0"])


def _case_engine(rank, world):
    from cpd_b200 import PrecisionConfig
    from cpd_b200.train import TrainEngine
    torch.manual_seed(1234 + rank)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Flatten(), torch.nn.Linear(8 * 64, 10))
    eng = TrainEngine(model, PrecisionConfig(grad_exp=5, grad_man=2, use_aps=True), batch_size=8, input_shape=(3, 8, 8), lr=0.1,
                      weight_decay=1e-4, use_lars=True, lars_eps=1e-9, cuda_graph=False)
    g = torch.Generator().manual_seed(rank)
    x, y = torch.randn(8, 3, 8, 8, generator=g), torch.randint(0, 10, (8,), generator=g)
    losses = [eng.train_step(x, y) for _ in range(5)]
    return {"losses": losses, "w": eng.arena.model_flat.clone(), "mode": eng.reducer.mode}


def test_engine_two_ranks_stay_in_sync():
    res = _run("_case_engine", 2)
    assert res[0]["mode"] == "gather"
    assert torch.equal(res[0]["w"], res[1]["w"])            # identical replicas after 5 steps
    assert all(np.isfinite(res[0]["losses"]))


def test_detect_launcher_variants():
    from cpd_b200.parallel import detect_launcher
    from cpd_b200.parallel.dist import _first_slurm_host
    assert detect_launcher({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3", "MASTER_ADDR": "10.0.0.1", "MASTER_PORT": "2222"}) == dict(
        launcher="torchrun", rank=3, world_size=8, local_rank=3, master_addr="10.0.0.1", master_port=2222)
    s = detect_launcher({"SLURM_PROCID": "5", "SLURM_NTASKS": "16", "SLURM_LOCALID": "5", "SLURM_NODELIST": "gpu[03-05,07]"})
    assert (s["launcher"], s["rank"], s["world_size"], s["master_addr"], s["master_port"]) == ("slurm", 5, 16, "gpu03", 12345)
    o = detect_launcher({"OMPI_COMM_WORLD_RANK": "1", "OMPI_COMM_WORLD_SIZE": "2", "OMPI_COMM_WORLD_LOCAL_RANK": "1"})
    assert (o["launcher"], o["rank"], o["world_size"], o["local_rank"]) == ("openmpi", 1, 2, 1)
    assert detect_launcher({})["launcher"] == "single"
    assert _first_slurm_host("nodeA,nodeB") == "nodeA" and _first_slurm_host("n[1,3]") == "n1" and _first_slurm_host("solo") == "solo"


def test_single_process_api_without_process_group():
    """world_size == 1, no init_process_group: sum_gradients still applies the APS round trip."""
    from cpd_b200 import sum_gradients
    from oracle import aps_fold_oracle
    model = _model(0)
    gs = _grads_for(0, model)
    for p, g in zip(model.parameters(), gs):
        p.grad = g.clone()
    sum_gradients(model, use_APS=True, grad_exp=4, grad_man=3)
    for p, g in zip(model.parameters(), gs):
        want, _ = aps_fold_oracle([g.reshape(-1).numpy()], 4, 3, True, False, 1.0)
        assert np.array_equal(p.grad.reshape(-1).numpy().view(np.uint32), want.view(np.uint32))
    # gradients replaced by fresh tensors (zero_grad(set_to_none=True) + backward) are re-adopted
    for p, g in zip(model.parameters(), gs):
        p.grad = None
    x = torch.randn(4, 20)
    model(x).sum().backward()
    before = [p.grad.clone() for p in model.parameters()]
    sum_gradients(model, use_APS=False, grad_exp=8, grad_man=23)
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.grad, b)
