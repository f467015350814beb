"""Small public helpers that the other test files do not touch: loaders, timers, writers, topology record,
process-group helpers, the model zoo's other depths."""
import json
import os

import pytest
import torch
import torch.distributed as dist

from cpd_b200.data import cifar_shape
from cpd_b200.data.synthetic import Batches
from cpd_b200.models import create_model
from cpd_b200.parallel.dist import Topology, discover_topology, shutdown, simple_group_split
from cpd_b200.utils import ScalarWriter, Timer, kernel_table, trace


def test_batches_wrapper_yields_device_dicts():
    ds = cifar_shape(40)
    b = Batches(ds, batch_size=16, shuffle=False, drop_last=False, device=torch.device("cpu"))
    items = list(b)
    assert len(b) == 3 and [it["input"].shape[0] for it in items] == [16, 16, 8]
    assert items[0]["target"].dtype == torch.long and items[0]["input"].shape[1:] == (3, 32, 32)


def test_timer_and_scalar_writer(tmp_path):
    t = Timer(sync=False)
    a = t()
    b = t(include_in_total=False)
    assert a >= 0 and b >= 0 and t.total_time == pytest.approx(a)
    w = ScalarWriter(str(tmp_path / "tb"), rank=0)
    w.add_scalar("loss_train", 1.5, 3)
    w.close()
    assert os.listdir(str(tmp_path / "tb"))                     # event file or the JSONL fallback
    off = ScalarWriter(str(tmp_path / "tb1"), rank=1)           # non-zero ranks write nothing
    off.add_scalar("x", 1.0, 1)
    off.close()
    assert not os.path.exists(str(tmp_path / "tb1"))


def test_topology_record_and_summary():
    topo = Topology(device_count=2, devices=[{"name": "B200"}, {"name": "B200"}],
                    p2p=[[{"access": True}, {"access": True}], [{"access": True}, {"access": True}]],
                    nvlink=[{"active_links": 18, "version": 7}, {"active_links": 18, "version": 7}], multicast=True)
    # NVML's link-generation enum: 7 = NVLink 5 (what a B200 reports), 50 GB/s per link per direction; 6 = NVLink 4, 25 GB/s
    assert topo.all_p2p and topo.nvlink_links == 18 and topo.est_link_gbs == 900.0
    assert Topology(device_count=1, nvlink=[{"active_links": 18, "version": 6}]).est_link_gbs == 450.0
    assert "2x B200" in topo.summary() and "multicast=True" in topo.summary()
    topo.p2p[0][1]["access"] = False
    assert not topo.all_p2p and "partial" in topo.summary()
    here = discover_topology()                                   # GPU-less container: an empty record, not an exception
    assert here.device_count >= 0 and isinstance(here.summary(), str)


def test_group_split_and_shutdown(tmp_path):
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        g = simple_group_split(1, 0, 1)
        t = torch.ones(3)
        dist.all_reduce(t, group=g)
        assert torch.equal(t, torch.ones(3))
    finally:
        shutdown()
    assert not dist.is_initialized()
    shutdown()                                                   # idempotent


@pytest.mark.parametrize("name,tensors,params", [("resnet34_cifar", 110, 21282122), ("resnet18", 62, 11689512),
                                                  ("resnet34", 110, 21797672), ("resnet101", 314, 44549160)])
def test_model_zoo_depths(name, tensors, params):
    m = create_model(name)
    ps = list(m.parameters())
    assert len(ps) == tensors and sum(p.numel() for p in ps) == params


def test_kernel_table_of_a_trace(tmp_path):
    with trace(str(tmp_path / "t.json")) as prof:
        torch.randn(64, 64) @ torch.randn(64, 64)
    assert "Self CPU" in kernel_table(prof, row_limit=5) and json.load(open(str(tmp_path / "t.json")))["traceEvents"]
