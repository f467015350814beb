"""Host-side planning logic: bucket ranges, wire selection, algorithm choice, arena views."""
import pytest
import torch

from cpd_b200 import PrecisionConfig
from cpd_b200.parallel import GradReducer, ModelArena
from cpd_b200.parallel.arena import ArenaLayout, _is_dense


def _net():
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Flatten(), torch.nn.Linear(8 * 36, 20), torch.nn.Linear(20, 4))


def test_layout_alignment_and_chunk_table():
    lay = ArenaLayout([10, 4096, 4097, 1])
    assert lay.offsets == [0, 16, 16 + 4096, 16 + 4096 + 4112, 16 + 4096 + 4112 + 16]
    tab = lay.table("cpu")
    assert tab.n_chunks == 1 + 1 + 2 + 1 and tab.tchunk_off.tolist() == [0, 1, 2, 4, 5]
    sub = lay.table("cpu", 1, 3)
    assert sub.base == 16 and sub.seg_off.tolist() == [0, 4096, 4096 + 4112] and sub.T == 2


def test_arena_views_share_storage_and_keep_memory_format():
    m = _net().to(memory_format=torch.channels_last)
    w = m[0].weight.data.clone()
    arena = ModelArena(list(m.parameters()))
    assert torch.equal(m[0].weight.data, w) and m[0].weight.stride() == w.stride()
    assert m[0].weight.data_ptr() == arena.model_flat.data_ptr()
    m(torch.randn(2, 3, 8, 8)).sum().backward()
    assert m[0].weight.grad.data_ptr() == arena.grad_flat.data_ptr()            # autograd accumulated into the arena
    assert float(arena.grad_flat.abs().sum()) > 0
    assert _is_dense(m[0].weight) and not _is_dense(torch.empty(4, 4)[:, ::2])


def test_bucket_ranges_cover_all_tensors_in_reverse_order():
    arena = ModelArena(list(_net().parameters()))
    T = arena.layout.T
    assert arena.bucket_ranges(0) == [(0, T)]
    r = arena.bucket_ranges(2000)
    assert r[0][1] == T and r[-1][0] == 0 and all(a < b for a, b in r)
    assert sorted(t for a, b in r for t in range(a, b)) == list(range(T))


def test_reducer_wire_and_mode_selection():
    arena = ModelArena(list(_net().parameters()))
    assert GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True)).wire == 8
    assert GradReducer(arena, PrecisionConfig(grad_exp=5, grad_man=10, use_aps=True)).wire == 16
    assert GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=False)).wire == 32          # raw fp32 travels
    assert GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=False, prequantize=True)).wire == 8
    assert GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, wire="16")).wire == 16
    with pytest.raises(ValueError):
        GradReducer(arena, PrecisionConfig(grad_exp=5, grad_man=10, use_aps=True, wire="8"))
    r = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, bucket_mb=0.001))
    assert r.mode == "local" and len(r.buckets) > 1
    for p in arena.params:
        p.grad.normal_()
    before = arena.grad_flat.clone()
    r.reduce()                                             # multi-bucket local path == single bucket
    one = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True))
    arena.grad_flat.copy_(before)
    one.reduce()
    after_one = arena.grad_flat.clone()
    arena.grad_flat.copy_(before)
    r.reduce()
    assert torch.equal(arena.grad_flat, after_one)


def test_variant_selection_follows_the_measured_thresholds():
    """algo_for(): raw (one flag round) for small APS buckets, packed one-shot up to the one-shot limit of wire codes,
    NVLS / two-shot above; plain fp32 -> in-switch sum; 2 ranks always one-shot; a forced algo is honoured.  The limits
    come from the topology (link bandwidth x flag-round latency) or from a calibration run."""
    from types import SimpleNamespace
    from cpd_b200.ops import native
    from cpd_b200.parallel import reducer as R
    th8 = R.Thresholds(8, link_gbs=900.0)
    assert 64_000 < th8.raw_max_elems < 200_000 and 512 << 10 <= th8.oneshot_max_bytes <= 2 << 20
    assert R.Thresholds(2, link_gbs=900.0).raw_max_elems > th8.raw_max_elems        # fewer peers: raw pays longer
    assert R.Thresholds(8, link_gbs=450.0).raw_max_elems < th8.raw_max_elems        # slower links: hand over earlier
    arena = ModelArena(list(_net().parameters()))
    r = GradReducer(arena, PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True))
    bucket = lambda n: SimpleNamespace(numel=n)                                     # noqa: E731
    r.world, r.symm = 8, SimpleNamespace(has_multicast=True)
    r.thresholds = th = R.Thresholds(8, raw_max_elems=4096, oneshot_max_bytes=1 << 19, source="test")
    assert r.algo_for(bucket(th.raw_max_elems)) == native.ALGO_ONESHOT_RAW
    assert r.algo_for(bucket(th.raw_max_elems + 16)) == native.ALGO_ONESHOT
    assert r.algo_for(bucket(th.oneshot_max_bytes)) == native.ALGO_ONESHOT            # 8-bit wire: bytes == elements
    assert r.algo_for(bucket(th.oneshot_max_bytes + 16)) == native.ALGO_NVLS
    r.plain_fp32 = True
    assert r.algo_for(bucket(1 << 20)) == native.ALGO_NVLS_SUM and r.algo_for(bucket(1024)) == native.ALGO_ONESHOT
    r.plain_fp32 = False
    r.symm = SimpleNamespace(has_multicast=False)
    assert r.algo_for(bucket(1 << 24)) == native.ALGO_TWOSHOT
    r.world = 2
    assert r.algo_for(bucket(1 << 24)) == native.ALGO_ONESHOT
    r.world = 8
    r.cfg = PrecisionConfig(grad_exp=4, grad_man=3, use_aps=False, prequantize=True)
    assert r.algo_for(bucket(64)) == native.ALGO_ONESHOT                            # raw variant needs APS exponents
    r.cfg = PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True, algo="twoshot")
    assert r.algo_for(bucket(64)) == native.ALGO_TWOSHOT


def test_steal_grads_mode_equals_accumulate_into_arena_mode():
    """TrainEngine(steal_grads=True): backward hands over fresh gradient tensors (p.grad None beforehand) and one
    multi-tensor copy moves them into the arena - same losses and weights as accumulating into the arena views,
    including a parameter that receives no gradient."""
    from cpd_b200.train import TrainEngine

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(48, 16)
            self.unused = torch.nn.Linear(4, 4)                 # never in the graph: its gradient must read as zero
            self.b = torch.nn.Linear(16, 10)

        def forward(self, x):
            return self.b(torch.relu(self.a(x.flatten(1))))

    def run(steal):
        torch.manual_seed(0)
        eng = TrainEngine(Net(), PrecisionConfig(grad_exp=4, grad_man=3, use_aps=True), batch_size=4, input_shape=(3, 4, 4), lr=0.05,
                          weight_decay=1e-3, cuda_graph=False, steal_grads=steal)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(4, 3, 4, 4, generator=g), torch.randint(0, 10, (4,), generator=g)
        losses = [eng.train_step(x, y) for _ in range(4)]
        assert all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(eng.arena.params, eng.arena.grad_views))
        return losses, eng.arena.model_flat.clone()

    l0, w0 = run(False)
    l1, w1 = run(True)
    assert l0 == l1 and torch.equal(w0, w1)
