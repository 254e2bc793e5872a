"""CPU: host-side pieces either side of the hot path: checkpoint interchange with the
reference's file format, shape scheduling (progressive resize, cluster batches), validation metrics."""
import collections
import os
import random

import pytest
import torch

from mnb200 import checkpoint as ck
from mnb200 import evaluate as ev
from mnb200 import schedule as sch
from mnb200.engine import ParamStore
from models.classifiers import FineTuneModelPool, load_model
from oracle import mnasnet_oracle as O


def _model(seed=0):
    torch.manual_seed(seed)
    return FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')


def test_state_dict_interchange_has_reference_keys_and_prefix():
    m = _model()
    sd = ck.reference_state_dict(m)
    assert len(sd) == 403 and all(k.startswith("module.") for k in sd)
    torch.manual_seed(42)
    ref_keys = list(O.init_state_dict().keys())
    assert [k[len("module."):] for k in sd] == ref_keys
    # round trip into a differently initialised model, with and without the DataParallel prefix
    for prefix in (True, False):
        m2 = _model(seed=1)
        ck.load_reference_state_dict(m2, ck.reference_state_dict(m, parallel_prefix=prefix))
        for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), k1
    # a checkpoint whose aliases of one shared block disagree is rejected
    bad = ck.reference_state_dict(m)
    alias = [g for g in ck._alias_groups(m)][0]
    bad["module." + alias[1]] = bad["module." + alias[1]] + 1
    with pytest.raises(ValueError):
        ck.load_reference_state_dict(_model(2), bad)


def test_oracle_checkpoint_loads_into_dropin_and_back():
    """A state_dict produced by the reference-faithful initialiser loads into the drop-in module unchanged."""
    torch.manual_seed(42)
    sd = O.init_state_dict()
    m = _model(3)
    ck.load_reference_state_dict(m, collections.OrderedDict(("module." + k, v.detach().clone()) for k, v in sd.items()))
    out = ck.reference_state_dict(m, parallel_prefix=False)
    for k, v in sd.items():
        assert torch.equal(out[k], v.detach()), k


def test_adam_state_matches_torch_optim_layout_and_trajectory(tmp_path):
    m = _model()
    store = ParamStore(m, "cpu")                       # parameters become views into one flat buffer
    params = [p for p in m.parameters() if p.requires_grad]
    assert len(params) == 112
    opt = torch.optim.Adam(params, lr=1e-3)             # src/train.py:219-221
    g = torch.Generator().manual_seed(5)
    for _ in range(2):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        opt.step()
    osd = opt.state_dict()
    step = ck.load_adam_state_dict(m, store, osd)
    assert step == 2
    back = ck.adam_state_dict(m, store, step)
    assert back["param_groups"][0]["params"] == osd["param_groups"][0]["params"]
    assert set(back["state"].keys()) == set(osd["state"].keys())
    for i, st in osd["state"].items():
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == float(st["step"]) == 2.0
    # torch accepts the exported state, and one more step of flat Adam (the checker's single-tensor Adam over the
    # flat buffers, i.e. what mnb_adam_step computes) lands where torch.optim.Adam lands
    opt2 = torch.optim.Adam(params, lr=1e-3)
    opt2.load_state_dict(back)
    for p in params:
        p.grad = torch.randn(p.shape, generator=g) * 1e-2
    for p in params:
        o, n = store.offsets[id(p)]
        store.grad[o:o + n].view(p.shape).copy_(p.grad)
    flat = store.flat.clone()
    mm, vv = store.adam_state()
    mm, vv = mm.clone(), vv.clone()
    O.adam_step(flat, store.grad, mm, vv, step + 1)
    opt2.step()
    for p in params:
        o, n = store.offsets[id(p)]
        torch.testing.assert_close(flat[o:o + n].view(p.shape), p.data, rtol=1e-6, atol=1e-8)
    # file round trip in the reference's checkpoint format (src/train.py:380-389, 652-655)
    state = {"epoch": 7, "optimizer": back, "state_dict": ck.reference_state_dict(m), "best_loss": 0.25}
    f, best = str(tmp_path / "w" / "0_checkpoint.pth.tar"), str(tmp_path / "w" / "0_best.pth.tar")
    ck.save_checkpoint(state, True, f, best)
    assert os.path.exists(f) and os.path.exists(best)
    m3 = _model(9)
    epoch, best_loss = ck.resume(best, m3)
    assert (epoch, best_loss) == (7, 0.25)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m3.state_dict().items()):
        assert torch.equal(v1, v2), k1
    # mismatched optimizer state is rejected
    with pytest.raises(ValueError):
        ck.load_adam_state_dict(m, store, {"state": {}, "param_groups": [{"params": [0, 1]}]})


def test_progressive_resize_follows_the_reference_rule():
    # epochs_grow_size=2, start at quarter size, batch 512: grow at epochs 1 and 3, then stay (size_ratio == 1)
    s = sch.progressive_resize(epochs=7, epochs_grow_size=2, size_ratio=0.25, batch_size=512)
    assert s == [(0, 0.25, 512), (1, 0.5, 128), (2, 0.5, 128), (3, 1.0, 32), (4, 1.0, 32), (5, 1.0, 32), (6, 1.0, 32)]
    assert sch.progressive_resize(3, 0, 0.5, 64) == [(0, 0.5, 64), (1, 0.5, 64), (2, 0.5, 64)]
    assert sch.cluster_shape(0, 0.5) == (192, 256) and sch.cluster_shape(2, 0.5) == (256, 192)
    shapes = sch.distinct_shapes(s, clusters_present=(0, 1, 2), n_ranks=8)
    assert shapes[0] == (64, 96, 128) and (4, 512, 384) in shapes and len(shapes) == 9


def test_cluster_batches_restate_the_reference_sampler():
    import importlib.util
    clusters = [list(range(0, 23)), list(range(100, 110)), list(range(200, 237))]
    b = sch.cluster_batches(clusters, 8, shuffle=False)
    assert [c for c, _ in b] == [0, 0, 1, 2, 2, 2, 2]                      # short chunks dropped
    assert b[0][1] == list(range(0, 8)) and b[2][1] == list(range(100, 108))
    bs = sch.cluster_batches(clusters, 8, shuffle=True, rng=random.Random(3))
    assert sorted(map(tuple, (i for _, i in bs))) == sorted(map(tuple, (i for _, i in b)))
    for c, idx in bs:
        assert len(idx) == 8 and all(i in clusters[c] for i in idx)        # one cluster per batch
    over = sch.cluster_batches([[1, 2, 3]], 2, shuffle=False, oversampling=[[2, 1, 3]])
    assert [i for _, i in over] == [[1, 1], [2, 3], [3, 3]]
    assert sch.step_shapes(b[:3], 0.5, n_ranks=2) == [(4, 192, 256), (4, 192, 256), (4, 256, 256)]
    # against the reference class itself when it is importable (this container only)
    path = "/root/reference/src/utils/cluster_random_sampler.py"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("ref_sampler", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ds = type("D", (), {"cluster_indices": clusters})()
        ref = mod.ClusterRandomSampler(ds, 8, shuffle=False)
        assert list(iter(ref)) == [i for _, idx in b for i in idx] and len(ref) == 56


def test_topk_and_meters():
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(64, 1000, generator=g)
    target = torch.randint(0, 1000, (64,), generator=g)
    target[:10] = logits[:10].argmax(1)                                    # 10 sure top-1 hits
    a1, a5 = ev.accuracy(logits, target, (1, 5))
    rank = (logits > logits.gather(1, target[:, None])).sum(1)
    assert abs(a1.item() - 100.0 * (rank < 1).float().mean().item()) < 1e-4
    assert abs(a5.item() - 100.0 * (rank < 5).float().mean().item()) < 1e-4
    assert a1.item() >= 100.0 * 10 / 64 - 1e-4
    mt = ev.DeviceMeter("cpu")
    mt.update(torch.tensor(2.0), 3)
    mt.update(torch.tensor(4.0), 1)
    assert abs(mt.avg - 2.5) < 1e-12


def test_kernel_selection_options():
    from mnb200 import _lib
    assert (_lib.get_option("pw_stream"), _lib.get_option("stem_mma"), _lib.get_option("dw_small"), _lib.get_option("pw_wide")) == (1, 1, 1, 0)
    _lib.set_option("dw_small", 0)
    assert _lib.get_option("dw_small") == 0
    _lib.set_option("dw_small", 2)
    assert _lib.get_option("dw_small") == 2
    _lib.set_option("dw_small", 1)
    assert _lib.get_option("dw_mma_tws") == 0 and _lib.get_option("c3_mma") == 1
    with pytest.raises(_lib.MnbError):
        _lib.set_option("no_such_option", 1)


def test_lr_regimes_match_torch_and_the_reference_cyclic_lr():
    import importlib.util
    from mnb200 import lr as L
    p = [torch.nn.Parameter(torch.zeros(1))]
    # ExponentialLR(gamma=0.99), one step per epoch (src/train.py:284-286, 335)
    opt = torch.optim.Adam(p, lr=1e-3)
    sch_t = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99)
    s = L.Schedule("auto_decay", 1e-3)
    for _ in range(30):
        opt.step()
        sch_t.step()
        assert abs(s.epoch_step() - opt.param_groups[0]["lr"]) < 1e-15
    # ReduceLROnPlateau(mode='min', factor=0.5, patience=5) (src/train.py:289-293, 337)
    opt = torch.optim.Adam(p, lr=1e-3)
    sch_t = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=5)
    s = L.Schedule("plateau_decay", 1e-3)
    g = torch.Generator().manual_seed(0)
    losses = [1.0 / (1 + 0.3 * i) if i < 8 else 0.3 + 0.01 * torch.rand((), generator=g).item() for i in range(60)]
    for v in losses:
        sch_t.step(v)
        assert abs(s.epoch_step(v) - opt.param_groups[0]["lr"]) < 1e-15
    assert opt.param_groups[0]["lr"] < 1e-3                      # it did decay
    # CyclicLR exp_range as configured at src/train.py:296-301, against the reference's own class when present
    path = "/root/reference/src/utils/cyclic_lr.py"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("ref_clr", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        opt = torch.optim.Adam(p, lr=1e-3)
        ref = mod.CyclicLR(optimizer=opt, base_lr=1e-4, max_lr=1e-2, step_size=1200, mode="exp_range", gamma=0.95)
        s = L.Schedule("clr", 1e-3)
        assert abs(s.lr - opt.param_groups[0]["lr"]) < 1e-15
        for _ in range(3000):
            ref.batch_step()
            assert abs(s.batch_step() - opt.param_groups[0]["lr"]) < 1e-12
    assert abs(L.cyclic(600, mode="triangular") - (1e-4 + (1e-2 - 1e-4) * 0.5)) < 1e-15
    assert L.epoch_decay(1e-3, 48) == 1e-3 and abs(L.epoch_decay(1e-3, 49) - 9e-4) < 1e-18
    with pytest.raises(ValueError):
        L.Schedule("nope", 1e-3)
