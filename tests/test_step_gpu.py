"""GPU: the fused training step as bench.py drives it (CUDA-graph replay), its state handling (capture rollback,
step counters, checkpoints taken from the graph path), frozen parameters (FineTuneModelPool.freeze, classifiers.py:94-105
+ train.py:219) and the dropout random stream.  Reference: src/train.py:433-440 (step), :219-221 (optimizer over
requires_grad parameters), :233-250 (resume), src/models/classifiers.py:80-105."""
import os

import numpy as np
import pytest
import torch

from oracle import mnasnet_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(dtype, seed=42, dropout_eval=True):
    from test_net_gpu import build
    m = build(dtype, seed=seed)
    if not dropout_eval:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.train()
    return m


def _state(eng):
    st = eng.store
    return [st.flat.clone(), st.fbuf.clone(), st.ibuf.clone(), st.m.clone(), st.v.clone(), eng.dev_step.clone(),
            eng.dev_fwd.clone()]


def test_graph_capture_leaves_training_state_bit_unchanged():
    """capture_step_graph runs one eager warm-up step and must roll it back completely: parameters, BN running
    statistics + num_batches_tracked, Adam moments, the optimizer step counter and the dropout forward counter."""
    from mnb200 import engine, schedule
    m = _build("fp32", dropout_eval=False)
    eng = engine.engine_for(m)
    x, t = O.synthetic_batch(4, 64, 64)
    xd, td = x.cuda(), t.cuda()
    eng.train_step(xd, td, lr=1e-3)                      # a non-trivial state to preserve
    before = _state(eng)
    eng.capture_step_graph(xd, td, 1e-3)
    built = schedule.warm_plans(eng, [(2, 96, 64), (4, 64, 64)], graphs=True)
    torch.cuda.synchronize()
    assert built == [(2, 96, 64), (4, 64, 64)] and len(eng.graphs) == 2
    for a, b in zip(before, _state(eng)):
        assert torch.equal(a, b)
    assert eng.host_step == 1


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_train_step_graph_matches_eager(dtype):
    """3 graph replays == 3 eager fused steps (same kernels, same order), and in fp32 both follow the golden Adam
    trajectory minted from the live reference.  The step counter advances on the device with every replay."""
    from mnb200 import engine
    x, t = O.synthetic_batch(8, 224, 224)
    xd, td = x.cuda(), t.cuda()
    m1, m2 = _build(dtype), _build(dtype)
    e1, e2 = engine.engine_for(m1), engine.engine_for(m2)
    le = [e1.train_step(xd, td, lr=1e-3).item() for _ in range(3)]
    lg = [e2.train_step_graph(xd, td, lr=1e-3).item() for _ in range(3)]
    torch.cuda.synchronize()
    print(dtype, "eager", le, "graph", lg)
    assert e1.host_step == 3 and e2.host_step == 3
    # first step: identical inputs and weights, only the atomics' order differs (in bf16 a 1e-7 difference in a BN
    # scale can flip the rounding of a stored activation: measured 4e-5)
    assert abs(le[0] - lg[0]) <= (1e-5 if dtype == "fp32" else 1e-3) * abs(le[0])
    # later steps sit behind Adam's sign-like updates of near-zero gradients (run-to-run noise ~2e-3)
    np.testing.assert_allclose(lg, le, rtol=6e-3)
    if dtype == "fp32":
        fx = np.load(os.path.join(GOLD, "adam_traj.npz"))
        np.testing.assert_allclose(lg, fx["loss_dropout_off"], rtol=6e-3)
        assert abs(lg[0] - fx["loss_dropout_off"][0]) / lg[0] < 1e-4
    for (k, p), q in zip(m1.named_parameters(), m2.parameters()):
        if not k.endswith("conv.bias"):
            assert (p - q).abs().max().item() < 3 * 2.2e-3, k       # 3 steps x 2*lr (+10 %: bias-corrected Adam steps can exceed lr)
    for (k, a), (_, b) in zip(m1.named_buffers(), m2.named_buffers()):
        if k.endswith("num_batches_tracked"):
            assert torch.equal(a, b), k


def test_graph_replay_changes_lr_without_recapture():
    """The learning rate is a device scalar: ExponentialLR (train.py:282-286) changes it between replays."""
    from mnb200 import engine
    m = _build("fp32")
    eng = engine.engine_for(m)
    x, t = O.synthetic_batch(2, 64, 64)
    xd, td = x.cuda(), t.cuda()
    eng.train_step_graph(xd, td, lr=1e-3)
    p1 = eng.store.flat.clone()
    eng.train_step_graph(xd, td, lr=0.0)                   # same graph, lr 0: parameters must not move
    torch.cuda.synchronize()
    assert len(eng.graphs) == 1 and torch.equal(p1, eng.store.flat) and eng.host_step == 2


def test_checkpoint_from_graph_path_resumes_into_eager_path():
    """step counters stay consistent on the graph path.  2 graph replays -> checkpoint (reference format,
    train.py:380-389) -> resume into a fresh model: restored state is BIT-identical (parameters, BN buffers, Adam
    moments, step) and the third step's loss agrees with the uninterrupted run to the step's run-to-run noise."""
    from mnb200 import checkpoint, engine
    x, t = O.synthetic_batch(4, 64, 64)
    xd, td = x.cuda(), t.cuda()
    m = _build("fp32")
    eng = engine.engine_for(m)
    for _ in range(2):
        eng.train_step_graph(xd, td, lr=1e-3)
    state = checkpoint.make_checkpoint(m, eng, epoch=3, best_loss=1.5)
    assert len(state["state_dict"]) == 403 and len(state["optimizer"]["state"]) == 112
    assert all(float(s["step"]) == 2.0 for s in state["optimizer"]["state"].values())
    m2 = _build("fp32", seed=7)
    eng2 = engine.engine_for(m2)
    epoch, best = checkpoint.resume(state, m2, eng2, load_optimizer=True)
    assert (epoch, best) == (3, 1.5) and eng2.host_step == 2
    torch.cuda.synchronize()
    assert torch.equal(eng.store.flat, eng2.store.flat)
    assert torch.equal(eng.store.fbuf, eng2.store.fbuf) and torch.equal(eng.store.ibuf, eng2.store.ibuf)
    assert torch.equal(eng.store.m, eng2.store.m) and torch.equal(eng.store.v, eng2.store.v)
    l3 = eng.train_step_graph(xd, td, lr=1e-3).item()
    l3b = eng2.train_step(xd, td, lr=1e-3).item()
    assert abs(l3 - l3b) <= 1e-5 * abs(l3)
    assert eng.host_step == eng2.host_step == 3


def test_freeze_is_honoured_by_the_fused_step():
    """FineTuneModelPool.freeze() (classifiers.py:94-99) + an optimizer built over requires_grad parameters
    (train.py:219): the fused step must leave the features' parameters, gradients and Adam moments untouched and
    train the classifier exactly as the unfrozen step would on its first step; BN buffers still update."""
    from mnb200 import engine
    x, t = O.synthetic_batch(4, 64, 64)
    xd, td = x.cuda(), t.cuda()
    ref = _build("fp32")
    eref = engine.engine_for(ref)
    eref.train_step(xd, td, lr=1e-3)
    m = _build("fp32")
    m.freeze()
    eng = engine.engine_for(m)
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    rm0 = m.features[0].bn.running_mean.clone()
    n_full = eref.launches_per_step(4, 64, 64)
    n_frozen = eng.launches_per_step(4, 64, 64)
    assert n_frozen < n_full - 150                        # no dgrad / wgrad / BN-backward below the head
    l1 = eng.train_step(xd, td, lr=1e-3)
    l2 = eng.train_step_graph(xd, td, lr=1e-3)            # and through the graph path
    torch.cuda.synchronize()
    assert torch.isfinite(l1).all() and torch.isfinite(l2).all()
    st = eng.store
    for k, p in m.named_parameters():
        o, n = st.offsets[id(p)]
        if k.startswith("features"):
            assert torch.equal(p, before[k]), k
            assert st.m[o:o + n].abs().max().item() == 0.0 and st.grad[o:o + n].abs().max().item() == 0.0, k
        else:
            assert not torch.equal(p, before[k]), k
    assert not torch.equal(rm0, m.features[0].bn.running_mean)
    # first frozen step == first unfrozen step on the classifier: same gradients there (the parameters themselves are
    # compared at 2*lr: Adam's first step is a sign step, and noise-level gradients of dead hidden units flip sign)
    m3 = _build("fp32")
    m3.freeze()
    e3 = engine.engine_for(m3)
    e3.train_step(xd, td, lr=1e-3)
    torch.cuda.synchronize()
    gr, g3 = eref.store.grad_views(), e3.store.grad_views()
    for (k, p), q in zip(ref.named_parameters(), m3.parameters()):
        if k.startswith("classifier"):
            a, b = gr[id(p)], g3[id(q)]
            assert ((a - b).norm() / a.norm()).item() < 1e-4, k
            assert (p - q).abs().max().item() <= 2e-3 + 1e-7, k
    # unfreeze re-plans: the features train again
    m.unfreeze()
    eng.train_step(xd, td, lr=1e-3)
    assert not torch.equal(m.features[0].conv.weight, before["features.0.conv.weight"])
    # autograd bridge: frozen parameters get no .grad
    m3.zero_grad(set_to_none=True)
    torch.nn.CrossEntropyLoss()(m3(xd), td).backward()
    assert m3.features[0].conv.weight.grad is None and m3.classifier[4].weight.grad is not None


def test_partial_freeze_keeps_backward_above_the_first_trainable_block():
    """Only features.6 onward trainable: gradients of those blocks equal the fully-trainable run's, the rest stay 0."""
    from mnb200 import engine
    x, t = O.synthetic_batch(3, 64, 96)
    xd, td = x.cuda(), t.cuda()
    full = _build("fp32")
    torch.nn.CrossEntropyLoss()(full(xd), td).backward()
    m = _build("fp32")
    for k, p in m.named_parameters():
        if k.startswith("features") and int(k.split(".")[1]) < 6:
            p.requires_grad = False
    torch.nn.CrossEntropyLoss()(m(xd), td).backward()
    torch.cuda.synchronize()
    # same kernels on both models; the two forwards differ by the order of the BN atomics (1e-7), which flips a few
    # ReLU masks below features.7 -- at N=3 that moves single gradients by ~1e-2, a missing application
    # of a shared block or a missing skip gradient would move them by O(1)
    live_a, live_b = [], []
    for (k, p), q in zip(full.named_parameters(), m.parameters()):
        if q.requires_grad:
            if k.endswith("conv.bias"):
                continue
            d = (p.grad - q.grad).norm() / p.grad.norm().clamp_min(1e-20)
            assert d.item() < (1e-4 if k.startswith("classifier") else 8e-2), (k, d.item())
            live_a.append(p.grad.reshape(-1))
            live_b.append(q.grad.reshape(-1))
        else:
            assert q.grad is None, k
    a, b = torch.cat(live_a), torch.cat(live_b)
    assert ((a - b).norm() / a.norm()).item() < 2e-2


def test_dropout_masks_change_every_forward_and_follow_the_seed():
    """nn.Dropout draws a fresh mask per call (classifiers.py:82,85).  The device RNG is keyed by a
    per-forward counter (not the optimizer step) and seeded from torch's seed and the rank."""
    from mnb200 import engine
    torch.manual_seed(1234)
    m = _build("fp32", dropout_eval=False)
    eng = engine.engine_for(m)
    x, t = O.synthetic_batch(4, 64, 64)
    xd = x.cuda()
    plan = eng.plan(4, 64, 64)
    masks = []
    for _ in range(3):                                   # plain drop-in loop: no optimizer step in between
        m(xd)
        masks.append([mk.clone() for mk, _ in plan.dropout_masks])
    torch.cuda.synchronize()
    assert int(eng.dev_fwd.item()) == 3 and int(eng.dev_step.item()) == 0
    for a, b in ((0, 1), (1, 2), (0, 2)):
        for ma, mb in zip(masks[a], masks[b]):
            assert not torch.equal(ma, mb)
    for mk, (_, p) in zip(masks[0], plan.dropout_masks):
        keep = mk.float().mean().item()
        assert abs(keep - (1 - p)) < 0.05, (keep, p)
    # the two dropout layers of one forward are independent
    n = min(masks[0][0].shape[1], masks[0][1].shape[1])
    assert not torch.equal(masks[0][0][:, :n], masks[0][1][:, :n])
    # same torch seed -> same stream; different seed or rank -> different stream
    torch.manual_seed(1234)
    s0 = engine._default_seed()
    torch.manual_seed(1234)
    assert engine._default_seed() == s0
    torch.manual_seed(99)
    assert engine._default_seed() != s0
    os.environ["RANK"] = "3"
    try:
        torch.manual_seed(1234)
        assert engine._default_seed() != s0
    finally:
        del os.environ["RANK"]
    # eval: no dropout, deterministic
    m.eval()
    with torch.no_grad():
        assert torch.equal(m(xd), m(xd))
