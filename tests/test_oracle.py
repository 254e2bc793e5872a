"""CPU: the oracle against the golden fixtures minted from the live reference (oracle/make_golden.py), and
the closed-form math contract of the fused kernels against autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import mnasnet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run_case(n, h, w, cfg, nc):
    torch.manual_seed(42)
    sd = O.init_state_dict(nc, cfg)
    x, t = O.synthetic_batch(n, h, w, nc)
    tr = O.Trainer(sd, classifier_config=cfg, num_classes=nc)
    logits, loss, g = tr.grads(x, t, dropout_masks="off")
    return sd, tr, logits, loss, g, x


@pytest.mark.parametrize("tag,n,h,w,cfg,nc", [("n3_96x128", 3, 96, 128, '512', 1000),
                                             ("n2_64_cfg320", 2, 64, 64, '320', 10)])
def test_oracle_matches_golden(tag, n, h, w, cfg, nc):
    fx = np.load(os.path.join(GOLD, f"step_{tag}.npz"))
    sd, tr, logits, loss, g, x = _run_case(n, h, w, cfg, nc)
    np.testing.assert_allclose(logits.numpy(), fx["logits"], rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    assert list(fx["names"]) == tr.names
    gn = np.array([g[k].double().norm().item() for k in tr.names])
    bias = np.array([k.endswith("conv.bias") for k in tr.names])
    # conv-bias gradients are analytically zero (rounding noise): not comparable across machines
    np.testing.assert_allclose(gn[~bias], fx["grad_norm"][~bias], rtol=2e-2)
    nbt = [int(sd[k]) for k in sd if k.endswith("num_batches_tracked")]
    assert nbt == list(fx["nbt"])
    with torch.no_grad():
        ev = O.forward(sd, x, False, cfg, nc)
    np.testing.assert_allclose(ev.numpy(), fx["eval_logits"], rtol=1e-4, atol=1e-6)


def test_state_dict_layout():
    torch.manual_seed(0)
    sd = O.init_state_dict()
    assert len(sd) == 403
    assert len({v.data_ptr() for v in sd.values() if v.numel() > 1 or True}) <= 193 + 27
    names = O.unique_param_names(sd)
    assert len(names) == 112
    assert sum(sd[n].numel() for n in names) == 2218400
    # shared blocks: every alias of a repeated MBConv_block is the same tensor
    assert sd["features.2.sequence.0.sequence.1.conv.weight"] is sd["features.2.sequence.2.sequence.1.conv.weight"]


def test_num_batches_tracked_increments():
    torch.manual_seed(0)
    sd = O.init_state_dict()
    x, _ = O.synthetic_batch(2, 64, 64)
    O.forward(sd, x, True, dropout_masks="off")
    got = {k: int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")}
    assert got["features.0.bn.num_batches_tracked"] == 1
    for s, layers in zip(range(2, 8), (3, 3, 3, 2, 4, 1)):
        assert got[f"features.{s}.sequence.0.sequence.0.bn.num_batches_tracked"] == layers
        assert got[f"features.{s}.sequence.{layers}.bn.num_batches_tracked"] == 1


def test_bn_closed_form_matches_autograd():
    torch.manual_seed(1)
    z = torch.randn(4, 16, 9, 7, dtype=torch.float64) * 0.3 + 0.5
    z.requires_grad_(True)
    gamma = torch.rand(16, dtype=torch.float64) + 0.5
    beta = torch.randn(16, dtype=torch.float64) * 0.1
    a = torch.relu(torch.nn.functional.batch_norm(z, None, None, gamma, beta, True, 0.1, O.BN_EPS))
    dA = torch.randn_like(a)
    gamma.requires_grad_(True); beta.requires_grad_(True)
    a = torch.relu(torch.nn.functional.batch_norm(z, None, None, gamma, beta, True, 0.1, O.BN_EPS))
    a.backward(dA)
    dZ, dg, db = O.conv_block_backward_explicit(z.detach(), dA, gamma.detach(), beta.detach())
    torch.testing.assert_close(dZ, z.grad, rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(dg, gamma.grad, rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(db, beta.grad, rtol=1e-9, atol=1e-11)
    s, t, mean, var, unb = O.bn_train_explicit(z.detach(), gamma.detach(), beta.detach())
    torch.testing.assert_close(torch.relu(s[None, :, None, None] * z.detach() + t[None, :, None, None]), a.detach())


def test_adam_matches_torch():
    torch.manual_seed(2)
    p = torch.randn(1000); g = [torch.randn(1000) for _ in range(4)]
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([q], lr=1e-3)
    m, v = torch.zeros(1000), torch.zeros(1000)
    pp = p.clone()
    for i, gi in enumerate(g):
        q.grad = gi.clone(); opt.step()
        O.adam_step(pp, gi, m, v, i + 1)
    assert torch.equal(pp, q.data)
