"""Pin the oracle against the LIVE reference and mint golden fixtures.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py            # asserts oracle == reference, writes tests/golden/*.npz

Checks (all must be bit-identical, same torch build / same CPU kernels):
  * state_dict key order, shapes, aliasing groups, init values (torch.manual_seed(42))
  * logits, loss and every parameter gradient for N=8, 3x224x224 (Dropout modules in eval())
  * BN buffers after the forward, num_batches_tracked increments
  * 3 Adam steps with Dropout ACTIVE (same RNG stream) -> loss trajectory
  * a small odd-shaped case (N=3, 96x128) incl. eval-mode forward
Fixtures written: small tensors only (logits, losses, per-parameter gradient norms + a strided
sample of gradient values, BN buffers), < 1 MB total.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import mnasnet_oracle as O  # noqa: E402


def build_ref(seed=42, num_classes=1000, cfg='512'):
    from models.classifiers import load_model, FineTuneModelPool  # the reference (read-only)
    torch.manual_seed(seed)
    return FineTuneModelPool(load_model('mnasnet'), 'mnasnet', num_classes, cfg)


def set_dropout_eval(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()


def grad_sample(g, n=64):
    f = g.reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx]


def case(n, h, w, cfg='512', num_classes=1000, tag=""):
    ref = build_ref(cfg=cfg, num_classes=num_classes)
    torch.manual_seed(42)
    sd = O.init_state_dict(num_classes, cfg)
    rsd = ref.state_dict()
    assert list(rsd.keys()) == list(sd.keys()), "state_dict key order differs"
    for k in rsd:
        assert rsd[k].shape == sd[k].shape and torch.equal(rsd[k], sd[k]), f"init mismatch {k}"
    # aliasing groups
    def groups(d):
        g = {}
        for k, v in d.items():
            g.setdefault(v.data_ptr() if v.numel() else id(v), []).append(k)
        return sorted(tuple(v) for v in g.values())
    assert groups(rsd) == groups(sd), "aliasing differs"

    x, t = O.synthetic_batch(n, h, w, num_classes)
    ref.train(); set_dropout_eval(ref)
    out = ref(x); loss = torch.nn.CrossEntropyLoss()(out, t)
    ref.zero_grad(); loss.backward()
    rg = {k: p.grad for k, p in ref.named_parameters()}

    tr = O.Trainer(sd, classifier_config=cfg, num_classes=num_classes)
    logits, oloss, og = tr.grads(x, t, dropout_masks="off")
    assert torch.equal(out.detach(), logits), "logits differ"
    assert torch.equal(loss.detach(), oloss), "loss differs"
    assert list(rg.keys()) == tr.names
    for k in rg:
        assert torch.equal(rg[k], og[k]), f"grad differs {k}"
    rsd2 = ref.state_dict()
    for k in rsd2:
        if "running" in k or "num_batches" in k:
            assert torch.equal(rsd2[k], sd[k]), f"buffer differs {k}"
    # eval forward
    ref.eval()
    with torch.no_grad():
        eo = ref(x)
        oo = O.forward(sd, x, False, cfg, num_classes)
    assert torch.equal(eo, oo), "eval logits differ"

    fx = {"logits": logits.numpy(), "loss": np.float64(oloss.item()), "eval_logits": oo.numpy(),
          "names": np.array(tr.names),
          "grad_norm": np.array([og[k].double().norm().item() for k in tr.names]),
          "grad_sample": np.stack([np.pad(grad_sample(og[k]).numpy(), (0, 64 - min(64, og[k].numel())))
                                   for k in tr.names]),
          "bn_keys": np.array([k for k in sd if "running" in k]),
          "nbt": np.array([int(sd[k]) for k in sd if k.endswith("num_batches_tracked")])}
    for k in sd:
        if "running" in k:
            fx["buf/" + k] = sd[k].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"step_{tag}.npz"), **fx)
    print(f"[ok] {tag}: loss {oloss.item():.9f}  max|logit| {logits.abs().max():.4f}")


def adam_trajectory():
    ref = build_ref()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, ref.parameters()), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    x, t = O.synthetic_batch(8, 224, 224)
    ref.train()
    torch.manual_seed(7)
    rl = []
    for _ in range(3):
        out = ref(x); loss = crit(out, t)
        opt.zero_grad(); loss.backward(); opt.step()
        rl.append(loss.item())
    torch.manual_seed(42)
    sd = O.init_state_dict()
    tr = O.Trainer(sd)
    torch.manual_seed(7)
    ol = [tr.step(x, t, dropout_masks=None)[1].item() for _ in range(3)]
    assert rl == ol, (rl, ol)
    rsd = ref.state_dict()
    for k in rsd:
        assert torch.equal(rsd[k], sd[k].detach()), f"post-Adam mismatch {k}"
    # dropout-off trajectory for the GPU parity test (T5)
    torch.manual_seed(42)
    sd = O.init_state_dict()
    tr = O.Trainer(sd)
    off = [tr.step(x, t, dropout_masks="off")[1].item() for _ in range(3)]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adam_traj.npz"),
                        loss_dropout_on_seed7=np.array(ol), loss_dropout_off=np.array(off),
                        bn_rm_f0=sd["features.0.bn.running_mean"].detach().numpy(),
                        bn_rv_f0=sd["features.0.bn.running_var"].detach().numpy(),
                        fc_w_sample=sd["classifier.4.weight"].detach().reshape(-1)[::5000].numpy())
    print("[ok] adam trajectory", ol, off)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(8)
    case(8, 224, 224, tag="n8_224")
    case(3, 96, 128, tag="n3_96x128")
    case(2, 64, 64, cfg='320', num_classes=10, tag="n2_64_cfg320")
    adam_trajectory()
    print("oracle pinned against the live reference; fixtures in tests/golden/")
