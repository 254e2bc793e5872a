"""Checkpoint interchange with the reference's training script.

The reference saves `{'epoch', 'optimizer': optimizer.state_dict(), 'state_dict': model.state_dict(), 'best_loss'}`
with `torch.save` (src/train.py:380-389, 652-655) from a `torch.nn.DataParallel`-wrapped model (src/train.py:202), so
every model key carries a `module.` prefix and the shared MBConv blocks appear once per alias (403 keys, 193 unique
tensors).  On resume it restores `epoch`, `best_loss` and the model (src/train.py:233-246; the optimizer line is
commented out there, the state is still part of the file).

Here the model's parameters are views into one flat fp32 buffer and the Adam moments are two more flat buffers in
backward-ready (reversed registration) order, so the optimizer state is converted to and from the per-parameter layout
of `torch.optim.Adam.state_dict()` in the order the reference builds its optimizer,
`filter(lambda p: p.requires_grad, model.parameters())` (src/train.py:219-221).  Everything in this module is
device-agnostic torch code: it runs on the flat buffers wherever they live."""
import collections
import os
import shutil

import torch
from torch import nn

PREFIX = "module."


def reference_state_dict(module: nn.Module, parallel_prefix: bool = True):
    """`model.state_dict()` as the reference writes it (src/train.py:383): CPU tensors, aliased shared-block keys
    included, `module.` prefix of the DataParallel wrapper when `parallel_prefix`."""
    out = collections.OrderedDict()
    for k, v in module.state_dict().items():
        out[(PREFIX + k) if parallel_prefix else k] = v.detach().to("cpu").clone()
    return out


def strip_prefix(sd):
    """Accepts checkpoints saved with or without the DataParallel `module.` prefix (src/train.py:238-246)."""
    if sd and all(k.startswith(PREFIX) for k in sd):
        return collections.OrderedDict((k[len(PREFIX):], v) for k, v in sd.items())
    return sd


def _alias_groups(module: nn.Module):
    """state_dict keys that name the same tensor (the shared MBConv blocks)."""
    groups = collections.defaultdict(list)
    for k, v in module.state_dict(keep_vars=True).items():
        groups[id(v)].append(k)
    return [g for g in groups.values() if len(g) > 1]


def load_reference_state_dict(module: nn.Module, sd, strict: bool = True):
    """`model.load_state_dict(checkpoint['state_dict'])` (src/train.py:242).  Aliased keys must agree: torch would
    silently keep whichever alias it copies last."""
    sd = strip_prefix(sd)
    for group in _alias_groups(module):
        present = [k for k in group if k in sd]
        for k in present[1:]:
            if not torch.equal(sd[present[0]].cpu(), sd[k].cpu()):
                raise ValueError(f"checkpoint disagrees on shared tensor: {present[0]} vs {k}")
    return module.load_state_dict(sd, strict=strict)


def _trainable_in_registration_order(module: nn.Module):
    # torch de-duplicates shared parameters in .parameters(); this is the reference optimizer's parameter order
    return [p for p in module.parameters() if p.requires_grad]


def adam_state_dict(module: nn.Module, store, step: int, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
    """The flat Adam state (store.m / store.v, `step` optimizer steps taken) in the layout of
    `torch.optim.Adam(...).state_dict()` for the reference's optimizer (src/train.py:219-221, 382)."""
    m, v = store.adam_state()
    params = _trainable_in_registration_order(module)
    state = {}
    for i, p in enumerate(params):
        if step == 0:
            continue                                  # torch creates per-parameter state lazily at the first step
        o, n = store.offsets[id(p)]
        state[i] = {"step": torch.tensor(float(step)),
                    "exp_avg": m[o:o + n].view(p.shape).detach().to("cpu").clone(),
                    "exp_avg_sq": v[o:o + n].view(p.shape).detach().to("cpu").clone()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False,
             "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(module: nn.Module, store, osd) -> int:
    """Inverse of adam_state_dict; returns the step count (0 for a fresh optimizer).  Extra param-group keys written
    by other torch versions (maximize, foreach, capturable, ...) are ignored."""
    m, v = store.adam_state()
    params = _trainable_in_registration_order(module)
    groups = osd["param_groups"]
    order = [i for g in groups for i in g["params"]]
    if len(order) != len(params):
        raise ValueError(f"optimizer state has {len(order)} parameters, the model has {len(params)} trainable ones")
    steps = set()
    m.zero_()
    v.zero_()
    for i, p in zip(order, params):
        st = osd["state"].get(i)
        if st is None:
            continue
        if tuple(st["exp_avg"].shape) != tuple(p.shape):
            raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} != parameter {tuple(p.shape)}")
        o, n = store.offsets[id(p)]
        m[o:o + n].view(p.shape).copy_(st["exp_avg"])
        v[o:o + n].view(p.shape).copy_(st["exp_avg_sq"])
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the flat optimizer keeps one counter")
    return steps.pop() if steps else 0


def save_checkpoint(state, is_best: bool, filename: str, best_filename: str):
    """src/train.py:652-655."""
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, best_filename)


def make_checkpoint(module: nn.Module, engine=None, epoch: int = 0, best_loss: float = float("inf"),
                    lr: float = 1e-3, parallel_prefix: bool = True):
    """The dict the reference saves at the end of an epoch (src/train.py:380-385).  `epoch` is the NEXT epoch."""
    ck = {"epoch": epoch, "state_dict": reference_state_dict(module, parallel_prefix), "best_loss": best_loss}
    if engine is not None and getattr(engine, "optimizer", "adam") == "adam":
        ck["optimizer"] = adam_state_dict(module, engine.store, engine.host_step, lr)
    return ck


def resume(path_or_dict, module: nn.Module, engine=None, load_optimizer: bool = False):
    """src/train.py:233-246: returns (start_epoch, best_loss).  `load_optimizer` additionally restores the Adam
    moments and step counter into the engine's flat buffers (the reference keeps that line commented out)."""
    ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                         weights_only=False)
    load_reference_state_dict(module, ck["state_dict"])
    if engine is not None:
        engine._check_store()
        if load_optimizer and "optimizer" in ck:
            step = load_adam_state_dict(module, engine.store, ck["optimizer"])
            engine.host_step = step
            engine.dev_step.fill_(step)
    return ck.get("epoch", 0), ck.get("best_loss", float("inf"))
