"""Validation forward around the hot path: `validate()` of the reference runs the
model in eval mode under no_grad, then CrossEntropyLoss and precision@1/@5 per batch, averaged with an AverageMeter
weighted by batch size (src/train.py:532-641, 687-700).

The eval forward is the engine's `fwd_eval` program (BatchNorm running statistics folded into per-channel scale /
shift, Dropout off, no statistics, nothing saved for backward); loss and top-k are a handful of torch ops on the
device logits (N x classes) with NO host synchronisation per batch -- the meters accumulate on the device and are read
once at the end."""
from typing import Iterable, Sequence, Tuple

import torch
import torch.nn.functional as F


def topk_hits(logits: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1, 5)):
    """Number of samples whose target is among the k highest logits, per k (device tensor of len(topk) floats).
    Same selection rule as the reference's accuracy(): `output.topk(maxk, 1, True, True)` (src/train.py:692)."""
    maxk = min(max(topk), logits.shape[1])
    pred = logits.topk(maxk, dim=1, largest=True, sorted=True).indices          # N x maxk
    hit = pred.eq(target.view(-1, 1))
    return torch.stack([hit[:, :min(k, maxk)].any(dim=1).sum() for k in topk]).to(torch.float32)


def accuracy(logits: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1,)):
    """precision@k in percent, one tensor per k (src/train.py:687-700)."""
    n = target.shape[0]
    return list(topk_hits(logits, target, topk) * (100.0 / n))


class DeviceMeter:
    """AverageMeter (src/train.py:657-672) whose sum / count stay on the device until `.avg` is read."""

    def __init__(self, device):
        self.sum = torch.zeros((), device=device, dtype=torch.float64)
        self.count = 0

    def update(self, val: torch.Tensor, n: int = 1):
        self.sum += val.detach().to(torch.float64) * n
        self.count += n

    @property
    def avg(self) -> float:
        return float(self.sum.item() / max(1, self.count))


def eval_logits(engine, x: torch.Tensor) -> torch.Tensor:
    """Eval-mode forward of the lowered module: N x num_classes fp32 device tensor (valid until the next forward)."""
    plan = engine.forward(x, train=False)
    return engine.logits(plan)


def validate(engine, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], topk: Sequence[int] = (1, 5)):
    """One pass over (input, target) batches: {'loss', 'acc1', 'acc5', 'n'} as the reference's validate() returns
    them (src/train.py:560-641).  Inputs may be host tensors (copied with non_blocking=True, train.py:562-563)."""
    dev = engine.device
    loss_m = DeviceMeter(dev)
    acc_m = [DeviceMeter(dev) for _ in topk]
    n_total = 0
    with torch.no_grad():
        for x, target in batches:
            x = x.float().to(dev, non_blocking=True)
            target = target.to(dev, non_blocking=True)
            logits = eval_logits(engine, x)
            n = target.shape[0]
            loss_m.update(F.cross_entropy(logits, target), n)
            for m, a in zip(acc_m, accuracy(logits, target, topk)):
                m.update(a, n)
            n_total += n
    out = {"loss": loss_m.avg, "n": n_total}
    for k, m in zip(topk, acc_m):
        out[f"acc{k}"] = m.avg
    return out
