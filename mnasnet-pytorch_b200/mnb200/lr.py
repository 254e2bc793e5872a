"""Learning-rate regimes of the reference's training script as plain host functions.

The fused step reads the learning rate from a device scalar (`Engine.train_step_graph(..., lr=...)`), so a schedule
is just a number computed on the host before each replay -- no optimizer object is involved.  Restated regimes
(`args.lr_regime`, src/train.py:282-302):

* 'auto_decay'    ExponentialLR(gamma=0.99), stepped once per epoch            (src/train.py:284-286, 334-335)
* 'plateau_decay' ReduceLROnPlateau(mode='min', factor=0.5, patience=5)        (src/train.py:289-293, 336-337)
* 'clr'           CyclicLR(base_lr=1e-4, max_lr=1e-2, step_size=1200, mode='exp_range', gamma=0.95), stepped once per
                  batch                                                        (src/train.py:296-301, 480; utils/cyclic_lr.py)
* the fixed rule lr = lr0 * 0.9 ** ((epoch + 1) // 50)                         (src/train.py:674-679)
"""
import math


def exponential(lr0: float, epochs_done: int, gamma: float = 0.99) -> float:
    """ExponentialLR after `epochs_done` calls of scheduler.step()."""
    return lr0 * gamma ** epochs_done


def epoch_decay(lr0: float, epoch: int) -> float:
    return lr0 * (0.9 ** ((epoch + 1) // 50))


def cyclic(iteration: int, base_lr: float = 1e-4, max_lr: float = 1e-2, step_size: int = 1200,
           mode: str = "exp_range", gamma: float = 0.95) -> float:
    """Cyclical learning rate at batch `iteration` (utils/cyclic_lr.py:126-140)."""
    step = float(step_size)
    cycle = math.floor(1 + iteration / (2 * step))
    x = abs(iteration / step - 2 * cycle + 1)
    height = (max_lr - base_lr) * max(0.0, 1 - x)
    if mode == "triangular":
        scale = 1.0
    elif mode == "triangular2":
        scale = 1 / (2.0 ** (cycle - 1))
    elif mode == "exp_range":
        scale = gamma ** iteration
    else:
        raise ValueError("mode must be triangular, triangular2 or exp_range")
    return base_lr + height * scale


class ReduceOnPlateau:
    """ReduceLROnPlateau(mode='min', threshold_mode='rel') with torch's defaults for everything the reference does
    not set (threshold 1e-4, cooldown 0, min_lr 0, eps 1e-8)."""

    def __init__(self, lr: float, factor: float = 0.5, patience: int = 5, threshold: float = 1e-4,
                 cooldown: int = 0, min_lr: float = 0.0, eps: float = 1e-8):
        self.lr, self.factor, self.patience, self.threshold = lr, factor, patience, threshold
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best = math.inf
        self.num_bad = 0
        self.cooldown_counter = 0

    def step(self, metric: float) -> float:
        if metric < self.best * (1 - self.threshold):
            self.best = metric
            self.num_bad = 0
        else:
            self.num_bad += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad = 0
        if self.num_bad > self.patience:
            new_lr = max(self.lr * self.factor, self.min_lr)
            if self.lr - new_lr > self.eps:
                self.lr = new_lr
            self.cooldown_counter = self.cooldown
            self.num_bad = 0
        return self.lr


class Schedule:
    """`args.lr_regime` as one object: `.lr` is the rate for the next batch; call `.batch_step()` after every batch
    (src/train.py:480) and `.epoch_step(val_loss)` after every epoch (src/train.py:334-337)."""

    def __init__(self, regime, lr: float):
        if regime not in (None, "auto_decay", "plateau_decay", "clr"):
            raise ValueError(f"unknown lr regime {regime!r}")
        self.regime, self.lr0 = regime, lr
        # the reference's CyclicLR sets iteration 0 in its constructor and AGAIN at the first batch_step()
        # (utils/cyclic_lr.py:113-120): the counter restarts at -1
        self.epochs_done, self.iteration = 0, -1
        self.plateau = ReduceOnPlateau(lr) if regime == "plateau_decay" else None
        self.lr = cyclic(0) if regime == "clr" else lr

    def batch_step(self) -> float:
        if self.regime == "clr":
            self.iteration += 1
            self.lr = cyclic(self.iteration)
        return self.lr

    def epoch_step(self, val_loss: float = None) -> float:
        self.epochs_done += 1
        if self.regime == "auto_decay":
            self.lr = exponential(self.lr0, self.epochs_done)
        elif self.regime == "plateau_decay":
            self.lr = self.plateau.step(val_loss)
        return self.lr
