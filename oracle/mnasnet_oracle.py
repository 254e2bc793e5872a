"""CPU oracle: a functional restatement of the reference's MNASNet training step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing in the product imports this.

The reference (snakers4/mnasnet-pytorch) is pure Python; its arithmetic lives in the third-party
dependency **PyTorch** (no pinned version in the reference: README.md:88-91 says "PyTorch 0.4",
this image has torch 2.11.0+cu128, CPU kernels = oneDNN conv + native batch-norm).  This file
restates the reference's *call sites* as a flat, table-driven functional program over a plain
``dict`` of tensors keyed exactly like the reference ``state_dict``:

  * architecture tables            <- src/models/mnasnet.py:175-193, src/models/classifiers.py:45-89
  * ``init_state_dict``            <- src/models/mnasnet.py:197-209 (+ construction-order RNG draws)
  * ``conv_block``                 <- src/models/mnasnet.py:37-62   (conv+bias -> BN -> ReLU)
  * ``forward``                    <- src/models/mnasnet.py:99-103,131-137,171-173,211-213,
                                      src/models/classifiers.py:107-111
  * ``loss_fn``                    <- src/train.py:277,435          (CrossEntropyLoss, mean)
  * ``adam_step``                  <- src/train.py:219-221,440      (torch.optim.Adam defaults)
  * ``train_step``                 <- src/train.py:433-440
  * ``bn_train_explicit`` / ``conv_block_backward_explicit`` : the closed-form math contract the
    fused CUDA kernels implement, checked against autograd in tests.

Parity pinning: the reference ships NO tests / golden vectors ("parity unpinned" by the reference
itself).  This oracle is pinned instead against the LIVE reference imported from /root/reference
in the build container: ``oracle/make_golden.py`` asserts bit-identical init, logits, loss and
gradients between this file and the reference modules and writes tests/golden/*.npz, which the
CPU test-suite re-checks on every run (the reference itself cannot travel to the GPU box).
"""
from __future__ import annotations

import math
import time
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default (mnasnet.py:55)
BN_MOMENTUM = 0.1    # ConvBlock's own `momentum` argument is ignored (mnasnet.py:46,55)

# (in, out, channel_factor, layers, kernel, reduce)      src/models/mnasnet.py:181-192
STAGES = [
    (16, 24, 3, 3, 3, True),
    (24, 40, 3, 3, 5, True),
    (40, 80, 6, 3, 5, True),
    (80, 96, 6, 2, 3, False),
    (96, 192, 6, 4, 5, True),
    (192, 320, 6, 1, 3, False),
]
FINAL_FEATURE_MAP = 320   # classifiers.py:48

# classifier_config -> list of ("dropout", p) | ("linear", in, out) | ("relu",)   classifiers.py:56-89
def head_spec(classifier_config: str, num_classes: int):
    c = FINAL_FEATURE_MAP
    if classifier_config == '256':
        return [("dropout", 0.5), ("linear", c, 256), ("relu",), ("dropout", 0.5), ("linear", 256, num_classes)]
    if classifier_config == '512_256':
        return [("dropout", 0.5), ("linear", c, 512), ("relu",), ("dropout", 0.5), ("linear", 512, 256),
                ("relu",), ("dropout", 0.5), ("linear", 256, num_classes)]
    if classifier_config == '320':
        return [("dropout", 0.2), ("linear", c, num_classes)]
    if classifier_config == '512':
        return [("dropout", 0.5), ("linear", c, 512), ("relu",), ("dropout", 0.5), ("linear", 512, num_classes)]
    raise ValueError("Finetuning not supported on this architecture yet")   # classifiers.py:89


class CB:
    """One unique ConvBlock (conv+bias, BN, ReLU) of the network."""
    __slots__ = ("keys", "cin", "cout", "k", "stride", "pad", "groups")

    def __init__(self, keys, cin, cout, k, stride, pad, groups):
        self.keys, self.cin, self.cout, self.k = keys, cin, cout, k
        self.stride, self.pad, self.groups = stride, pad, groups

    @property
    def key(self):          # canonical (first) state_dict prefix
        return self.keys[0]

    @property
    def kind(self):
        if self.groups > 1:
            return "dw"
        return "pw" if self.k == 1 else "dense"


def conv_blocks(cut_channels_first: bool = False) -> List[CB]:
    """Unique ConvBlocks in ``modules()`` order (== init order, mnasnet.py:197-209)."""
    out = [CB(["features.0"], 3, 32, 3, 2, 1, 1),                       # mnasnet.py:179
           CB(["features.1.sequence.0"], 32, 32, 3, 1, 1, 32),          # mnasnet.py:180, 86-91
           CB(["features.1.sequence.1"], 32, 16, 1, 1, 0, 1)]           # mnasnet.py:92-95
    for si, (cin, cout, f, layers, k, reduce) in enumerate(STAGES):
        s = si + 2
        cb = cout if cut_channels_first else cin                        # mnasnet.py:150-153
        if cut_channels_first:
            trans_idx, blk_idx = 0, list(range(1, layers + 1))
        else:
            trans_idx, blk_idx = layers, list(range(layers))            # reversed, mnasnet.py:165-168
        trans = CB([f"features.{s}.sequence.{trans_idx}"], cin, cout, 3, 2 if reduce else 1, 1, 1)
        blk = [CB([f"features.{s}.sequence.{j}.sequence.0" for j in blk_idx], cb, cb * f, 1, 1, 0, 1),
               CB([f"features.{s}.sequence.{j}.sequence.1" for j in blk_idx], cb * f, cb * f, k, 1, k // 2, cb * f),
               CB([f"features.{s}.sequence.{j}.sequence.2" for j in blk_idx], cb * f, cb, 1, 1, 0, 1)]
        out += ([trans] + blk) if cut_channels_first else (blk + [trans])
    return out


# ----------------------------------------------------------------------------------------------
# init  (mnasnet.py:197-209, with the RNG draws made by module construction replayed in order)
# ----------------------------------------------------------------------------------------------
def _default_conv_draws(cout, cin_g, k):
    """RNG consumed by nn.Conv2d.__init__ (reset_parameters): kaiming_uniform_(a=sqrt 5) + bias."""
    w = torch.empty(cout, cin_g, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = cin_g * k * k
    bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
    torch.nn.init.uniform_(torch.empty(cout), -bound, bound)


def _default_linear(cin, cout):
    w = torch.empty(cout, cin)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = torch.empty(cout)
    bound = 1 / math.sqrt(cin)
    torch.nn.init.uniform_(b, -bound, bound)
    return w, b


def init_state_dict(num_classes=1000, classifier_config='512', cut_channels_first=False,
                    dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Build the reference's state_dict (403 keys for the default head; aliased shared blocks) using
    the global torch RNG exactly as ``FineTuneModelPool(load_model('mnasnet'), ...)`` would."""
    # 1) construction-order draws (discarded afterwards by init_params, but they advance the RNG)
    _default_conv_draws(32, 3, 3)                                   # features.0
    for _ in range(2):                                              # SepConv builds the pair twice; `* repeat`
        _default_conv_draws(32, 1, 3)                               # (repeat=0) drops the first pair
        _default_conv_draws(32 if _ == 0 else 16, 32, 1)            # mnasnet.py:76-95
    for (cin, cout, f, layers, k, reduce) in STAGES:
        cb = cout if cut_channels_first else cin
        _default_conv_draws(cout, cin, 3)                           # transition built first, mnasnet.py:157-161
        _default_conv_draws(cb * f, cb, 1)
        _default_conv_draws(cb * f, 1, k)
        _default_conv_draws(cb, cb * f, 1)
    # 2) init_params over modules() order
    uniq = OrderedDict()
    for cb in conv_blocks(cut_channels_first):
        w = torch.empty(cb.cout, cb.cin // cb.groups, cb.k, cb.k)
        torch.nn.init.kaiming_normal_(w, mode='fan_out')            # mnasnet.py:200
        uniq[cb.key] = {
            "conv.weight": w, "conv.bias": torch.zeros(cb.cout),
            "bn.weight": torch.ones(cb.cout), "bn.bias": torch.zeros(cb.cout),
            "bn.running_mean": torch.zeros(cb.cout), "bn.running_var": torch.ones(cb.cout),
            "bn.num_batches_tracked": torch.tensor(0, dtype=torch.long),
        }
    # 3) classifier built afterwards with default nn.Linear init (classifiers.py:56-89)
    head = OrderedDict()
    for i, op in enumerate(head_spec(classifier_config, num_classes)):
        if op[0] == "linear":
            w, b = _default_linear(op[1], op[2])
            head[f"classifier.{i}.weight"] = w
            head[f"classifier.{i}.bias"] = b
    # 4) emit keys in the reference's state_dict order (aliases share storage)
    sd = OrderedDict()
    order = _state_dict_prefix_order(cut_channels_first)
    alias = {}
    for cb in conv_blocks(cut_channels_first):
        for kx in cb.keys:
            alias[kx] = cb.key
    for prefix in order:
        for name, t in uniq[alias[prefix]].items():
            sd[f"{prefix}.{name}"] = t
    sd.update(head)
    if dtype != torch.float32:
        conv = {}
        for k_, v in sd.items():
            if v.is_floating_point():
                if id(v) not in conv:
                    conv[id(v)] = v.to(dtype)
                sd[k_] = conv[id(v)]
    return sd


def _state_dict_prefix_order(cut_channels_first=False) -> List[str]:
    order = ["features.0", "features.1.sequence.0", "features.1.sequence.1"]
    for si, (cin, cout, f, layers, k, reduce) in enumerate(STAGES):
        s = si + 2
        n = layers + 1
        for j in range(n):
            is_trans = (j == 0) if cut_channels_first else (j == layers)
            if is_trans:
                order.append(f"features.{s}.sequence.{j}")
            else:
                order += [f"features.{s}.sequence.{j}.sequence.{q}" for q in range(3)]
    return order


def unique_param_names(sd) -> List[str]:
    """Parameter names as ``named_parameters()`` would list them (first alias only)."""
    seen, out = set(), []
    for k_, v in sd.items():
        if k_.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        if id(v) in seen:
            continue
        seen.add(id(v))
        out.append(k_)
    return out


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def conv_block(sd, prefix, x, stride, pad, groups, train, capture=None):
    """ConvBlock.forward (mnasnet.py:58-62): relu(bn(conv(x)+b)); BN buffers updated in train mode."""
    z = F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=stride, padding=pad,
                 groups=groups)
    if train:
        sd[prefix + ".bn.num_batches_tracked"] += 1          # torch:nn/modules/batchnorm.py:163-178
    a = F.batch_norm(z, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                     sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], train, BN_MOMENTUM, BN_EPS)
    a = F.relu(a)
    if capture is not None:
        for t in (x, z, a):
            if t.requires_grad:
                t.retain_grad()
        capture.append({"prefix": prefix, "x": x, "z": z, "a": a,
                        "stride": stride, "pad": pad, "groups": groups})
    return a


def features_forward(sd, x, train=True, capture=None, cut_channels_first=False):
    x = conv_block(sd, "features.0", x, 2, 1, 1, train, capture)                        # mnasnet.py:179
    x = conv_block(sd, "features.1.sequence.0", x, 1, 1, 32, train, capture)            # mnasnet.py:180
    x = conv_block(sd, "features.1.sequence.1", x, 1, 0, 1, train, capture)
    for si, (cin, cout, f, layers, k, reduce) in enumerate(STAGES):
        s = si + 2
        cb = cout if cut_channels_first else cin

        def trans(x, j):
            return conv_block(sd, f"features.{s}.sequence.{j}", x, 2 if reduce else 1, 1, 1, train, capture)

        def block(x, j):                                                                # mnasnet.py:131-137
            p = f"features.{s}.sequence.{j}.sequence."
            y = conv_block(sd, p + "0", x, 1, 0, 1, train, capture)
            y = conv_block(sd, p + "1", y, 1, k // 2, cb * f, train, capture)
            y = conv_block(sd, p + "2", y, 1, 0, 1, train, capture)
            return x + y
        if cut_channels_first:
            x = trans(x, 0)
            for j in range(1, layers + 1):
                x = block(x, j)
        else:
            for j in range(layers):
                x = block(x, j)
            x = trans(x, layers)
    return x


def head_forward(sd, f, train, classifier_config='512', num_classes=1000, dropout_masks=None):
    """pooling + classifier (classifiers.py:107-111).  ``dropout_masks``: None -> torch RNG dropout in
    train mode; "off" -> identity (Dropout modules in eval()); list of {0,1} keep-masks -> injected."""
    f = F.adaptive_avg_pool2d(f, 1)
    y = f.view(f.size(0), -1)
    di = 0
    for i, op in enumerate(head_spec(classifier_config, num_classes)):
        if op[0] == "dropout":
            if not train or dropout_masks == "off":
                pass
            elif dropout_masks is None:
                y = F.dropout(y, op[1], True)
            else:
                y = y * dropout_masks[di].to(y.dtype) / (1.0 - op[1])
            di += 1
        elif op[0] == "linear":
            y = F.linear(y, sd[f"classifier.{i}.weight"], sd[f"classifier.{i}.bias"])
        else:
            y = F.relu(y)
    return y


def forward(sd, x, train=True, classifier_config='512', num_classes=1000, dropout_masks=None, capture=None,
            cut_channels_first=False):
    f = features_forward(sd, x, train, capture, cut_channels_first)
    if capture is not None:
        capture.append({"prefix": "features_out", "a": f})
    return head_forward(sd, f, train, classifier_config, num_classes, dropout_masks)


def loss_fn(logits, target):
    return F.cross_entropy(logits, target)                  # train.py:277 (mean reduction)


# ----------------------------------------------------------------------------------------------
# optimizer  (torch.optim.Adam defaults: betas (0.9,0.999), eps 1e-8, wd 0, amsgrad False)
# ----------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """Single-tensor Adam, torch:optim/adam.py `_single_tensor_adam`; `step` counts from 1. In place."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


class Trainer:
    """src/train.py:419-440 on synthetic tensors: model.train(); forward; CE; zero_grad; backward; Adam."""

    def __init__(self, sd, lr=1e-3, classifier_config='512', num_classes=1000):
        self.sd = sd
        self.lr = lr
        self.cfg, self.nc = classifier_config, num_classes
        self.names = unique_param_names(sd)
        for n in self.names:
            sd[n].requires_grad_(True)
        self.m = {n: torch.zeros_like(sd[n]) for n in self.names}
        self.v = {n: torch.zeros_like(sd[n]) for n in self.names}
        self.t = 0

    def grads(self, x, target, dropout_masks="off", capture=None):
        for n in self.names:
            self.sd[n].grad = None
        logits = forward(self.sd, x, True, self.cfg, self.nc, dropout_masks, capture)
        loss = loss_fn(logits, target)
        loss.backward()
        return logits.detach(), loss.detach(), {n: self.sd[n].grad for n in self.names}

    def step(self, x, target, dropout_masks="off"):
        logits, loss, g = self.grads(x, target, dropout_masks)
        self.t += 1
        with torch.no_grad():
            for n in self.names:
                adam_step(self.sd[n], g[n], self.m[n], self.v[n], self.t, self.lr)
        return logits, loss


def synthetic_batch(n, h, w, num_classes=1000, seed=0, dtype=torch.float32):
    """x ~ N(0,1) NCHW, targets uniform in [0, num_classes)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g).to(dtype)
    t = torch.randint(0, num_classes, (n,), generator=g)
    return x, t


# ----------------------------------------------------------------------------------------------
# closed-form math contract of the fused kernels -- checked against autograd
# ----------------------------------------------------------------------------------------------
def bn_train_explicit(z, gamma, beta, eps=BN_EPS):
    """Returns (scale s, shift t, mean, biased var, unbiased var): A = relu(s*Z+t) per channel (NCHW z)."""
    m = z.numel() // z.shape[1]
    mean = z.mean(dim=(0, 2, 3))
    var = ((z - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    s = gamma / torch.sqrt(var + eps)
    t = beta - mean * s
    return s, t, mean, var, var * m / max(m - 1, 1)


def conv_block_backward_explicit(z, dA, gamma, beta, eps=BN_EPS):
    """Given raw conv output Z and dA (grad wrt post-ReLU output) returns (dZ, dgamma, dbeta) using the
    per-channel coefficient form the CUDA kernels use:  dZ = a*G + b*Z + c,  G = dA*[s*Z+t>0]."""
    s, t, mean, var, _ = bn_train_explicit(z, gamma, beta, eps)
    m = z.numel() // z.shape[1]
    bc = lambda v: v[None, :, None, None]
    G = dA * ((bc(s) * z + bc(t)) > 0)
    sum_g = G.sum(dim=(0, 2, 3))
    sum_gz = (G * z).sum(dim=(0, 2, 3))
    inv_std = 1.0 / torch.sqrt(var + eps)
    dbeta = sum_g
    dgamma = inv_std * (sum_gz - mean * sum_g)
    # zhat = (Z-mean)*inv_std ; dZ = s*(G - dbeta/m - zhat*dgamma/m)
    a = s
    b = -s * inv_std * dgamma / m
    c = -s * dbeta / m + s * inv_std * mean * dgamma / m
    dZ = bc(a) * G + bc(b) * z + bc(c)
    return dZ, dgamma, dbeta


# ----------------------------------------------------------------------------------------------
# CPU baseline timing (bench.py cpu_baseline / --impl reference)
# ----------------------------------------------------------------------------------------------
def time_cpu_steps(n=8, h=224, w=224, steps=5, warmup=2, threads=None, seed=42):
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    sd = init_state_dict()
    tr = Trainer(sd)
    x, t = synthetic_batch(n, h, w)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tr.step(x, t, dropout_masks=None)     # Dropout active, like train.py
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return {"img_per_s": n / med, "s_per_step": med, "threads": torch.get_num_threads(), "times": times}
