"""Lowering of the reference's module tree to a static program of libmnb200 kernel launches.

Host side of the drop-in boundary: the reference drives the hot path with six lines
(src/train.py:433-440: ``out = model(input); loss = criterion(out, target); optimizer.zero_grad();
loss.backward(); optimizer.step()``).  ``run_module`` makes ``model(input)`` / ``loss.backward()`` work
unchanged through ``torch.autograd.Function``; ``Engine.train_step`` is the fused variant (loss, backward and
Adam also in our kernels, optional NCCL gradient all-reduce, optional CUDA-graph replay).

Data layout in HBM: activations NHWC (bf16 or fp32).  For every ConvBlock *application* the
program keeps the raw conv output Z plus per-channel (scale, shift, mean, invstd); A = relu(scale*Z+shift)
is never stored -- consumers recompute it on load.  Residual-block inputs/outputs (narrow tensors) are
materialised once.  Parameters / gradients / Adam moments live in flat fp32 buffers ordered
backward-ready-first so NCCL buckets are contiguous slices; nn.Parameters are views into them.
"""
from __future__ import annotations

import os
import weakref
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check

DTYPES = {"fp32": (_lib.MNB_F32, torch.float32), "bf16": (_lib.MNB_BF16, torch.bfloat16)}
IMPLS = {"auto": 0, "simt": 1, "tc": 2, "stream": 3}
BN_EPS_DEFAULT = 1e-5


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mnb200: a CUDA device (B200, sm_100a) is required; there is no CPU fallback")


class Ref:
    """An activation in HBM: tensor t viewed as [N,H,W,C]; if scale is set the VALUE is relu(scale*t+shift)."""
    __slots__ = ("t", "N", "H", "W", "C", "scale", "shift", "nchw")

    def __init__(self, t, N, H, W, C, scale=None, shift=None, nchw=False):
        self.t, self.N, self.H, self.W, self.C = t, N, H, W, C
        self.scale, self.shift, self.nchw = scale, shift, nchw

    @property
    def M(self):
        return self.N * self.H * self.W


class ParamStore:
    """Flat fp32 parameter / gradient buffers (+ flat BN buffers).  Order: reverse of registration order,
    i.e. the order in which gradients become ready in backward (classifier first, stem last)."""
    ALIGN = 8

    def __init__(self, module: nn.Module, device):
        self.device = device
        seen, params = set(), []
        for name, p in module.named_parameters():          # de-duplicated by torch (shared blocks once)
            if id(p) not in seen:
                seen.add(id(p))
                params.append((name, p))
        params = params[::-1]
        self.names = [n for n, _ in params]
        self.params = [p for _, p in params]
        off, self.offsets = 0, {}
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel())
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32)
        for p in self.params:
            o, n = self.offsets[id(p)]
            v = self.flat[o:o + n].view(p.shape)
            v.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = v
        # BN buffers
        fb, ib, seenb = [], [], set()
        for name, b in module.named_buffers():
            if id(b) in seenb:
                continue
            seenb.add(id(b))
            (ib if b.dtype == torch.long else fb).append(b)
        self.fbuf = torch.zeros(max(1, sum(b.numel() for b in fb)), device=device, dtype=torch.float32)
        self.ibuf = torch.zeros(max(1, sum(b.numel() for b in ib)), device=device, dtype=torch.long)
        self._fb, self._ib = fb, ib
        o = 0
        for b in fb:
            v = self.fbuf[o:o + b.numel()].view(b.shape)
            v.copy_(b.data.to(device))
            b.data = v
            o += b.numel()
        o = 0
        for b in ib:
            v = self.ibuf[o:o + b.numel()].view(b.shape)
            v.copy_(b.data.to(device))
            b.data = v
            o += b.numel()
        self.m = None
        self.v = None

    def valid(self):
        """Every parameter and buffer still aliases its slot of the flat stores (a sub-module lowered on its own, or
        model.to(), re-homes them: the engine then rebuilds, see Engine._check_store)."""
        base = self.flat.data_ptr()
        for p in self.params:
            if p.data_ptr() != base + 4 * self.offsets[id(p)][0]:
                return False
        for lst, buf, es in ((self._fb, self.fbuf, 4), (self._ib, self.ibuf, 8)):
            o = buf.data_ptr()
            for b in lst:
                if b.data_ptr() != o:
                    return False
                o += es * b.numel()
        return True

    def trainable_ranges(self):
        """Contiguous [lo, hi) slices of the flat buffers whose parameters have requires_grad (the reference builds
        its optimizer from filter(lambda p: p.requires_grad, ...), src/train.py:219): after
        FineTuneModelPool.freeze() (classifiers.py:94-99) this is the classifier slice alone."""
        out = []
        for p in self.params:
            if not p.requires_grad:
                continue
            o, n = self.offsets[id(p)]
            hi = o + (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            if out and out[-1][1] == o:
                out[-1][1] = hi
            else:
                out.append([o, hi])
        return [tuple(r) for r in out]

    def trainable_sig(self):
        return tuple(p.requires_grad for p in self.params)

    def gptr(self, p):
        return self.grad.data_ptr() + 4 * self.offsets[id(p)][0]

    def grad_views(self, src=None):
        src = self.grad if src is None else src
        out = {}
        for p in self.params:
            o, n = self.offsets[id(p)]
            out[id(p)] = src[o:o + n].view(p.shape)
        return out

    def adam_state(self):
        if self.m is None:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
        return self.m, self.v

    def buckets(self):
        """Contiguous slices of the flat gradient buffer in backward-ready order, cut at top-level-stage
        boundaries: classifier+features.7 | features.6+features.5 | rest."""
        groups = [("classifier", "features.7"), ("features.6", "features.5")]

        def gidx(name):
            parts = name.split(".")
            s_ = parts[0] if parts[0] != "features" else "features." + parts[1]
            for i, g in enumerate(groups):
                if s_ in g:
                    return i
            return len(groups)
        edges, prev = [0], None
        for name, p in zip(self.names, self.params):
            g = gidx(name)
            if prev is not None and g != prev:
                edges.append(self.offsets[id(p)][0])
            prev = g
        edges.append(self.total)
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1) if edges[i + 1] > edges[i]]


class _ConvApp:
    """One application of a ConvBlock inside a plan."""
    __slots__ = ("cb", "inp", "z", "stats", "sums", "scale", "shift", "mean", "invstd", "coef", "kind",
                 "Ho", "Wo", "Cout", "k", "stride", "pad", "need_dgrad", "m", "in_bytes", "out_bytes", "label",
                 "index", "reduce_fused", "in_layout", "in_sc", "in_sh")


class Plan:
    """Static program for one input shape (N,H,W): buffers + forward/backward op lists."""

    def __init__(self, eng: "Engine", N: int, H: int, W: int, input_u8: bool = False):
        self.eng, self.N, self.H, self.W = eng, N, H, W
        self.input_u8 = input_u8     # network input is N x H x W x 3 uint8: the stem normalises it
        self.dev = eng.device
        self.code, self.tdtype = DTYPES[eng.dtype]
        self.esize = 4 if eng.dtype == "fp32" else 2
        self.keep = []              # tensors owned by the plan
        self.fwd, self.fwd_eval, self.bwd = [], [], []
        self.apps: List[_ConvApp] = []
        self.tape = []
        self.generation = 0
        self.prog_graphs = {}          # CUDA graphs of the module-path programs (run_graphed)
        self.static_in = None
        self.last_write = {}
        self.packed = {}
        self.n_events = 0
        self.side_last = None
        self.events = None
        self._arena_f, self._arena_d = [], []   # (numel) requests
        self._build()

    # ---- allocation helpers ---------------------------------------------------------------------
    def _act(self, N, H, W, C):
        t = torch.empty((N, H, W, C), device=self.dev, dtype=self.tdtype)
        self.keep.append(t)
        return t

    def _f32(self, *shape):
        t = torch.zeros(shape, device=self.dev, dtype=torch.float32)
        self.keep.append(t)
        return t

    # ---- program construction --------------------------------------------------------------------
    def _build(self):
        eng = self.eng
        mod = eng.module
        N, H, W = self.N, self.H, self.W
        # fp64 arena for BN statistics (fwd stats + bwd sums), zeroed once per step
        self._dstat_chunks = []
        kind = getattr(mod, "_mnb", None)
        if kind == "net":
            ref = Ref(None, N, H, W, 3, nchw=True)
            ref = self._emit(mod.features, ref)
            self._emit_head(mod, ref)
            if self.dropout_masks:
                self._op(self.fwd, "mnb_counter_inc", eng.dev_fwd, label="fwd_counter")
        else:
            cin = eng.in_channels
            # standalone sub-module: convert the NCHW fp32 input to NHWC once, run, convert back
            xin = self._act(N, H, W, cin)
            self.sub_in = xin
            ref = Ref(xin, N, H, W, cin)
            out = self._emit(mod, ref)
            out = self._materialize(out)
            self.sub_out = out
            self.out_nchw = torch.empty((N, out.C, out.H, out.W), device=self.dev, dtype=torch.float32)
            self.keep.append(self.out_nchw)
        # allocate the fp64 arena
        total = sum(n for n, _ in self._dstat_chunks)
        self.dstats = torch.zeros(max(total, 1), device=self.dev, dtype=torch.float64)
        self._dstat_base = self.dstats.data_ptr()
        self._emit_backward()

    def _dstat(self, n):
        """Reserve n doubles in the arena; returns a lazy pointer (resolved after allocation)."""
        off = sum(c for c, _ in self._dstat_chunks)
        h = _Lazy(self, off * 8)
        self._dstat_chunks.append((n, h))
        return h

    def _emit(self, mod, ref: Ref) -> Ref:
        kind = getattr(mod, "_mnb", None)
        if kind == "convblock":
            return self._emit_convblock(mod, ref)
        if kind == "chain":
            for child in mod.sequence:
                ref = self._emit(child, ref)
            return ref
        if kind == "resblock":
            x = self._materialize(ref)
            self.tape.append(("res_begin", x))
            r = x
            for child in mod.sequence:
                r = self._emit(child, r)
            y = self._act(x.N, x.H, x.W, x.C)
            M = x.M
            nb = 3 * M * x.C * self.esize
            self._op(self.fwd, "mnb_bn_relu_apply", r.t, r.scale, r.shift, x.t, y, M, x.C, self.code, nbytes=nb,
                     label="bn_relu_apply+res")
            self._op(self.fwd_eval, "mnb_bn_relu_apply", r.t, r.scale, r.shift, x.t, y, M, x.C, self.code)
            self.tape.append(("res_end", x))
            return Ref(y, x.N, x.H, x.W, x.C)
        if isinstance(mod, nn.Sequential):
            for child in mod:
                ref = self._emit(child, ref)
            return ref
        raise TypeError(f"mnb200 cannot lower module of type {type(mod).__name__}")

    def _materialize(self, ref: Ref) -> Ref:
        if ref.scale is None and not ref.nchw:
            return ref
        if ref.nchw:
            raise RuntimeError("NCHW input can only feed a dense ConvBlock")
        y = self._act(ref.N, ref.H, ref.W, ref.C)
        for ops in (self.fwd, self.fwd_eval):
            self._op(ops, "mnb_bn_relu_apply", ref.t, ref.scale, ref.shift, None, y, ref.M, ref.C, self.code,
                     nbytes=2 * ref.M * ref.C * self.esize, label="bn_relu_apply")
        self.tape.append(("materialize", ref))
        return Ref(y, ref.N, ref.H, ref.W, ref.C)

    def _emit_convblock(self, cb, ref: Ref) -> Ref:
        conv, bn = cb.conv, cb.bn
        k, stride, pad, groups = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.groups
        cin, cout = conv.in_channels, conv.out_channels
        if cin != ref.C:
            raise ValueError(f"channel mismatch: ConvBlock expects {cin}, got {ref.C}")
        Ho = (ref.H + 2 * pad - k) // stride + 1
        Wo = (ref.W + 2 * pad - k) // stride + 1
        a = _ConvApp()
        a.cb, a.inp, a.Ho, a.Wo, a.Cout, a.k, a.stride, a.pad = cb, ref, Ho, Wo, cout, k, stride, pad
        a.kind = "dw" if groups > 1 else "dense"
        if a.kind == "dw" and (groups != cin or cin != cout or stride != 1 or pad != k // 2):
            raise ValueError("only depthwise stride-1 'same' grouped convs are lowered")
        a.z = self._act(ref.N, Ho, Wo, cout)
        a.m = float(ref.N * Ho * Wo)
        a.stats = self._dstat(2 * cout)
        a.sums = self._dstat(2 * cout)
        vec = self._f32(4, cout)
        a.scale, a.shift, a.mean, a.invstd = vec[0], vec[1], vec[2], vec[3]
        a.coef = self._f32(3, cout)
        a.need_dgrad = not ref.nchw        # the stem reads the network input: no data gradient
        eng = self.eng
        x_t = ref.t if not ref.nchw else _InputPtr(self)
        layout = _lib.LAYOUT_NHWC
        in_sc, in_sh = ref.scale, ref.shift
        if ref.nchw:
            layout = _lib.LAYOUT_NHWC_U8 if self.input_u8 else _lib.LAYOUT_NCHW_F32
            if self.input_u8:
                in_sc, in_sh = eng.norm_constants(cin)       # Normalize(mean, std) happens inside the stem kernels
        a.in_layout, a.in_sc, a.in_sh = layout, in_sc, in_sh
        es = self.esize
        in_b = ref.M * cin * ((1 if self.input_u8 else 4) if ref.nchw else es)
        out_b = int(a.m) * cout * es
        a.in_bytes, a.out_bytes = in_b, out_b
        a.label = f"dw{k}x{k}" if a.kind == "dw" else (f"conv{k}x{k}" if k > 1 else "pw1x1")
        self._cur_detail = f"{ref.H}x{ref.W} {cin}->{cout} k{k}s{stride}"
        wpk = self._packed(cb) if (a.kind == "dense" and not ref.nchw and eng.dtype == "bf16") else (None, None)
        for ops, train in ((self.fwd, True), (self.fwd_eval, False)):
            st = a.stats if train else None
            if a.kind == "dense":
                self._op(ops, "mnb_conv_fwd_packed", x_t, in_sc, in_sh, conv.weight, wpk[0], None, a.z, st,
                         ref.N, ref.H, ref.W, cin, cout, k, stride, pad, self.code, layout, eng.impl,
                         nbytes=in_b + out_b, label=a.label + "_fwd")
            else:
                self._op(ops, "mnb_dw_fwd", x_t, ref.scale, ref.shift, conv.weight, None, a.z, st,
                         ref.N, ref.H, ref.W, cin, k, self.code, nbytes=in_b + out_b, label=a.label + "_fwd")
            if train:
                # the conv bias is folded into BN (train-mode BN cancels it): Z is stored without it
                self._op(ops, "mnb_bn_finalize_fb", a.stats, bn.weight, bn.bias, conv.bias, bn.running_mean, bn.running_var,
                         bn.num_batches_tracked, a.scale, a.shift, a.mean, a.invstd, cout, a.m,
                         float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1))
            else:
                self._op(ops, "mnb_bn_eval_coeffs_fb", bn.weight, bn.bias, conv.bias, bn.running_mean, bn.running_var,
                         a.scale, a.shift, cout, float(bn.eps))
        a.index, a.reduce_fused = len(self.apps), False
        self.apps.append(a)
        self.tape.append(("conv", a))
        return Ref(a.z, ref.N, Ho, Wo, cout, a.scale, a.shift)

    def _packed(self, cb):
        """bf16 K-major packings (forward, backward-data) of a dense ConvBlock's weight, refreshed once per
        forward (shared-weight blocks are packed once, before their first application)."""
        key = id(cb)
        if key not in self.packed:
            conv = cb.conv
            n = conv.weight.numel()
            pf = torch.empty(n, device=self.dev, dtype=torch.bfloat16)
            pd = torch.empty(n, device=self.dev, dtype=torch.bfloat16)
            self.keep += [pf, pd]
            self.packed[key] = (pf, pd)
            for ops in (self.fwd, self.fwd_eval):
                self._op(ops, "mnb_pack_weights", conv.weight, pf, pd if ops is self.fwd else None,
                         conv.out_channels, conv.in_channels, conv.kernel_size[0], label="pack_weights")
        return self.packed[key]

    def _emit_head(self, mod, ref: Ref):
        """AdaptiveAvgPool2d(1) + classifier Sequential of Dropout / Linear / ReLU (classifiers.py:56-111)."""
        N, C = ref.N, ref.C
        self.head_in = ref
        self.f = self._f32(N, C)
        children = list(mod.classifier)
        # bf16 + tensor-pipe head: the pooling kernel writes the first Linear's bf16 operand itself (dropout applied),
        # so GAP feeds the FC GEMM without a pass in between; otherwise a plain pooling kernel
        lin0 = next((ch for ch in children if isinstance(ch, nn.Linear)), None)
        gap_fused = (self.eng.dtype == "bf16" and self.eng.impl != 1 and lin0 is not None and C % 8 == 0 and
                     lin0.out_features % 8 == 0)
        if not gap_fused:
            for ops in (self.fwd, self.fwd_eval):
                self._op(ops, "mnb_gap_fwd", ref.t, ref.scale, ref.shift, self.f, N, ref.H * ref.W, C, self.code)
        cur, width, pending = self.f, C, None
        self.head = []
        self.dropout_masks = []
        di = 0
        for i, ch in enumerate(children):
            if isinstance(ch, nn.Dropout):
                mask = torch.ones((N, width), device=self.dev, dtype=torch.uint8)
                self.keep.append(mask)
                self.dropout_masks.append((mask, float(ch.p)))
                pending = (mask, float(ch.p), di)
                di += 1
            elif isinstance(ch, nn.Linear):
                relu = i + 1 < len(children) and isinstance(children[i + 1], nn.ReLU)
                O = ch.out_features
                y = self._f32(N, O)
                rec = {"x": cur, "mask": None, "ms": 1.0, "lin": ch, "y": y, "relu": relu, "K": width, "O": O,
                       "tc": False}
                drop = pending is not None and pending[1] > 0
                if drop:
                    mask, p, idx = pending
                    rec.update(mask=mask, ms=1.0 / (1.0 - p), p=p, idx=idx)
                    self.fwd.append(_DropoutOp(self, mask, N * width, p, idx))
                maskp = _MaskPtr(self, rec["mask"]) if drop else None
                msc = _MaskScale(self, rec["ms"]) if drop else 1.0
                # bf16 mode: the classifier GEMMs run on the tensor pipe (tcgen05), fp32 accumulate and output
                if self.eng.dtype == "bf16" and self.eng.impl != 1 and width % 8 == 0 and O % 8 == 0:
                    xb = torch.empty((N, width), device=self.dev, dtype=torch.bfloat16)
                    dyb = torch.empty((N, O), device=self.dev, dtype=torch.bfloat16)
                    pf = torch.empty(O * width, device=self.dev, dtype=torch.bfloat16)
                    pd = torch.empty(O * width, device=self.dev, dtype=torch.bfloat16)
                    self.keep += [xb, dyb, pf, pd]
                    rec.update(tc=True, xb=xb, dyb=dyb, pf=pf, pd=pd)
                    for ops, train in ((self.fwd, True), (self.fwd_eval, False)):
                        self._op(ops, "mnb_pack_weights", ch.weight, pf, pd if train else None, O, width, 1,
                                 label="pack_weights")
                        if cur is self.f and gap_fused:
                            self._op(ops, "mnb_gap_fc_prep", ref.t, ref.scale, ref.shift, self.f,
                                     maskp if train else None, msc if train else 1.0, xb, N, ref.H * ref.W, C,
                                     self.code, label="gap+fc_prep")
                        else:
                            self._op(ops, "mnb_fc_prep_bf16", cur, maskp if train else None, msc if train else 1.0,
                                     xb, N * width)
                        self._op(ops, "mnb_fc_fwd_tc", xb, ch.weight, pf, ch.bias, y, int(relu), N, width, O,
                                 label="fc_fwd(tcgen05)")
                else:
                    self._op(self.fwd, "mnb_fc_fwd", cur, maskp, msc, ch.weight, ch.bias, y, int(relu), N, width, O)
                    self._op(self.fwd_eval, "mnb_fc_fwd", cur, None, 1.0, ch.weight, ch.bias, y, int(relu), N,
                             width, O)
                self.head.append(rec)
                cur, width, pending = y, O, None
            elif isinstance(ch, nn.ReLU):
                pass
            else:
                raise TypeError(f"classifier child {type(ch).__name__} not lowered")
        self.logits = cur
        self.num_out = width
        self.dlogits = self._f32(N, width)
        self.loss = self._f32(1)
        self.target = torch.zeros(N, device=self.dev, dtype=torch.long)

    # ---- backward program ----------------------------------------------------------------------------
    def _emit_backward(self):
        eng = self.eng
        maxel = 1
        for a in self.apps:
            maxel = max(maxel, a.inp.M * a.inp.C if not a.inp.nchw else 1, int(a.m) * a.Cout)
        # 4 buffers are live at most on the main chain; the extra ones give the weight-gradient kernels (side
        # stream, off the critical path) time to finish before their dZ operand is recycled
        pool = [torch.empty(maxel, device=self.dev, dtype=self.tdtype) for _ in range(4 + eng.wgrad_slack)]
        self.keep += pool
        free = list(pool)
        ops = self.bwd
        pending = {}                       # id(buffer) -> event id of the side-stream wgrad still reading it
        side = eng.wgrad_slack > 0

        def take():
            b = free.pop(0)                # FIFO: the buffer released longest ago
            ev = pending.pop(id(b), None)
            if ev is not None:
                ops.append(_EventOp("wait", ev, 0))
            return b
        net = getattr(eng.module, "_mnb", None) == "net"

        def rg(p):
            return p is not None and p.requires_grad

        def Gp(p):                         # gradient slot, or NULL for a frozen parameter (freeze(), classifiers.py:94)
            return _G(p) if rg(p) else None
        # frozen parameters get no gradient and backward stops below the first (in forward order) ConvBlock that
        # still has a trainable parameter -- what autograd does for requires_grad=False leaves (train.py:219)
        first_live = None
        for ti, entry in enumerate(self.tape):
            if entry[0] == "conv":
                cbk = entry[1].cb
                if rg(cbk.conv.weight) or rg(cbk.conv.bias) or rg(cbk.bn.weight) or rg(cbk.bn.bias):
                    first_live = ti
                    break
        sub_needs_dx = not net             # a stand-alone sub-module returns the input gradient
        if first_live is None and not sub_needs_dx:
            first_live = len(self.tape)    # nothing trainable in the features: backward ends at the head
        if sub_needs_dx:
            first_live = -1                # walk the whole tape and keep the first conv's data gradient
        self.bwd_stop = first_live
        if net:
            # head backward
            g = self.dlogits
            head_live = [rg(r["lin"].weight) or rg(r["lin"].bias) for r in self.head]
            for i in range(len(self.head) - 1, -1, -1):
                if first_live >= len(self.tape) and not any(head_live[:i + 1]):
                    break                  # nothing trainable below this layer
                r = self.head[i]
                lin = r["lin"]
                maskp = _MaskPtr(self, r["mask"]) if r["mask"] is not None else None
                ms = _MaskScale(self, r["ms"]) if r["mask"] is not None else 1.0
                dx = self._f32(self.N, r["K"])
                relu_ref = r["x"] if (i > 0 and self.head[i - 1]["relu"]) else None
                if r["tc"]:
                    self._op(ops, "mnb_fc_prep_bf16", g, None, 1.0, r["dyb"], self.N * r["O"])
                    if rg(lin.weight):
                        self._op(ops, "mnb_conv_wgrad", r["xb"], None, None, r["dyb"], _G(lin.weight), self.N, 1, 1,
                                 r["K"], r["O"], 1, 1, 0, _lib.MNB_BF16, _lib.LAYOUT_NHWC, 0,
                                 label="fc_wgrad(tcgen05)")
                    if rg(lin.bias):
                        self._op(ops, "mnb_fc_bias_grad", g, _G(lin.bias), self.N, r["O"])
                    self._op(ops, "mnb_fc_dgrad_tc", r["dyb"], lin.weight, r["pd"], dx, self.N, r["K"], r["O"],
                             label="fc_dgrad(tcgen05)")
                    self._op(ops, "mnb_fc_gate", dx, maskp, ms, relu_ref, self.N * r["K"])
                else:
                    if rg(lin.weight) or rg(lin.bias):
                        if not (rg(lin.weight) and rg(lin.bias)):
                            raise NotImplementedError("mnb200: freeze a Linear's weight and bias together")
                        self._op(ops, "mnb_fc_wgrad", r["x"], maskp, ms, g, _G(lin.weight), _G(lin.bias), self.N,
                                 r["K"], r["O"])
                    self._op(ops, "mnb_fc_dgrad", g, lin.weight, maskp, ms, relu_ref, dx, self.N, r["K"], r["O"])
                g = dx
            gbuf = None
            if first_live < len(self.tape):
                gbuf = take()
                hi = self.head_in
                self._op(ops, "mnb_gap_bwd", g, gbuf, self.N, hi.H * hi.W, hi.C, self.code)
        else:
            gbuf = take()
            self.sub_dout = gbuf          # filled from the NCHW grad_output at run time
        held = []                          # stack of residual-skip gradients
        for idx in range(len(self.tape) - 1, max(first_live, 0) - 1, -1):
            entry = self.tape[idx]
            tag = entry[0]
            if tag == "res_end":
                held.append(gbuf)          # dY of the block: feeds CB3 as dA and the skip add of CB1's dgrad
            elif tag == "res_begin":
                skip = held.pop()
                if skip is not gbuf:
                    free.append(skip)
            elif tag == "materialize":
                pass                       # dY == dA of the producing conv
            elif tag == "conv":
                a = entry[1]
                conv, bn = a.cb.conv, a.cb.bn
                M, C = int(a.m), a.Cout
                self._cur_detail = f"{a.inp.H}x{a.inp.W} {a.inp.C}->{a.Cout} k{a.k}s{a.stride}"
                if self._fused_dw_backward(a, idx, first_live):
                    # one kernel: BN-backward elementwise pass + backward-data + backward-weight of this depthwise
                    # block + the BN-backward reductions of the block that produced its input (csrc/dw_mma.cu)
                    r = a.inp
                    pa = self.apps[a.index - 1]
                    if not a.reduce_fused:
                        self._op(ops, "mnb_bn_bwd_reduce", gbuf, a.z, a.scale, a.shift, a.sums, M, C, self.code,
                                 nbytes=2 * a.out_bytes, label="bn_bwd_reduce")
                    dx = take()
                    self._op(ops, "mnb_dw_bwd_fused", gbuf, a.z, a.scale, a.shift, a.sums, a.mean, a.invstd,
                             Gp(bn.weight), Gp(bn.bias), Gp(conv.bias), r.t, r.scale, r.shift, conv.weight, dx,
                             Gp(conv.weight), pa.sums, r.N, r.H, r.W, C, a.k, a.m, self.code,
                             nbytes=4 * a.out_bytes, label=a.label + "_bwd_fused")
                    pa.reduce_fused = True
                    if not any(gbuf is h for h in held):
                        free.append(gbuf)
                    gbuf = dx
                    continue
                if self._fused_pw_backward(a, idx, first_live):
                    # one kernel: BN-backward elementwise pass + backward-data (+ skip gradient) + backward-weight of
                    # this 1x1 block + the BN-backward reductions of the block that produced its input
                    # (csrc/pw_bwd_fused.cu)
                    r = a.inp
                    pa = self.apps[a.index - 1] if (a.index > 0 and r.scale is not None and
                                                    self.apps[a.index - 1].z is r.t) else None
                    if not a.reduce_fused:
                        self._op(ops, "mnb_bn_bwd_reduce", gbuf, a.z, a.scale, a.shift, a.sums, M, C, self.code,
                                 nbytes=2 * a.out_bytes, label="bn_bwd_reduce")
                    dx = take()
                    add = held[-1] if (idx > 0 and self.tape[idx - 1][0] == "res_begin") else None
                    self._op(ops, "mnb_pw_bwd_fused", gbuf, a.z, a.scale, a.shift, a.sums, a.mean, a.invstd,
                             Gp(bn.weight), Gp(bn.bias), Gp(conv.bias), r.t, r.scale, r.shift, conv.weight, add, dx,
                             Gp(conv.weight), pa.sums if pa is not None else None, r.M, r.C, C, a.m, self.code,
                             nbytes=2 * a.out_bytes + a.in_bytes * (3 if add is not None else 2),
                             label=a.label + "_bwd_fused")
                    if pa is not None:
                        pa.reduce_fused = True
                    if not any(gbuf is h for h in held):
                        free.append(gbuf)
                    gbuf = dx
                    continue
                dz = take()
                if not a.reduce_fused:     # else: done in the epilogue of the dgrad that produced gbuf
                    self._op(ops, "mnb_bn_bwd_reduce", gbuf, a.z, a.scale, a.shift, a.sums, M, C, self.code,
                             nbytes=2 * a.out_bytes, label="bn_bwd_reduce")
                self._op(ops, "mnb_bn_bwd_apply_fused", gbuf, a.z, a.scale, a.shift, a.sums, a.mean, a.invstd,
                         Gp(bn.weight), Gp(bn.bias), Gp(conv.bias), dz, M, C, a.m, self.code,
                         nbytes=3 * a.out_bytes, label="bn_bwd_apply")
                in_held = any(gbuf is h for h in held)
                if not in_held:
                    free.append(gbuf)
                r = a.inp
                if self._proj_pw_backward(a, idx, first_live):
                    # project 1x1 block (wide -> narrow): backward-data (+ skip gradient) + backward-weight + the
                    # BN-backward reductions of the block that produced its input in one kernel (csrc/pw_proj_bwd.cu)
                    pa = self.apps[a.index - 1] if (a.index > 0 and r.scale is not None and
                                                    self.apps[a.index - 1].z is r.t) else None
                    dx = take()
                    add = held[-1] if (idx > 0 and self.tape[idx - 1][0] == "res_begin") else None
                    self._op(ops, "mnb_pw_proj_bwd", dz, r.t, r.scale, r.shift, conv.weight, add, dx, Gp(conv.weight),
                             pa.sums if pa is not None else None, r.M, r.C, C, self.code,
                             nbytes=a.out_bytes + a.in_bytes * (3 if add is not None else 2),
                             label=a.label + "_bwd_proj")
                    if pa is not None:
                        pa.reduce_fused = True
                    free.append(dz)
                    gbuf = dx
                    continue
                x_t = r.t if not r.nchw else _InputPtr(self)
                layout = a.in_layout
                wg = rg(conv.weight)
                if side and wg:            # fork: the side stream may start once dZ is complete
                    ev_dz = self._new_event()
                    ops.append(_EventOp("record", ev_dz, 0))
                    ops.append(_EventOp("wait", ev_dz, 1))
                if wg and a.kind == "dense":
                    self._op(ops, "mnb_conv_wgrad", x_t, a.in_sc, a.in_sh, dz, _G(conv.weight), r.N, r.H, r.W,
                             r.C, C, a.k, a.stride, a.pad, self.code, layout, eng.impl,
                             nbytes=a.in_bytes + a.out_bytes, label=a.label + "_wgrad")
                elif wg:
                    self._op(ops, "mnb_dw_wgrad", x_t, r.scale, r.shift, dz, _G(conv.weight), r.N, r.H, r.W, r.C,
                             a.k, self.code, nbytes=a.in_bytes + a.out_bytes, label=a.label + "_wgrad")
                if side and wg:
                    ops[-1].stream_id = 1
                    ev_w = self._new_event()
                    ops.append(_EventOp("record", ev_w, 1))
                    pending[id(dz)] = ev_w
                    self.side_last = ev_w
                if a.need_dgrad and idx > first_live:
                    dx = take()
                    # the first conv of a residual block adds the skip gradient (dY) into its dgrad output
                    add = None
                    if idx > 0 and self.tape[idx - 1][0] == "res_begin":
                        add = held[-1]
                    # dx is the dA of the ConvBlock applied just before this one: fuse its BN-backward
                    # reduction (sum G, sum G*z) into this dgrad's epilogue
                    pa = self.apps[a.index - 1] if (a.index > 0 and eng.fuse_bn_reduce) else None
                    bn = (pa.z, pa.scale, pa.shift, pa.sums) if pa is not None else (None, None, None, None)
                    if pa is not None:
                        pa.reduce_fused = True
                    extra = a.in_bytes if pa is not None else 0
                    if a.kind == "dense":
                        wpk_d = self.packed[id(a.cb)][1] if id(a.cb) in self.packed else None
                        self._op(ops, "mnb_conv_dgrad_packed", dz, conv.weight, wpk_d, add, dx, *bn, r.N, r.H, r.W, r.C, C, a.k,
                                 a.stride, a.pad, self.code, eng.impl,
                                 nbytes=a.in_bytes * (2 if add is not None else 1) + a.out_bytes + extra,
                                 label=a.label + "_dgrad")
                    else:
                        self._op(ops, "mnb_dw_dgrad", dz, conv.weight, dx, *bn, r.N, r.H, r.W, r.C, a.k, self.code,
                                 nbytes=a.in_bytes + a.out_bytes + extra, label=a.label + "_dgrad")
                    free.append(dz)
                    gbuf = dx
                else:
                    free.append(dz)
                    gbuf = None
        self.sub_din = gbuf
        if side and getattr(self, "side_last", None) is not None:
            ops.append(_EventOp("wait", self.side_last, 0))      # join: optimizer / caller see all weight grads

    def _fused_dw_backward(self, a, idx, first_live) -> bool:
        """Where the fused depthwise backward kernels are used (bf16; the input must come straight from a ConvBlock: raw z
        + scale / shift) -- the layers where they measured faster than the unfused chain: the row-streaming kernel
        (dw_mma.cu) for 3x3 on maps of >= 56 rows with a channel count the 24-channel geometry tiles exactly
        (profiles/r2_exp_dw_mma.json) when the whole-tile kernels are switched off; by default the whole-tile kernel
        (dw_small.cu) on every map of >= 12 rows (5x5 up to 64 rows): 242 vs 277, 148 vs 172, 76 vs 136 us on the 28 x 28 /
        14 x 14 blocks (scripts/exp_dw_small.py bwd), and 14.85 -> 14.25 ms per step from the 56 x 56 5x5 and 112 x 112 3x3
        blocks."""
        eng = self.eng
        if eng.dtype != "bf16" or not eng.fuse_dw_bwd or a.kind != "dw":
            return False
        r = a.inp
        if r.scale is None or a.index == 0 or self.apps[a.index - 1].z is not r.t:
            return False
        if not (a.need_dgrad and idx > first_live):
            return False
        if eng.fuse_dw_bwd == 2:                 # forced (tests): every shape the kernel supports
            return r.H >= 1
        if _lib.get_option("dw_small") == 1:      # whole-tile fused kernel (dw_small.cu): every map it is dispatched for
            return r.H >= 12 and r.W >= 12 and (a.k == 3 or r.H <= 64)
        return a.k == 3 and r.H >= 56 and r.W >= 24 and a.Cout % 24 == 0

    PW_FUSED_SHAPES = {(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)}       # (Cin, Cout) instantiated in pw_bwd_fused.cu

    def _fused_pw_backward(self, a, idx, first_live) -> bool:
        """Where the fused pointwise backward kernel is used: bf16 1x1 blocks of the shapes it is instantiated for (the
        expand / project blocks of the 112x112 and 56x56 stages), whose data gradient is needed."""
        eng = self.eng
        if eng.dtype != "bf16" or not eng.fuse_pw_bwd or a.kind != "dense" or a.k != 1 or a.stride != 1:
            return False
        if a.inp.nchw or not (a.need_dgrad and idx > first_live):
            return False
        return (a.inp.C, a.Cout) in self.PW_FUSED_SHAPES

    def _proj_pw_backward(self, a, idx, first_live) -> bool:
        """Where the fused project-block backward (after the BN pass) is used: bf16 1x1 blocks 240->40, 480->80, 576->96
        (the shapes csrc/pw_proj_bwd.cu is instantiated for) whose data gradient is needed."""
        eng = self.eng
        if eng.dtype != "bf16" or not eng.fuse_proj_bwd or a.kind != "dense" or a.k != 1 or a.stride != 1:
            return False
        if a.inp.nchw or not (a.need_dgrad and idx > first_live):
            return False
        cin, cout = a.inp.C, a.Cout
        return (cout in (96, 80) and cin % 96 == 0 and cin >= 4 * cout) or (cout == 40 and cin % 80 == 0 and cin >= 4 * cout)

    def _new_event(self):
        self.n_events += 1
        return self.n_events - 1

    # ---- op plumbing -----------------------------------------------------------------------------------
    def _op(self, ops, name, *args, nbytes=0, label=None):
        fn = getattr(lib, name)
        conv = []
        if ops is self.bwd:
            for a in args:
                if isinstance(a, _G):
                    self.last_write[id(a.p)] = len(ops) + 1     # gradient slot final after this op index
        for a in args:
            if isinstance(a, torch.Tensor):
                conv.append(_TensorPtr(a))
            elif isinstance(a, nn.Parameter):
                conv.append(_TensorPtr(a))
            else:
                conv.append(a)
        op = _Op(name, fn, conv, self)
        op.nbytes, op.label = nbytes, label or name
        op.stream_id = 0
        op.detail = getattr(self, "_cur_detail", "")
        ops.append(op)

    def run_graphed(self, ops, pre=None):
        """Runs a launch program (forward or backward of the module path) through a CUDA graph: the first call of a
        program runs eagerly (warm-up, lazy module loading), the second captures it, later calls replay.  Every operand
        of the program lives at a static address (plan buffers, the engine's flat parameter / gradient stores, the
        static input copy), device-side counters carry the dropout seed, so a replay is the same work as the eager loop
        minus ~200 Python / ctypes launches -- at 14x14 / 7x7 the kernels are shorter than a launch and the GPU starved.
        Graphs are keyed by the program list, the dropout state and the kernel-option epoch (a rebuilt program or a
        changed option recaptures).  `pre` = in-place zeroing that belongs to the program."""
        key = (id(ops), pre is not None, getattr(self, "dropout_active", False), getattr(self, "masks_injected", False),
               _lib.OPTION_EPOCH)
        ent = self.prog_graphs.get(key)
        stream = torch.cuda.current_stream().cuda_stream
        if ent is None:                       # first call: eager
            if len(self.prog_graphs) > 16:
                self.prog_graphs.clear()
            self.prog_graphs[key] = ("warm", ops)       # holds the list: its id cannot be recycled by a rebuilt program
            if pre is not None:
                pre()
            self.run(ops, stream)
            return
        if ent[0] == "warm":                  # second call: capture, then replay
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if pre is not None:
                    pre()
                self.run(ops, torch.cuda.current_stream().cuda_stream)
            self.prog_graphs[key] = ent = (g, ops)      # holds the list: its id cannot be recycled while the graph lives
        ent[0].replay()

    def run(self, ops, stream, after_op=None):
        prof = self.eng.profile
        if prof is None:
            if self.n_events and self.events is None:
                self.events = [torch.cuda.Event() for _ in range(self.n_events)]
            main = torch.cuda.current_stream()
            side = self.eng.side_stream
            side_ptr = side.cuda_stream
            for i, op in enumerate(ops):
                if op.__class__ is _EventOp:
                    st = main if op.stream_id == 0 else side
                    if op.kind == "record":
                        self.events[op.ev].record(st)
                    else:
                        st.wait_event(self.events[op.ev])
                else:
                    op(stream if getattr(op, "stream_id", 0) == 0 else side_ptr)
                if after_op is not None:
                    after_op(i)
            return
        cur = torch.cuda.current_stream()
        for op in ops:          # per-launch CUDA-event timing on the launching stream (bench.py roofline pass)
            if op.__class__ is _EventOp:
                continue        # profiling pass is single-stream: no cross-stream dependencies needed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            op(stream)
            e1.record(cur)
            prof.append((getattr(op, "label", op.name), getattr(op, "nbytes", 0), e0, e1,
                         getattr(op, "detail", "")))


class _TensorPtr:
    """Pointer of a tensor resolved at bind time (plan buffers, and parameters/buffers which are views into
    the engine's flat stores -- validity is checked once per call by ParamStore.valid())."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t

    def resolve(self, plan):
        return self.t.data_ptr()


class _G:
    """Pointer of the gradient slot of a parameter in the flat gradient buffer."""
    __slots__ = ("p",)

    def __init__(self, p):
        self.p = p

    def resolve(self, plan):
        return plan.eng.store.gptr(self.p)


class _Lazy:
    __slots__ = ("plan", "off")

    def __init__(self, plan, off):
        self.plan, self.off = plan, off

    def resolve(self, plan):
        return plan._dstat_base + self.off


class _InputPtr:
    """The NCHW fp32 network input: bound per call."""
    __slots__ = ("plan",)

    def __init__(self, plan):
        self.plan = plan

    def resolve(self, plan):
        return None        # dynamic


class _MaskPtr:
    """Dropout mask pointer, NULL when dropout is disabled for this call."""
    __slots__ = ("plan", "mask")

    def __init__(self, plan, mask):
        self.plan, self.mask = plan, mask

    def resolve(self, plan):
        return None        # dynamic


class _MaskScale:
    __slots__ = ("plan", "ms")

    def __init__(self, plan, ms):
        self.plan, self.ms = plan, ms

    def resolve(self, plan):
        return None        # dynamic


class _Op:
    """A bound kernel launch.  Static arguments are resolved once (first run after (re)binding); dynamic
    ones (network input pointer, dropout on/off) are patched per call."""

    def __init__(self, name, fn, args, plan):
        self.name, self.fn, self.args, self.plan = name, fn, args, plan
        self.bound = None
        self.dyn = [(i, a) for i, a in enumerate(args) if isinstance(a, (_InputPtr, _MaskPtr, _MaskScale))]

    def bind(self):
        plan = self.plan
        self.bound = [a.resolve(plan) if hasattr(a, "resolve") else a for a in self.args]

    def __call__(self, stream):
        if self.bound is None:
            self.bind()
        b = self.bound
        plan = self.plan
        for i, a in self.dyn:
            if isinstance(a, _InputPtr):
                b[i] = plan.cur_input_ptr
            elif isinstance(a, _MaskPtr):
                b[i] = a.mask.data_ptr() if plan.dropout_active else None
            else:
                b[i] = a.ms if plan.dropout_active else 1.0
        rc = self.fn(*b, stream)
        if rc != 0:
            check(rc, self.name)


class _EventOp:
    """Cross-stream dependency inside a program: record / wait of plan event `ev` on stream 0 (main) or 1 (the
    weight-gradient side stream)."""
    __slots__ = ("kind", "ev", "stream_id", "name")

    def __init__(self, kind, ev, stream_id):
        self.kind, self.ev, self.stream_id = kind, ev, stream_id
        self.name = "event_" + kind


class _DropoutOp:
    """Generates a dropout keep-mask unless dropout is off or masks were injected for this call."""

    def __init__(self, plan, mask, n, p, idx):
        self.plan, self.mask, self.n, self.p, self.idx = plan, mask, n, p, idx
        self.name = "mnb_dropout_mask"

    def __call__(self, stream):
        plan = self.plan
        if not plan.dropout_active or plan.masks_injected:
            return
        eng = plan.eng
        # counter-based RNG keyed by (seed, per-forward counter, layer): a fresh mask per training forward
        # (nn.Dropout draws one per call, classifiers.py:82,85), independent of the optimizer step
        check(lib.mnb_dropout_mask(self.mask.data_ptr(), self.n, self.p, eng.seed,
                                   (self.idx + 1) << 40, eng.dev_fwd.data_ptr(), stream), self.name)


class Engine:
    """Owns the flat parameter store and the per-shape plans of one lowered module."""

    def __init__(self, module: nn.Module, dtype: str = "bf16", impl: str = "auto", device=None, seed=None):
        _require_cuda()
        if dtype not in DTYPES:
            raise ValueError(f"dtype must be one of {list(DTYPES)}")
        self.module = module
        self.dtype, self.impl = dtype, IMPLS[impl]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.in_channels = _first_conv(module).in_channels
        self.store = ParamStore(module, self.device)
        self.plans = {}
        # dropout stream: torch's seed (torch.manual_seed, src/train.py:110) folded with the rank, so replicas draw
        # different masks like DataParallel's per-replica nn.Dropout
        self.seed = _default_seed() if seed is None else int(seed)
        self.dev_step = torch.zeros(1, device=self.device, dtype=torch.long)   # optimizer step counter (device truth)
        self.dev_fwd = torch.zeros(1, device=self.device, dtype=torch.long)    # training forwards run (dropout RNG)
        self.dev_lr = torch.zeros(1, device=self.device, dtype=torch.float32)
        self.dropout = "on"            # "on" | "off"
        self.grad_hook = None          # DDP: callable(engine, stage) invoked while backward is enqueued
        self.graphs = {}
        self.world_size = 1
        self.wgrad_slack = 4           # extra scratch buffers = how far weight-gradient kernels may trail
        self.side_stream = torch.cuda.Stream(device=self.device)
        # module path (model(x) / loss.backward()): replay the forward / backward programs as CUDA graphs (MNB_MODULE_GRAPHS=0: eager)
        self.module_graphs = os.environ.get("MNB_MODULE_GRAPHS", "1") != "0"
        self.fuse_bn_reduce = False    # BN-backward reductions in the producing dgrad epilogue (tested; off:
                                       # the dgrad epilogues are the bottleneck, the separate kernel is faster)
        # fused depthwise ConvBlock backward (csrc/dw_mma.cu): 0 off, 1 where it wins, 2 always (MNB_FUSE_DW_BWD overrides)
        self.fuse_dw_bwd = int(os.environ.get("MNB_FUSE_DW_BWD", "1"))
        self.fuse_pw_bwd = 1           # fused pointwise ConvBlock backward (csrc/pw_bwd_fused.cu) for its shapes
        self.fuse_proj_bwd = int(os.environ.get("MNB_FUSE_PROJ_BWD", "1"))   # project blocks: dgrad + wgrad + producer reduction (csrc/pw_proj_bwd.cu)
        self.profile = None            # list -> Plan.run records (label, bytes, ev0, ev1) per launch
        self.optimizer = "adam"        # 'adam' | 'rmsprop' | 'sgd' (train.py:218-231); set before the first graph capture

    @property
    def host_step(self) -> int:
        """Optimizer steps taken.  The device counter is the only copy (graph replays advance it without the host
        seeing them); reading it synchronises, which only checkpointing does."""
        return int(self.dev_step.item())

    @host_step.setter
    def host_step(self, value: int):
        self.dev_step.fill_(int(value))

    def plan(self, N, H, W, input_u8: bool = False) -> Plan:
        key = (N, H, W, self.store.trainable_sig(), bool(input_u8))
        p = self.plans.get(key)
        if p is None:
            p = Plan(self, N, H, W, bool(input_u8))
            self.plans[key] = p
        return p

    def norm_constants(self, C=3):
        """Normalize(mean, std) of the input pipeline (module.mean / .std, classifiers.py:91-92) as device vectors."""
        if getattr(self, "_norm", None) is None:
            mean = getattr(self.module, "mean", (0.0,) * C)
            std = getattr(self.module, "std", (1.0,) * C)
            self._norm = (torch.tensor(mean, device=self.device, dtype=torch.float32),
                          torch.tensor(std, device=self.device, dtype=torch.float32), {})
        mean, std, _ = self._norm
        if mean.numel() != C:
            raise ValueError(f"{C} input channels, but module.mean has {mean.numel()}")
        return mean, std

    def _check_store(self):
        if not self.store.valid():
            # parameters were re-allocated (e.g. model.to()/load on another device): rebuild everything
            self.store = ParamStore(self.module, self.device)
            self.plans.clear()
            self.graphs.clear()
        return self.store

    # ---- forward / backward on the current stream ------------------------------------------------
    def forward(self, x: torch.Tensor, train: bool, dropout_masks=None) -> Plan:
        self._check_store()
        if x.dim() != 4:
            raise ValueError("expected N x C x H x W input")
        if not x.is_cuda:
            raise RuntimeError("mnb200: input must be a CUDA tensor (no CPU fallback)")
        u8 = x.dtype == torch.uint8
        if u8:
            # N x H x W x 3 uint8 as decoded: ToTensor + Normalize are fused into the stem (whole-network engines only)
            if getattr(self.module, "_mnb", None) != "net" or x.shape[-1] != self.in_channels:
                raise ValueError("uint8 input: expected N x H x W x 3 on a FineTuneModelPool engine")
            x = x.contiguous()
            N, H, W, C = x.shape
        else:
            x = x.contiguous().float()
            N, C, H, W = x.shape
        plan = self.plan(N, H, W, u8)
        stream = torch.cuda.current_stream().cuda_stream
        plan.cur_input = x
        plan.cur_input_ptr = x.data_ptr()
        drop_on = self.dropout == "on" and any(isinstance(m, nn.Dropout) and m.training
                                               for m in self.module.modules())
        plan.dropout_active = bool(train) and (drop_on or dropout_masks is not None)
        plan.masks_injected = dropout_masks is not None
        if dropout_masks is not None:
            for (mask, _), src in zip(plan.dropout_masks, dropout_masks):
                mask.copy_(src.to(device=self.device, dtype=torch.uint8))
        sub = getattr(self.module, "_mnb", None) != "net"
        if sub:
            check(lib.mnb_nchw_f32_to_nhwc(x.data_ptr(), plan.sub_in.data_ptr(), N, H, W, C, plan.code, stream),
                  "nchw_to_nhwc")
        if train and self._graph_ok(forward=True) and not sub and dropout_masks is None:
            if plan.static_in is None or plan.static_in.shape != x.shape or plan.static_in.dtype != x.dtype:
                plan.static_in = torch.empty_like(x)
            plan.static_in.copy_(x)             # the program reads its input at a static address
            plan.cur_input = plan.static_in
            plan.cur_input_ptr = plan.static_in.data_ptr()
            plan.run_graphed(plan.fwd, plan.dstats.zero_)
        elif train:
            plan.dstats.zero_()
            plan.run(plan.fwd, stream)
        else:
            plan.run(plan.fwd_eval, stream)
        if sub:
            o = plan.sub_out
            check(lib.mnb_nhwc_to_nchw_f32(o.t.data_ptr(), plan.out_nchw.data_ptr(), o.N, o.H, o.W, o.C, plan.code,
                                           stream), "nhwc_to_nchw")
        plan.generation += 1
        return plan

    def backward(self, plan: Plan, zero_grads: bool = True):
        """Runs the backward program; dlogits (or sub_dout) must already be in place."""
        stream = torch.cuda.current_stream().cuda_stream
        hook = self.grad_hook
        if hook is None and self._graph_ok() and getattr(self.module, "_mnb", None) == "net" and not plan.masks_injected:
            plan.run_graphed(plan.bwd, self.store.grad.zero_ if zero_grads else None)
            return
        if zero_grads:
            self.store.grad.zero_()
        if hook is None:
            plan.run(plan.bwd, stream)
        else:
            hook.run_backward(self, plan, stream)

    def _graph_ok(self, forward=False) -> bool:
        """Module-path programs replay as CUDA graphs unless profiling or already inside a capture (the fused-step
        graph); under data parallel only the forward does (the backward's bucket hooks interleave NCCL calls with the
        launches)."""
        return (self.module_graphs and self.profile is None and (forward or self.grad_hook is None)
                and not torch.cuda.is_current_stream_capturing())

    # ---- fused training step (train.py:433-440 entirely in libmnb200) -----------------------------
    def train_step(self, x, target, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """forward + CrossEntropyLoss(mean) + backward (+ bucketed all-reduce when a mnb200.ddp.GradSync is
        attached) + Adam.  Returns the device loss tensor of shape [1] (this rank's mean loss)."""
        plan = self.forward(x, True)
        stream = torch.cuda.current_stream().cuda_stream
        plan.target.copy_(target, non_blocking=True)
        plan.loss.zero_()
        check(lib.mnb_xent_fwd_bwd(plan.logits.data_ptr(), plan.target.data_ptr(), plan.loss.data_ptr(),
                                   plan.dlogits.data_ptr(), plan.N, plan.num_out, 1.0, stream), "xent")
        self.backward(plan)
        gscale = 1.0
        if self.grad_hook is not None:
            self.grad_hook.wait()
            gscale = 1.0 / self.grad_hook.world
        self.optimizer_step(lr, betas, eps, gscale)
        return plan.loss

    def normalize_u8(self, x_u8: torch.Tensor) -> torch.Tensor:
        """uint8 N x H x W x C image batch (host or device) -> the normalised fp32 N x C x H x W device tensor the step
        consumes: ToTensor + Normalize(module.mean, module.std) on the device (utils/datasets.py:460-462), so the
        host-to-device copy carries one byte per value.  The result lives in a per-shape buffer."""
        if x_u8.dtype != torch.uint8 or x_u8.dim() != 4:
            raise ValueError("expected a uint8 N x H x W x C tensor")
        x_u8 = x_u8.to(self.device, non_blocking=True).contiguous()
        N, H, W, C = x_u8.shape
        mean, std = self.norm_constants(C)
        bufs = self._norm[2]
        y = bufs.get((N, H, W))
        if y is None:
            y = bufs[(N, H, W)] = torch.empty((N, C, H, W), device=self.device, dtype=torch.float32)
        check(lib.mnb_u8hwc_to_nchw_f32(x_u8.data_ptr(), mean.data_ptr(), std.data_ptr(), y.data_ptr(), N, H, W, C,
                                        torch.cuda.current_stream().cuda_stream), "u8hwc_to_nchw")
        return y

    def optimizer_step(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, lr_on_device=False):
        """`optimizer.step()` of the optimizer selected by `Engine.optimizer` (train.py:218-231: 'adam' | 'rmsprop' |
        'sgd', all with torch's defaults) over the flat parameter / gradient buffers."""
        if self.optimizer == "adam":
            return self.adam(lr, betas, eps, grad_scale, lr_on_device)
        stream = torch.cuda.current_stream().cuda_stream
        st = self.store
        dev_lr = self.dev_lr.data_ptr() if lr_on_device else None
        if self.optimizer not in ("sgd", "rmsprop"):
            raise ValueError(f"unknown optimizer {self.optimizer!r}")
        check(lib.mnb_counter_inc(self.dev_step.data_ptr(), stream), "counter_inc")
        pp, gp = st.flat.data_ptr(), st.grad.data_ptr()
        for lo, hi in st.trainable_ranges():    # only parameters with requires_grad are stepped (train.py:219)
            if self.optimizer == "sgd":
                check(lib.mnb_sgd_step(pp + 4 * lo, gp + 4 * lo, hi - lo, float(lr), grad_scale, dev_lr, stream), "sgd")
            else:
                _, v = st.adam_state()          # the second-moment buffer doubles as RMSprop's square_avg
                check(lib.mnb_rmsprop_step(pp + 4 * lo, gp + 4 * lo, v.data_ptr() + 4 * lo, hi - lo, float(lr),
                                           0.99, 1e-8, grad_scale, dev_lr, stream), "rmsprop")

    def adam(self, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, lr_on_device=False):
        stream = torch.cuda.current_stream().cuda_stream
        m, v = self.store.adam_state()
        check(lib.mnb_counter_inc(self.dev_step.data_ptr(), stream), "counter_inc")
        st = self.store
        for lo, hi in st.trainable_ranges():    # only parameters with requires_grad are stepped (train.py:219)
            check(lib.mnb_adam_step(st.flat.data_ptr() + 4 * lo, st.grad.data_ptr() + 4 * lo, m.data_ptr() + 4 * lo,
                                    v.data_ptr() + 4 * lo, hi - lo, float(lr), betas[0], betas[1], eps, 1, grad_scale,
                                    self.dev_lr.data_ptr() if lr_on_device else None,
                                    self.dev_step.data_ptr(), stream), "adam")

    # ---- CUDA-graph replay of the whole step (single GPU): ~450 launches -> one graph launch ---------
    def train_step_graph(self, x, target, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """Same as train_step, replayed from a CUDA graph captured per input shape.  x / target are copied
        into static buffers; lr and the step counter are read from device scalars so the schedule can change
        between replays (train.py:282-302,334)."""
        graph, sx, st, plan = self.capture_step_graph(x, target, lr, betas, eps)
        sx.copy_(x, non_blocking=True)
        st.copy_(target, non_blocking=True)
        self.dev_lr.fill_(float(lr))
        graph.replay()
        return plan.loss

    def capture_step_graph(self, x, target, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        """Captures (once per input shape) the CUDA graph of the fused step without advancing the training state:
        the eager warm-up step it needs is rolled back (parameters, BN buffers, optimizer moments, counters)."""
        if self.grad_hook is not None:
            raise RuntimeError("graph replay is single-GPU; use train_step with GradSync for data parallel")
        self._check_store()
        u8 = x.dtype == torch.uint8                     # N x H x W x 3 uint8: normalised inside the stem
        if u8:
            N, H, W, C = x.shape
        else:
            N, C, H, W = x.shape
        key = (N, H, W, self.store.trainable_sig(), self.optimizer, self.dropout, u8)
        g = self.graphs.get(key)
        if g is None:
            plan = self.plan(N, H, W, u8)
            self.store.adam_state()
            sx = torch.empty(tuple(x.shape), device=self.device, dtype=torch.uint8 if u8 else torch.float32)
            st = torch.zeros(N, device=self.device, dtype=torch.long)
            sx.copy_(x)
            st.copy_(target)
            self.dev_lr.fill_(float(lr))
            # one eager step on a side stream warms every kernel (and lazy module loading) before capture
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            saved = (self.store.flat.clone(), self.store.fbuf.clone(), self.store.ibuf.clone(),
                     self.dev_step.clone(), self.dev_fwd.clone(), self.store.m.clone(), self.store.v.clone())
            with torch.cuda.stream(s):
                self._graph_body(sx, st, betas, eps)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.store.flat.copy_(saved[0]); self.store.fbuf.copy_(saved[1]); self.store.ibuf.copy_(saved[2])
            self.dev_step.copy_(saved[3]); self.dev_fwd.copy_(saved[4])
            self.store.m.copy_(saved[5]); self.store.v.copy_(saved[6])     # (a resumed optimizer state survives)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._graph_body(sx, st, betas, eps)
            # capture does not execute: state is untouched
            g = (graph, sx, st, plan)
            self.graphs[key] = g
        return g

    def _graph_body(self, sx, st, betas, eps):
        plan = self.forward(sx, True)
        stream = torch.cuda.current_stream().cuda_stream
        plan.target.copy_(st)
        plan.loss.zero_()
        check(lib.mnb_xent_fwd_bwd(plan.logits.data_ptr(), plan.target.data_ptr(), plan.loss.data_ptr(),
                                   plan.dlogits.data_ptr(), plan.N, plan.num_out, 1.0, stream), "xent")
        self.backward(plan)
        self.optimizer_step(0.0, betas, eps, 1.0, lr_on_device=True)

    def launches_per_step(self, N, H, W):
        """Number of libmnb200 kernel launches in one fused training step of this shape."""
        plan = self.plan(N, H, W)
        n = 0
        for op in plan.fwd + plan.bwd:
            if op.__class__ is _EventOp:
                continue
            n += 2 if getattr(op, "name", "") == "mnb_fc_wgrad" else 1
        return n + 3          # xent, counter_inc, adam

    def logits(self, plan):
        return plan.logits


def _default_seed() -> int:
    """torch's global seed (torch.manual_seed(args.seed), src/train.py:110) mixed with the data-parallel rank."""
    import os
    rank = 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank = dist.get_rank()
        else:
            rank = int(os.environ.get("RANK", "0"))
    except Exception:
        rank = 0
    return (int(torch.initial_seed()) ^ (0x9E3779B97F4A7C15 * (rank + 1))) & 0x7FFFFFFFFFFFFFFF


def _first_conv(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            return m
    raise ValueError("module has no convolution to lower")


# -------------------------------------------------------------------------------------------------------
# autograd bridge: model(input) -> logits with loss.backward() filling param.grad (train.py:433-440)
# -------------------------------------------------------------------------------------------------------
class _ProgramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, dropout_masks, *params):
        plan = eng.forward(x, True, dropout_masks)
        ctx.eng, ctx.plan, ctx.gen = eng, plan, plan.generation
        ctx.x_needs_grad = x.requires_grad
        if getattr(eng.module, "_mnb", None) == "net":
            return plan.logits.clone()
        return plan.out_nchw.clone()

    @staticmethod
    def backward(ctx, dout):
        eng, plan = ctx.eng, ctx.plan
        if plan.generation != ctx.gen:
            raise RuntimeError("mnb200: the activations of this forward were overwritten by a later forward of "
                               "the same shape; call backward before the next forward")
        stream = torch.cuda.current_stream().cuda_stream
        dout = dout.contiguous().float()
        dx = None
        if getattr(eng.module, "_mnb", None) == "net":
            plan.dlogits.copy_(dout)
        else:
            o = plan.sub_out
            check(lib.mnb_nchw_f32_to_nhwc(dout.data_ptr(), plan.sub_dout.data_ptr(), o.N, o.H, o.W, o.C, plan.code,
                                           stream), "nchw_to_nhwc(grad)")
        eng.backward(plan)
        if getattr(eng.module, "_mnb", None) != "net" and ctx.x_needs_grad and plan.sub_din is not None:
            dx = torch.empty_like(plan.cur_input)
            N, C, H, W = dx.shape
            check(lib.mnb_nhwc_to_nchw_f32(plan.sub_din.data_ptr(), dx.data_ptr(), N, H, W, C, plan.code, stream),
                  "nhwc_to_nchw(grad)")
        if eng.grad_hook is not None:              # data parallel: buckets were all-reduced during backward
            eng.grad_hook.wait()
            g = eng.store.grad / float(eng.grad_hook.world)
        else:
            g = eng.store.grad.clone()             # one copy; .grad tensors are views of it
        views = eng.store.grad_views(g)
        grads = tuple(views[id(p)] if p.requires_grad else None for p in eng.store.params)
        return (None, dx, None) + grads


_ENGINES = weakref.WeakKeyDictionary()
_DEFAULTS = {"dtype": "bf16", "impl": "auto"}


def configure(module: Optional[nn.Module] = None, dtype: Optional[str] = None, impl: Optional[str] = None):
    """Select the activation dtype ("bf16" | "fp32") / GEMM implementation for a module (or the default for
    modules lowered later).  Must be called before the first forward of that module."""
    tgt = _DEFAULTS if module is None else module.__dict__.setdefault("_mnb_cfg", dict(_DEFAULTS))
    if dtype is not None:
        tgt["dtype"] = dtype
    if impl is not None:
        tgt["impl"] = impl
    if module is not None and module in _ENGINES:
        del _ENGINES[module]


def release(module: nn.Module):
    """Drops the engine of a module (plans, activation buffers, CUDA graphs).  The engine references the module, so
    the weak registry alone never frees it; long-running scripts that build many models call this."""
    eng = _ENGINES.pop(module, None)
    if eng is not None:
        eng.plans.clear()
        eng.graphs.clear()


def engine_for(module: nn.Module) -> Engine:
    eng = _ENGINES.get(module)
    if eng is None:
        cfg = module.__dict__.get("_mnb_cfg", _DEFAULTS)
        p = next(module.parameters())
        if not p.is_cuda:
            raise RuntimeError("mnb200: move the model to a CUDA device first (model.to('cuda')); "
                               "there is no CPU fallback")
        eng = Engine(module, cfg["dtype"], cfg["impl"], device=p.device)
        _ENGINES[module] = eng
    return eng


def run_module(module: nn.Module, x: torch.Tensor, dropout_masks=None) -> torch.Tensor:
    """``module(x)`` for any lowered module (FineTuneModelPool, Mnasnet, MBConv, MBConv_block, SepConv,
    ConvBlock): NCHW fp32 in, NCHW fp32 (or logits) out."""
    _require_cuda()
    eng = engine_for(module)
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in eng.store.params))
    if module.training and needs_grad:
        eng._check_store()
        return _ProgramFn.apply(eng, x, dropout_masks, *eng.store.params)
    if needs_grad and not module.training:
        raise RuntimeError("mnb200: gradients through eval-mode (running-stat) BatchNorm are not lowered")
    with torch.no_grad():
        plan = eng.forward(x, module.training, dropout_masks)
        if getattr(module, "_mnb", None) == "net":
            return plan.logits.clone()
        return plan.out_nchw.clone()
