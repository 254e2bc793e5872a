"""One process per GPU + NCCL bucketed gradient all-reduce overlapped with backward.

Replaces ``torch.nn.DataParallel(model)`` (src/train.py:202; torch:nn/parallel/data_parallel.py:173-198):
no per-step weight broadcast, no input scatter / logits gather -- every rank owns a replica and a batch shard,
BatchNorm statistics stay per rank (DataParallel has no SyncBN either), and the only exchange is one
all-reduce(sum) of the flat fp32 gradient buffer, issued per bucket on a communication stream as soon as the
last kernel writing into that bucket has been enqueued; the result is scaled by 1/world inside Adam.
DataParallel computes mean-CE over the gathered batch and sums replica gradients, which equals the average
of per-rank mean-CE gradients for equal shards.

Works with any torch.distributed backend: "nccl" on the B200 box, "gloo" in the CPU unit tests (which drive
``bucket_schedule`` / ``allreduce_buckets`` on host tensors).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


def bucket_schedule(buckets, offsets_last_write):
    """buckets: [(lo, hi)] slices of the flat gradient buffer in backward-ready order.
    offsets_last_write: [(offset, last_op_index)] for every parameter.
    Returns [(op_index, lo, hi)]: bucket (lo,hi) is final once op `op_index` (1-based) has been enqueued."""
    out = []
    for lo, hi in buckets:
        idx = max((w for off, w in offsets_last_write if lo <= off < hi), default=0)
        out.append((idx, lo, hi))
    # a later bucket can never be ready before an earlier one is enqueued on the same stream: keep order
    fixed, run = [], 0
    for idx, lo, hi in out:
        run = max(run, idx)
        fixed.append((run, lo, hi))
    return fixed


def allreduce_buckets(flat, buckets, group=None, async_op=False):
    """Sum `flat` across ranks bucket by bucket (used directly by the CPU tests)."""
    works = []
    for lo, hi in buckets:
        w = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


class GradSync:
    """Engine.grad_hook implementation: runs the backward program and overlaps the bucket all-reduces."""

    def __init__(self, engine, group=None, overlap=True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.engine, self.group, self.overlap = engine, group, overlap
        self.world = dist.get_world_size(group)
        self.comm_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._sched = {}
        self.works = []
        engine.grad_hook = self
        engine.world_size = self.world

    def schedule(self, plan):
        key = id(plan)
        if key not in self._sched:
            st = self.engine.store
            olw = [(st.offsets[id(p)][0], plan.last_write.get(id(p), len(plan.bwd))) for p in st.params]
            self._sched[key] = bucket_schedule(st.buckets(), olw)
        return self._sched[key]

    def run_backward(self, engine, plan, stream):
        sched = self.schedule(plan)
        grad = engine.store.grad
        cur = torch.cuda.current_stream()
        self.works = []
        state = {"si": 0}

        def after(i):
            while state["si"] < len(sched) and sched[state["si"]][0] <= i + 1:
                self._launch(cur, grad, sched[state["si"]][1], sched[state["si"]][2], plan, engine)
                state["si"] += 1
        plan.run(plan.bwd, stream, after_op=after)
        while state["si"] < len(sched):
            self._launch(cur, grad, sched[state["si"]][1], sched[state["si"]][2], plan, engine)
            state["si"] += 1

    def _launch(self, cur, grad, lo, hi, plan=None, engine=None):
        if self.world == 1:
            return
        if self.overlap:
            ev = torch.cuda.Event()
            ev.record(cur)
            ev2 = None
            if engine is not None and getattr(engine, "wgrad_slack", 0) > 0:
                ev2 = torch.cuda.Event()              # weight gradients are written on the engine's side stream
                ev2.record(engine.side_stream)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if ev2 is not None:
                    self.comm_stream.wait_event(ev2)
                self.works.append(dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                  async_op=True))
        else:
            dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def wait(self):
        """Make the current stream wait for every outstanding bucket."""
        for w in self.works:
            w.wait()
        self.works = []


def broadcast_parameters(engine, src=0, group=None):
    """Start-of-training sync (what DataParallel's replicate does every step, done once here)."""
    dist.broadcast(engine.store.flat, src=src, group=group)
    dist.broadcast(engine.store.fbuf, src=src, group=group)
    dist.broadcast(engine.store.ibuf, src=src, group=group)


class DataParallel(nn.Module):
    """``model = torch.nn.DataParallel(model)`` (src/train.py:202) for the process-per-GPU layout: the wrapper the
    training script keeps calling -- ``model(input)``, ``model.module.parameters()`` (:210-212), ``model.state_dict()``
    with ``module.``-prefixed keys (:383), ``load_state_dict`` (:242), ``train()/eval()`` -- while every process owns
    one replica and one batch shard.  With torch.distributed initialised (torchrun) the replicas are synchronised once
    from rank 0 and a GradSync is attached to the module's engine, so ``loss.backward()`` all-reduces (averages) the
    gradients bucket by bucket while backward runs; without it (single GPU) this is a transparent wrapper.
    ``device_ids`` / ``output_device`` / ``dim`` are accepted for signature compatibility and ignored."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, dim=0, group=None, overlap=True):
        super().__init__()
        from .engine import engine_for
        self.module = module
        self.device_ids, self.output_device, self.dim = device_ids, output_device, dim
        self.sync = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            eng = engine_for(module)
            broadcast_parameters(eng, 0, group)
            self.sync = GradSync(eng, group, overlap)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)
