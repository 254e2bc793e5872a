"""CPU (gloo, world_size 2): host-side logic of the data-parallel path -- bucket schedule over the flat
gradient buffer, bucketed all-reduce + averaging equals DataParallel's summed-gradient semantics, and the
oracle-level equivalence "per-rank mean loss + grad average == global-batch mean loss" (SURVEY.md section 5)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "mnasnet-pytorch_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mnb200 import ddp
    from oracle import mnasnet_oracle as O
    torch.set_num_threads(2)
    # each rank: oracle gradients on its own shard (BN statistics per rank, like DataParallel replicas)
    torch.manual_seed(42)
    sd = O.init_state_dict(10, '320')
    tr = O.Trainer(sd, classifier_config='320', num_classes=10)
    x, t = O.synthetic_batch(4, 64, 64, 10)
    shard = slice(rank * 2, rank * 2 + 2)
    _, loss, g = tr.grads(x[shard], t[shard], dropout_masks="off")
    names = tr.names[::-1]                      # backward-ready order, as ParamStore lays the flat buffer out
    sizes = [g[n].numel() for n in names]
    flat = torch.cat([g[n].reshape(-1) for n in names])
    offs = [0]
    for s_ in sizes:
        offs.append(offs[-1] + s_)
    # three buckets cut at stage boundaries; last-writer op indices increase along the buffer
    cut1 = next(o for o, n in zip(offs, names) if n.startswith("features.6"))
    cut2 = next(o for o, n in zip(offs, names) if n.startswith("features.4"))
    buckets = [(0, cut1), (cut1, cut2), (cut2, offs[-1])]
    olw = [(o, 10 * (i + 1)) for i, o in enumerate(offs[:-1])]
    sched = ddp.bucket_schedule(buckets, olw)
    assert [b[1:] for b in sched] == buckets
    assert sched[0][0] <= sched[1][0] <= sched[2][0] == 10 * len(names)
    works = ddp.allreduce_buckets(flat, buckets, async_op=True)
    for w in works:
        w.wait()
    flat /= world
    q.put((rank, loss.item(), flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_matches_per_shard_oracle():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0][2], res[1][2])            # every rank ends with the same averaged gradient
    # reference: both shards on one process, gradients averaged by hand
    from oracle import mnasnet_oracle as O
    gs = []
    nthreads = torch.get_num_threads()
    torch.set_num_threads(2)                # same CPU kernel path as the workers (gradients are thread-count
    for r in range(world):                  # sensitive at the 3e-3 level)
        torch.manual_seed(42)
        sd = O.init_state_dict(10, '320')
        tr = O.Trainer(sd, classifier_config='320', num_classes=10)
        x, t = O.synthetic_batch(4, 64, 64, 10)
        _, loss, g = tr.grads(x[r * 2:r * 2 + 2], t[r * 2:r * 2 + 2], dropout_masks="off")
        gs.append(torch.cat([g[n].reshape(-1) for n in tr.names[::-1]]))
        assert abs(loss.item() - res[r][1]) < 1e-5
    torch.set_num_threads(nthreads)
    ref = (gs[0] + gs[1]) / 2
    # (a loaded host can change the CPU kernels' work split between the worker and this process: one failure in ~20 runs
    # at rtol 1e-5 / atol 1e-7; the all-reduce itself is exact up to fp32 summation order)
    torch.testing.assert_close(res[0][2], ref, rtol=1e-4, atol=1e-6)


def test_bucket_schedule_monotone():
    from mnb200 import ddp
    sched = ddp.bucket_schedule([(0, 10), (10, 20), (20, 30)], [(0, 50), (5, 7), (10, 3), (20, 99)])
    assert sched == [(50, 0, 10), (50, 10, 20), (99, 20, 30)]
