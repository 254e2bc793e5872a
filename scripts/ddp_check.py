"""T6: R ranks (one per GPU) vs the CPU oracle run per shard with gradients averaged.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 scripts/ddp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
import contextlib
import io

import torch
import torch.distributed as dist

from mnb200 import ddp, engine
from models.classifiers import FineTuneModelPool, load_model
from oracle import mnasnet_oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("MNB_DDP_BACKEND", "nccl")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        # gloo moves CUDA tensors too: lets R ranks share ONE GPU, so the data-parallel path (bucket schedule,
        # overlapped all-reduce, averaged gradients, lock-step Adam) is testable on a single-GPU box
        torch.cuda.set_device(local % torch.cuda.device_count())
        dist.init_process_group(backend)
    per, h, w = 4, 96, 96
    torch.manual_seed(42 + rank)              # deliberately different init per rank: broadcast must fix it
    with contextlib.redirect_stdout(io.StringIO()):
        m = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
    engine.configure(m, dtype="fp32")
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    eng = engine.engine_for(m)
    if os.environ.get("MNB_NO_SIDE"):
        eng.wgrad_slack = 0
    if rank == 0:                              # rank 0 carries the reference weights
        torch.manual_seed(42)
        sd0 = O.init_state_dict()
        m.load_state_dict(sd0)
    wrapped = ddp.DataParallel(m)             # = nn.DataParallel(model) of src/train.py:202: broadcast + GradSync
    assert wrapped.module is m and eng.grad_hook is wrapped.sync
    assert all(k.startswith("module.") for k in wrapped.state_dict()) and len(wrapped.state_dict()) == 403
    x, t = O.synthetic_batch(per * world, h, w)
    xs, ts = x[rank * per:(rank + 1) * per].cuda(), t[rank * per:(rank + 1) * per].cuda()
    out = wrapped(xs)
    loss = torch.nn.CrossEntropyLoss()(out, ts)
    loss.backward()                            # bucketed all-reduce overlapped with backward, averaged
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    ok = True
    if rank == 0:
        acc = None
        for r in range(world):
            torch.manual_seed(42)
            sd = O.init_state_dict(dtype=torch.float64)
            tr = O.Trainer(sd)
            _, l, g = tr.grads(x[r * per:(r + 1) * per].double(), t[r * per:(r + 1) * per], dropout_masks="off")
            acc = g if acc is None else {k: acc[k] + g[k] for k in g}
            if r == 0:
                c = abs(l.item() - loss.item()) / l.item() < 1e-4
                print("  loss check", c, l.item(), loss.item())
                ok &= c
                bn0 = sd["features.0.bn.running_mean"].clone()
        names = [k for k in acc if not k.endswith("conv.bias")]
        a = torch.cat([ours[k].double().reshape(-1) for k in names])
        b = torch.cat([(acc[k] / world).reshape(-1) for k in names])
        err = ((a - b).norm() / b.norm()).item()
        rm = (m.features[0].bn.running_mean.cpu().double() - bn0).abs().max().item()
        print(f"ddp_check world={world}: averaged-grad rel-L2 vs per-shard fp64 oracle {err:.3e}; rank-0 BN buffer diff {rm:.2e}")
        ok &= err < 2.5e-2 and rm < 1e-5        # gradient noise floor of the fp32 oracle itself is 1e-2 (F9)
        # the last linear layer is well conditioned (tight); layers behind a ReLU see occasional fp32-vs-fp64 mask
        # flips (observed once in five 8-rank runs: classifier.1 7e-3 with classifier.4 at 1e-5)
        for k, tol in (("classifier.4.weight", 1e-4), ("classifier.1.weight", 2.5e-2)):
            e = ((ours[k].double() - acc[k] / world).norm() / (acc[k] / world).norm()).item()
            print("  head grad", k, f"{e:.3e}")
            ok &= e < tol
    # every rank holds identical gradients
    flat = torch.cat([ours[k].reshape(-1) for k in sorted(ours)]).cuda()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(flat, ref)
    if not same:
        print(f"  rank {rank}: gradients differ from rank 0: max abs {(flat - ref).abs().max().item():.3e}")
    print(f"  rank {rank}: ok={ok} same={same}", flush=True)
    flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # two fused steps keep the replicas in lock-step (identical Adam update on every rank)
    for _ in range(2):
        eng.train_step(xs, ts, lr=1e-3)
    torch.cuda.synchronize()
    p = eng.store.flat.clone()
    dist.broadcast(p, 0)
    lock = torch.equal(p, eng.store.flat)
    flag2 = torch.tensor([1 if lock else 0], device="cuda")
    dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP_CHECK", "PASS" if int(flag) == 1 and int(flag2) == 1 else "FAIL", int(flag), int(flag2))
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 and int(flag2) == 1 else 1)


if __name__ == "__main__":
    main()
