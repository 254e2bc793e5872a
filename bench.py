#!/usr/bin/env python
"""MNASNet-224 training throughput on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (forward + CrossEntropyLoss + backward + [NCCL gradient all-reduce] + Adam)
over one synthetic batch of 256 images/GPU at 3x224x224, bf16 activations (BASELINE.json configs[1]; weak
scaling for N > 1 = configs[2] at 256/GPU).  Prints ONE JSON line on rank 0.

  value    images/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      the reference's own loop (src/train.py:427-440: .to(device), model(input), criterion, zero_grad,
           backward, optimizer.step) through the drop-in modules with pinned HOST inputs: the H2D copy of every
           step's batch and the D2H read of its loss are inside the timed region
  roofline dominant kernel FAMILY of the step (depthwise / pointwise / dense 3x3; the BatchNorm passes carry no
           algorithmic bytes of their own under SURVEY.md section 8d): section-8d bytes of that family's conv calls /
           CUDA-event time of its launches vs MEASURED_PEAKS.json; `families` holds the same for every family
  cpu_baseline  the CPU oracle (port of the reference loop, fp32, torch CPU kernels) on the host's cores
  gpu_eager_baseline  the same architecture through stock PyTorch eager on this GPU (bf16 autocast, channels_last)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "mnasnet-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "MNASNet-224 train images/sec"
WORKLOAD = "MnasNet (cut_channels_first=False, head '512', 1000 classes) 224x224 bf16 training, batch 256/GPU"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """Reference arm: the reference's training loop (CPU oracle port: same torch CPU kernels as the
    reference modules, oracle/make_golden.py proves bit-identity) on all host cores.  Each step is a bounded
    sample of the workload: batch `--ref-batch` (default 32) instead of 256, fp32 (the reference has no bf16)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import mnasnet_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    nb = args.ref_batch
    torch.manual_seed(42)
    sd = O.init_state_dict()
    tr = O.Trainer(sd)
    x, t = O.synthetic_batch(nb, 224, 224)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tr.step(x, t, dropout_masks=None)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    val = nb * len(times) / tot
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"batch {nb} per step on the host CPU (fp32, NCHW)"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(times)} steps x batch {nb}, 3x224x224 fp32, fwd+CE+bwd+Adam"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(seconds_budget=20.0):
    import torch
    from oracle import mnasnet_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    sd = O.init_state_dict()
    tr = O.Trainer(sd)
    nb = 8
    x, t = O.synthetic_batch(nb, 224, 224)
    times, t_start = [], time.perf_counter()
    for i in range(2 + 50):
        t0 = time.perf_counter()
        tr.step(x, t, dropout_masks=None)
        dt = time.perf_counter() - t0
        if i >= 2:
            times.append(dt)
        if time.perf_counter() - t_start > seconds_budget and len(times) >= 3:
            break
    med = statistics.median(times)
    return {"value": nb / med, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"BASELINE configs[0]: batch {nb}, 3x224x224 fp32 fwd+CE+bwd+Adam, median of {len(times)} "
                      f"steps after 2 warm-up ({med * 1e3:.0f} ms/step)"}


def gpu_eager_baseline(n, size, steps=5, warm=3):
    """Stock PyTorch on the same GPU: the reference's architecture and loop (oracle port = the reference's own torch
    ops; a checker-side baseline like cpu_baseline, never on the product path) under bf16 autocast + channels_last,
    cuDNN autotuned.  This is what the reference would run on a B200 today."""
    import torch
    from oracle import mnasnet_oracle as O
    try:
        old = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True
        torch.manual_seed(42)
        sd = {k: v.cuda() for k, v in O.init_state_dict().items()}
        tr = O.Trainer(sd)
        x = torch.randn(n, 3, size, size, device="cuda").contiguous(memory_format=torch.channels_last)
        t = torch.randint(0, 1000, (n,), device="cuda")

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return tr.step(x, t, dropout_masks=None)
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        torch.backends.cudnn.benchmark = old
        del tr, sd, x
        torch.cuda.empty_cache()
        return {"value": n / (ms / 1e3), "unit": "images/s", "ms_per_step": ms,
                "what": f"torch {torch.__version__} eager, autocast bf16 + channels_last, cuDNN benchmark, batch {n}, "
                        f"{steps} steps after {warm} warm-up, same GPU"}
    except Exception as e:          # never lose the main measurement over a context number
        return {"value": None, "error": repr(e)[:200]}


FAMILIES = ("depthwise", "pointwise", "dense3x3", "batchnorm", "other")


def family_of(label):
    if label.startswith("dw"):
        return "depthwise"
    if label.startswith("pw1x1"):
        return "pointwise"
    if label.startswith("conv3x3"):
        return "dense3x3"
    if label.startswith("bn_") or label.startswith("mnb_bn_"):
        return "batchnorm"
    return "other"


def algorithmic_work(plan):
    """SURVEY.md section 8d per family for this plan's shapes: bytes = fwd (|X|+|Y|) + bwd (2|X|+2|Y|) per conv call at
    the stored element size, flops = 6 * MACs (fwd 2, bwd 4)."""
    out = {f: {"bytes": 0.0, "flops": 0.0} for f in FAMILIES}
    for a in plan.apps:
        conv = a.cb.conv
        fam = "depthwise" if a.kind == "dw" else ("pointwise" if a.k == 1 else "dense3x3")
        out[fam]["bytes"] += 3.0 * (a.in_bytes + a.out_bytes)
        out[fam]["flops"] += 6.0 * a.m * a.Cout * a.k * a.k * (conv.in_channels // conv.groups)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mnb200")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--height", type=int, default=0, help="rectangular inputs (BASELINE configs[4]): H, with --width")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE configs[2]): fixed global batch split over the ranks")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--ref-batch", type=int, default=32)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table (JSON) here")
    ap.add_argument("--dump-ops", default=None, help="write the launch sequence of one step (labels) here")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mnb200 import ddp, engine
    from models.classifiers import FineTuneModelPool, load_model

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import contextlib, io
    torch.manual_seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        model = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
    engine.configure(model, dtype=args.dtype, impl=args.gemm)
    model = model.to(dev).train()                      # Dropout ACTIVE, BN in train mode: the real step
    eng = engine.engine_for(model)
    sync = None
    if world > 1:
        ddp.broadcast_parameters(eng)
        sync = ddp.GradSync(eng)
    N, S = args.batch, args.size
    if args.global_batch:
        N = args.global_batch // world
    SH, SW = (args.height or S), (args.width or S)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    nbuf = 2                                           # alternate two resident batches (each 154 MB > L2)
    xs = [torch.randn(N, 3, SH, SW, device=dev, generator=g) for _ in range(nbuf)]
    ts = [torch.randint(0, 1000, (N,), device=dev, generator=g) for _ in range(nbuf)]
    if args.dump_ops and rank == 0:
        plan = eng.plan(N, SH, SW)
        seq = []
        for op in plan.fwd:
            seq.append([getattr(op, "label", op.name), getattr(op, "detail", ""), op.name])
        seq.append(["xent", "", "mnb_xent_fwd_bwd"])
        for op in plan.bwd:
            if op.name.startswith("event_"):
                continue
            seq.append([getattr(op, "label", op.name), getattr(op, "detail", ""), op.name])
            if op.name == "mnb_fc_wgrad":
                seq.append(["fc_wgrad_bias", "", "mnb_fc_wgrad"])
        seq += [["counter_inc", "", "mnb_counter_inc"], ["adam", "", "mnb_adam_step"]]
        with open(args.dump_ops, "w") as f:
            json.dump(seq, f)
    use_graph = world == 1 and not args.no_graph
    step_fn = eng.train_step_graph if use_graph else eng.train_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_fn(xs[i % nbuf], ts[i % nbuf], lr=1e-3)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step_fn(xs[i % nbuf], ts[i % nbuf], lr=1e-3)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    last_loss = float(loss.item())
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = world * N * args.steps / (ms / 1e3)
    scaling = "strong" if args.global_batch else "weak"
    launches = eng.launches_per_step(N, SH, SW) * args.steps

    # ---- end to end through the reference's own loop with pinned host inputs ---------------------------
    # e2e: the loader hands over uint8 HWC batches as decoded (utils/datasets.py:456-462 before ToTensor); ToTensor +
    # Normalize run inside the stem kernels, so a step ships 1 byte per value.  e2e_fp32_input: the same loop fed the
    # already-normalised fp32 NCHW tensor exactly as src/train.py:427 does (4 bytes per value).
    e2e, e2e_f32 = None, None
    if not args.no_e2e:
        crit = torch.nn.CrossEntropyLoss()
        opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3)
        copy_stream = torch.cuda.Stream()

        def measure(u8):
            if u8:
                hx = [torch.randint(0, 256, (N, SH, SW, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
            else:
                hx = [torch.randn(N, 3, SH, SW).pin_memory() for _ in range(2)]
            ht = [torch.randint(0, 1000, (N,)).pin_memory() for _ in range(2)]

            class Prefetch:                              # the DataLoader side of train.py:423-431
                def __init__(self):
                    self.i = 0
                    self.nxt = None
                    self.load()

                def load(self):
                    with torch.cuda.stream(copy_stream):
                        x = hx[self.i % 2].to(dev, non_blocking=True)
                        if not u8:
                            x = x.float()
                        t = ht[self.i % 2].to(dev, non_blocking=True).long()
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    self.nxt = (x, t, ev)
                    self.i += 1

                def get(self):
                    x, t, ev = self.nxt
                    torch.cuda.current_stream().wait_event(ev)
                    x.record_stream(torch.cuda.current_stream())
                    t.record_stream(torch.cuda.current_stream())
                    self.load()
                    return x, t

            def loop(k, pf):
                last = None
                for _ in range(k):
                    inp, tgt = pf.get()
                    out = model(inp)
                    l = crit(out, tgt)
                    opt.zero_grad()
                    l.backward()                         # DDP: buckets all-reduced + averaged inside backward
                    opt.step()
                    last = l.item()                      # D2H read of the step's loss (train.py:447)
                return last
            pf = Prefetch()
            loop(max(3, args.warmup // 2), pf)
            barrier()
            t0 = time.perf_counter()
            loop(args.steps, pf)
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                tm = torch.tensor([dt], device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dt = float(tm.item())
            return {"value": world * N * args.steps / dt, "unit": "images/s",
                    "h2d_bytes_per_step": N * 3 * SH * SW * (1 if u8 else 4) + N * 8, "d2h_bytes_per_step": 4,
                    "api": "src/train.py:423-440 loop on the drop-in models (model(input), CrossEntropyLoss, zero_grad, "
                           "backward, torch.optim.Adam.step); pinned host batch prefetched on a copy stream; input "
                           + ("uint8 N x H x W x 3, ToTensor + Normalize fused into the stem kernels" if u8 else
                              "fp32 N x 3 x H x W, already normalised (train.py:427)")}
        e2e = measure(True)
        e2e_f32 = measure(False)

    # ---- roofline pass: per-launch CUDA-event timing of every kernel of the step (rank 0) ----------------
    roof, table = None, None
    if rank == 0:
        hbm, tf, which = peaks()
        eng.profile = []
        hook, eng.grad_hook = eng.grad_hook, None
        for i in range(3):
            eng.train_step(xs[i % nbuf], ts[i % nbuf], lr=0.0)
        torch.cuda.synchronize()
        rec, eng.profile, eng.grad_hook = eng.profile, None, hook
        agg = {}
        per_op = {}
        for label, nbytes, a, b, detail in rec:
            d = agg.setdefault(label, [0, 0.0, 0])
            ms_ = a.elapsed_time(b)
            d[0] += 1
            d[1] += ms_
            d[2] += nbytes
            q = per_op.setdefault((label, detail), [0, 0.0, 0])
            q[0] += 1; q[1] += ms_; q[2] += nbytes
        tot = sum(v[1] for v in agg.values())
        table = sorted(({"kernel": k, "family": family_of(k), "launches_per_step": v[0] // 3, "ms_per_step": v[1] / 3,
                         "share": v[1] / tot, "kernel_traffic_GB_per_step": v[2] / 3 / 1e9,
                         "achieved_GBps": (v[2] / 1e9) / (v[1] / 1e3) if v[1] > 0 else 0.0,
                         "frac_of_hbm_peak": ((v[2] / 1e9) / (v[1] / 1e3)) / hbm if v[1] > 0 else 0.0}
                        for k, v in agg.items()), key=lambda r: -r["ms_per_step"])
        # per family: time of its launches (CUDA events) vs the section-8d algorithmic bytes / flops of its conv calls
        work = algorithmic_work(eng.plan(N, SH, SW))
        ncu = None
        try:    # dram__bytes_read+write per kernel class from the committed ncu pass of the same command (profiles/)
            with open(os.path.join(ROOT, "profiles", "r2_ncu_step_summary.json")) as f:
                ncu = json.load(f)["classes"]
        except Exception:
            ncu = None
        families = {}
        for fam in FAMILIES:
            ms_f = sum(r["ms_per_step"] for r in table if r["family"] == fam)
            gb, gf = work[fam]["bytes"] / 1e9, work[fam]["flops"] / 1e9
            d = {"ms_per_step": ms_f, "share": ms_f * 3 / tot, "algorithmic_GB": gb, "algorithmic_GFLOP": gf}
            if ms_f > 0 and gb > 0:
                d["achieved_GBps"] = gb / (ms_f / 1e3)
                d["frac_of_hbm_peak"] = d["achieved_GBps"] / hbm
                d["achieved_TFLOPs"] = gf / ms_f
                d["frac_of_bf16_peak"] = d["achieved_TFLOPs"] / tf
            if ncu is not None:
                tr_ = [ncu[r["kernel"]]["dram_GB_per_step"] for r in table if r["family"] == fam and r["kernel"] in ncu]
                d["dram_GB_ncu"] = sum(tr_) if tr_ else None
            families[fam] = d
        families["batchnorm"]["note"] = "BN-backward / residual passes: 0 algorithmic bytes of their own under section 8d"
        top_f = max(("depthwise", "pointwise", "dense3x3"), key=lambda f_: families[f_]["ms_per_step"])
        top = families[top_f]
        roof = {"bound": "hbm", "kernel": top_f + " family: " + ", ".join(sorted(
                    {r["kernel"] for r in table if r["family"] == top_f})),
                "achieved": top["achieved_GBps"], "peak": hbm, "unit": "GB/s", "frac": top["frac_of_hbm_peak"],
                "traffic": top.get("dram_GB_ncu"),
                "traffic_unit": "GB per step over this family's launches (ncu dram__bytes_read+write, profiles/r2_ncu_step_summary.json)",
                "algorithmic_GB_per_step": top["algorithmic_GB"], "ms_per_step": top["ms_per_step"],
                "peak_source": which, "share_of_step": top["share"],
                "families": families,
                "step_algorithmic_GB": sum(families[f_]["algorithmic_GB"] for f_ in FAMILIES),
                "step_kernel_traffic_GB": sum(r["kernel_traffic_GB_per_step"] for r in table),
                "step_frac_of_hbm_peak": (value / world) * sum(families[f_]["algorithmic_GB"] for f_ in FAMILIES) / N / hbm}
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                layers = sorted(({"kernel": k[0], "layer": k[1], "launches_per_step": v[0] // 3,
                                  "ms_per_step": v[1] / 3, "algorithmic_GB_per_step": v[2] / 3 / 1e9,
                                  "achieved_GBps": (v[2] / 1e9) / (v[1] / 1e3) if v[1] > 0 else 0.0}
                                 for k, v in per_op.items()), key=lambda r: -r["ms_per_step"])
                json.dump({"ms_per_step_sum": tot / 3, "families": families, "kernels": table, "layers": layers[:120]},
                          f, indent=1)

    cpu, eager = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        eager = gpu_eager_baseline(N, S) if (SH, SW) == (S, S) else None
        cpu = cpu_baseline()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": WORKLOAD if (N, SH, SW) == (256, 224, 224) else
                           f"MnasNet (cut_channels_first=False, head '512') {SH}x{SW} bf16 training, batch {N}/GPU",
                           "global_batch": N * world, "parallelism": f"dp{world}", "optimizer": "Adam lr 1e-3",
                           "dropout": "active", "cuda_graph": bool(use_graph), "gemm_impl": args.gemm,
                           "kernel_options": _kernel_options(),
                           "l2": "no flush needed: each step streams a 154 MB input batch (2 alternating "
                                 "buffers) and ~6 GB of activations, both >> 126 MB L2"},
                "e2e": e2e, "e2e_fp32_input": e2e_f32, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "gpu_eager_baseline": eager, "final_loss": last_loss}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _kernel_options():
    """Which kernel families `impl` auto selects (mnb_get_option): recorded with every bench line."""
    try:
        from mnb200 import _lib
        return {n: _lib.get_option(n) for n in ("pw_stream", "stem_mma", "dw_mma", "dw_small", "c3_mma")}
    except Exception as e:      # never lose a measurement over a label
        return {"error": str(e)}


if __name__ == "__main__":
    main()
