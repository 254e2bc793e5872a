"""One-shot GPU experiment for the warp-streaming 1x1 kernels (csrc/pw_stream.cu, impl=3) and the tensor-pipe
stem backward-weight kernel: parity against the tcgen05 / SIMT kernels they would replace (which are themselves
parity-tested against the CPU checker in tests/), stand-alone timings on the real layer shapes, and a whole-step
comparison Engine(impl="auto") vs Engine(impl="stream").  Writes gpurun_out/exp_stream.json after every item so a
cut-off run still leaves results.

    python scripts/exp_stream.py            # everything
"""
import json
import math
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "exp_stream.json")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
T0 = time.time()
RES = {"items": [], "fatal": None}


def flush():
    RES["elapsed_s"] = round(time.time() - T0, 1)
    with open(OUT, "w") as f:
        json.dump(RES, f, indent=1)


def item(name, **kw):
    kw["name"] = name
    kw["t"] = round(time.time() - T0, 1)
    RES["items"].append(kw)
    print(json.dumps(kw), flush=True)
    flush()


flush()
import torch  # noqa: E402

from mnb200 import _lib as L  # noqa: E402

item("import", torch=torch.__version__, dev=torch.cuda.get_device_name(0))
dev = "cuda"
BF = torch.bfloat16


def P(t):
    return None if t is None else t.data_ptr()


def S():
    return torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3      # us


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ulp_stats(a, b):
    """bf16 tensors: share of elements that differ, share outside (4 bf16 ulps + a cancellation floor), rel-L2."""
    a, b = a.float(), b.float()
    d = (a - b).abs()
    tol = torch.maximum(b.abs(), a.abs()) * 2.0 ** -6 + 1e-4 * b.abs().mean()
    return {"differ": (d > 0).float().mean().item(), "bad": int((d > tol).sum().item()), "rel_l2": relerr(a, b)}


def pw_case(N, H, W, Cin, Cout, xform, timing):
    g = torch.Generator(device=dev).manual_seed(N * 1000 + Cin * 7 + Cout)
    M = N * H * W
    x = torch.randn(M, Cin, device=dev, generator=g).to(BF)
    w = (torch.randn(Cout, Cin, 1, 1, device=dev, generator=g) / math.sqrt(Cin)).float()
    b = (torch.randn(Cout, device=dev, generator=g) * 0.1).float()
    sc = (torch.rand(Cin, device=dev, generator=g) + 0.5).float() if xform else None
    sh = (torch.randn(Cin, device=dev, generator=g) * 0.3).float() if xform else None
    pf = torch.empty(Cout * Cin, device=dev, dtype=BF)
    pd = torch.empty(Cout * Cin, device=dev, dtype=BF)
    L.call("mnb_pack_weights", P(w), P(pf), P(pd), Cout, Cin, 1, S())
    out = {"shape": f"{N}x{H}x{W} {Cin}->{Cout} xf={int(xform)}"}
    # forward
    zs, sts = {}, {}
    for impl in (2, 3):
        z = torch.full((M, Cout), float("nan"), device=dev, dtype=BF)
        st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64)

        def f(z=z, st=st, impl=impl):
            L.call("mnb_conv_fwd_packed", P(x), P(sc), P(sh), P(w), P(pf), P(b), P(z), P(st), N, H, W, Cin, Cout, 1, 1, 0,
                   1, 0, impl, S())
        f()
        torch.cuda.synchronize()
        zs[impl], sts[impl] = z, st.clone()
        if timing:
            out[f"fwd_us_impl{impl}"] = round(timeit(f), 1)
    out["fwd"] = ulp_stats(zs[3], zs[2])
    out["fwd_nan"] = int(torch.isnan(zs[3].float()).sum().item())
    out["fwd_stats_rel"] = relerr(sts[3], sts[2])
    # dgrad (with residual add)
    dz = torch.randn(M, Cout, device=dev, generator=g).to(BF)
    add = torch.randn(M, Cin, device=dev, generator=g).to(BF)
    for use_add in (True, False):
        dxs = {}
        for impl in (2, 3):
            dx = torch.full((M, Cin), float("nan"), device=dev, dtype=BF)

            def f(dx=dx, impl=impl):
                L.call("mnb_conv_dgrad_packed", P(dz), P(w), P(pd), P(add) if use_add else None, P(dx), None, None, None,
                       None, N, H, W, Cin, Cout, 1, 1, 0, 1, impl, S())
            f()
            torch.cuda.synchronize()
            dxs[impl] = dx
            if timing and use_add:
                out[f"dgrad_us_impl{impl}"] = round(timeit(f), 1)
        out["dgrad_add" if use_add else "dgrad"] = ulp_stats(dxs[3], dxs[2])
        out["dgrad_nan"] = out.get("dgrad_nan", 0) + int(torch.isnan(dxs[3].float()).sum().item())
    # wgrad
    dws = {}
    for impl in (2, 3):
        dw = torch.zeros(Cout, Cin, device=dev, dtype=torch.float32)

        def f(dw=dw, impl=impl):
            L.call("mnb_conv_wgrad", P(x), P(sc), P(sh), P(dz), P(dw), N, H, W, Cin, Cout, 1, 1, 0, 1, 0, impl, S())
        f()
        torch.cuda.synchronize()
        dws[impl] = dw.clone()
        if timing:
            out[f"wgrad_us_impl{impl}"] = round(timeit(f), 1)
    out["wgrad_rel"] = relerr(dws[3], dws[2])
    # independent check of the streaming results against torch fp32 math on the same rounded operands
    a = x.float()
    if xform:
        a = torch.relu(a * sc + sh).to(BF).float()
    zref = a @ w.view(Cout, Cin).to(BF).float().t() + b
    out["fwd_vs_torch"] = relerr(zs[3].float(), zref)
    out["wgrad_vs_torch"] = relerr(dws[3], dz.float().t() @ a)
    ok = (out["fwd_nan"] == 0 and out["dgrad_nan"] == 0 and
          all(out[k]["bad"] == 0 and out[k]["rel_l2"] < 2e-3 for k in ("fwd", "dgrad_add", "dgrad")) and
          out["fwd_stats_rel"] < 1e-4 and out["wgrad_rel"] < 1e-4 and out["fwd_vs_torch"] < 5e-3 and
          out["wgrad_vs_torch"] < 1e-4)
    out["ok"] = bool(ok)
    return out


def stem_case(N, H, W, timing):
    g = torch.Generator(device=dev).manual_seed(N + H)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    x = torch.randn(N, 3, H, W, device=dev, generator=g)
    dz = torch.randn(N, Ho, Wo, 32, device=dev, generator=g).to(BF)
    out = {"shape": f"{N}x{H}x{W}"}
    dws = {}
    for impl in (0, 3):
        dw = torch.zeros(32, 3, 3, 3, device=dev, dtype=torch.float32)

        def f(dw=dw, impl=impl):
            L.call("mnb_conv_wgrad", P(x), None, None, P(dz), P(dw), N, H, W, 3, 32, 3, 2, 1, 1, 1, impl, S())
        f()
        torch.cuda.synchronize()
        dws[impl] = dw.clone()
        if timing:
            out[f"us_impl{impl}"] = round(timeit(f, n=5, warm=1), 1)
    out["rel_vs_old"] = relerr(dws[3], dws[0])
    out["ok"] = bool(out["rel_vs_old"] < 5e-3)          # x is rounded to bf16 in the new kernel
    return out


def net_case(n_big, steps):
    from mnb200 import engine
    from models.classifiers import FineTuneModelPool, load_model
    from oracle import mnasnet_oracle as O
    out = {}
    # small step against the CPU checker (same recipe as __graft_entry__.smoke)
    n, h, w = 4, 64, 64
    x, t = O.synthetic_batch(n, h, w)
    torch.manual_seed(42)
    tr = O.Trainer(O.init_state_dict())
    _, oloss = tr.step(x, t, dropout_masks="off")
    grads = {}
    for impl in ("auto", "stream"):
        torch.manual_seed(42)
        m = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
        engine.configure(m, dtype="bf16", impl=impl)
        m = m.cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.eval()
        eng = engine.engine_for(m)
        loss = eng.train_step(x.cuda(), t.cuda(), lr=1e-3)
        torch.cuda.synchronize()
        out[f"small_loss_{impl}"] = loss.item()
        out[f"small_loss_rel_oracle_{impl}"] = abs(loss.item() - oloss.item()) / oloss.item()
        grads[impl] = eng.store.grad.clone()
    out["small_grad_rel_stream_vs_auto"] = relerr(grads["stream"], grads["auto"])
    item("net_small", **out)
    # full-size step timing, graph replay
    big = {}
    xb = torch.randn(n_big, 3, 224, 224, device=dev)
    tb = torch.randint(0, 1000, (n_big,), device=dev)
    for impl in ("auto", "stream"):
        torch.manual_seed(42)
        m = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
        engine.configure(m, dtype="bf16", impl=impl)
        m = m.cuda().train()
        eng = engine.engine_for(m)
        losses = []
        for _ in range(3):
            losses.append(eng.train_step_graph(xb, tb, lr=1e-3).item())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            l = eng.train_step_graph(xb, tb, lr=1e-3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        big[f"ms_per_step_{impl}"] = round(ms, 3)
        big[f"img_per_s_{impl}"] = round(n_big / ms * 1e3, 1)
        big[f"loss_first_{impl}"] = losses[0]
        big[f"loss_last_{impl}"] = l.item()
        del eng, m
        torch.cuda.empty_cache()
    item("net_big", **big)


def main():
    only = sys.argv[1:] or ["pw_small", "stem_small", "net", "pw_big", "stem_big"]
    PW = [(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)]
    try:
        if "pw_small" in only:
            for cin, cout in PW:
                for (N, H, W) in ((1, 5, 7), (2, 16, 16)):
                    for xf in (True, False):
                        item("pw_small", **pw_case(N, H, W, cin, cout, xf, False))
        if "stem_small" in only:
            for (N, H, W) in ((2, 30, 26), (3, 64, 64), (1, 17, 23)):
                item("stem_small", **stem_case(N, H, W, False))
        if "net" in only:
            net_case(256, 10)
        if "pw_big" in only:
            for (H, cin, cout) in ((112, 16, 48), (112, 48, 16), (112, 32, 16), (56, 24, 72), (56, 72, 24)):
                item("pw_big", **pw_case(256, H, H, cin, cout, True, True))
        if "stem_big" in only:
            item("stem_big", **stem_case(256, 224, 224, True))
    except Exception as e:  # a CUDA fault poisons the context: record and stop
        RES["fatal"] = f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"
        flush()
        print(RES["fatal"])
        sys.exit(1)
    flush()
    oks = [i.get("ok") for i in RES["items"] if "ok" in i]
    print(f"EXP done: {sum(bool(o) for o in oks)}/{len(oks)} kernel items ok, elapsed {RES['elapsed_s']} s")


if __name__ == "__main__":
    main()
