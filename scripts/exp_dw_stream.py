"""GPU experiment for the depthwise row-stream kernels (csrc/dwconv_stream.cu, option "dw_stream"): parity against the
shared-memory tile kernels and stand-alone timings on every depthwise layer shape of MNASNet-224 at batch 256, then the
whole training step with the option off / on.  Writes gpurun_out/exp_dw_stream.json incrementally.

    python scripts/exp_dw_stream.py [layers] [net]
"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "exp_dw_stream.json")
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

BF = torch.bfloat16
dev = "cuda"
# (H = W, C, k) of the depthwise ConvBlocks, batch 256
LAYERS = [(112, 32, 3), (112, 48, 3), (56, 72, 5), (28, 240, 5), (14, 480, 3), (14, 576, 5), (7, 1152, 3)]


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
    return e0.elapsed_time(e1) / n * 1e3


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def layer_case(N, H, C, k):
    W = H
    g = torch.Generator(device=dev).manual_seed(H * 7 + C)
    x = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)
    dz = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device=dev, generator=g) / k).float()
    sc = (torch.rand(C, device=dev, generator=g) + 0.5).float()
    sh = (torch.randn(C, device=dev, generator=g) * 0.3).float()
    out = {"shape": f"{N}x{H}x{W}x{C} k{k}", "bytes_fwd": 2 * x.numel() * 2}
    res = {}
    for tag, opt, pd, tw8 in (("tile", 0, 1, 0), ("stream", 1, 1, 0), ("stream_pd2", 1, 2, 0), ("stream_pd3", 1, 3, 0),
                              ("stream_tw8", 1, 1, 1), ("stream_tw8_pd3", 1, 3, 1)):
        if tw8 and k != 3:
            continue
        L.set_option("dw_stream", opt)
        L.set_option("dw_stream_pd", pd)
        L.set_option("dw_stream_tw8", tw8)
        z = torch.full_like(x, float("nan"))
        st = torch.zeros(2 * C, device=dev, dtype=torch.float64)
        dx = torch.full_like(x, float("nan"))
        dw = torch.zeros(C, 1, k, k, device=dev)

        def f_fwd(z=z, st=st):
            L.call("mnb_dw_fwd", P(x), P(sc), P(sh), P(w), None, P(z), P(st), N, H, W, C, k, 1, S())

        def f_dg(dx=dx):
            L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())

        def f_wg(dw=dw):
            L.call("mnb_dw_wgrad", P(x), P(sc), P(sh), P(dz), P(dw), N, H, W, C, k, 1, S())
        f_fwd(); f_dg(); f_wg()
        torch.cuda.synchronize()
        res[tag] = (z.clone(), st.clone(), dx.clone(), dw.clone())
        out[f"fwd_us_{tag}"] = round(timeit(f_fwd), 1)
        out[f"dgrad_us_{tag}"] = round(timeit(f_dg), 1)
        out[f"wgrad_us_{tag}"] = round(timeit(f_wg), 1)
    L.set_option("dw_stream", 0)
    L.set_option("dw_stream_pd", 1)
    L.set_option("dw_stream_tw8", 0)
    z0, st0, dx0, dw0 = res["tile"]
    ok = True
    for tag in [t for t in res if t != "tile"]:
        z1, st1, dx1, dw1 = res[tag]
        nan = int(torch.isnan(z1.float()).sum().item() + torch.isnan(dx1.float()).sum().item())
        rels = (relerr(z1.float(), z0.float()), relerr(st1, st0), relerr(dx1.float(), dx0.float()), relerr(dw1, dw0))
        out[f"rel_{tag}"] = [float(f"{r:.2e}") for r in rels]          # fwd, stats, dgrad, wgrad
        ok = ok and nan == 0 and all(r < 5e-3 for r in rels)
    out["ok"] = bool(ok)
    return out


def net_case(n_big, steps):
    from mnb200 import engine
    from models.classifiers import FineTuneModelPool, load_model
    xb = torch.randn(n_big, 3, 224, 224, device=dev)
    tb = torch.randint(0, 1000, (n_big,), device=dev)
    big = {}
    for opt, pd, tw8 in ((0, 1, 0), (1, 1, 0), (2, 1, 0), (1, 3, 0), (2, 3, 0), (1, 3, 1)):
        L.set_option("dw_stream", opt)              # graphs captured now keep the kernels selected now
        L.set_option("dw_stream_pd", pd)
        L.set_option("dw_stream_tw8", tw8)
        torch.manual_seed(42)
        m = FineTuneModelPool(load_model('mnasnet'), 'mnasnet', 1000, '512')
        engine.configure(m, dtype="bf16")
        m = m.cuda().train()
        eng = engine.engine_for(m)
        first = None
        for _ in range(3):
            l = eng.train_step_graph(xb, tb, lr=1e-3).item()
            first = l if first is None else first
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            l = eng.train_step_graph(xb, tb, lr=1e-3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tag = ("tile", "stream", "auto")[opt] + (f"_pd{pd}" if pd > 1 else "") + ("_tw8" if tw8 else "")
        big[f"ms_per_step_{tag}"] = round(ms, 3)
        big[f"img_per_s_{tag}"] = round(n_big / ms * 1e3, 1)
        big[f"loss_first_{tag}"] = first
        big[f"loss_last_{tag}"] = l.item()
        del eng, m
        torch.cuda.empty_cache()
    L.set_option("dw_stream", 0)
    L.set_option("dw_stream_pd", 1)
    L.set_option("dw_stream_tw8", 0)
    item("net_big", **big)


def main():
    only = sys.argv[1:] or ["layers", "net"]
    try:
        item("import", torch=torch.__version__, dev=torch.cuda.get_device_name(0))
        if "layers" in only:
            item("small", **layer_case(2, 14, 72, 5))
            for (H, C, k) in LAYERS:
                item("layer", **layer_case(256, H, C, k))
        if "net" in only:
            net_case(256, 10)
    except Exception as e:
        RES["fatal"] = f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"
        flush()
        print(RES["fatal"])
        sys.exit(1)
    flush()
    oks = [i.get("ok") for i in RES["items"] if "ok" in i]
    print(f"EXP done: {sum(bool(o) for o in oks)}/{len(oks)} layer items ok, elapsed {RES['elapsed_s']} s")


if __name__ == "__main__":
    main()
