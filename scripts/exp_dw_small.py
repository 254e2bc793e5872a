"""GPU experiment for the whole-tile depthwise kernels (csrc/dw_small.cu): parity against torch fp32 math on the same bf16
operands and stand-alone timings at batch 256 against the row-streaming kernels (dw_small = 0).
   python scripts/exp_dw_small.py [parity] [time]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from mnb200 import _lib as L
BF = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream
torch.backends.cudnn.allow_tf32 = False


def timeit(fn, n=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make(N, H, W, C, k):
    g = torch.Generator(device="cuda").manual_seed(3 + C + 7 * H + k)
    x = (torch.randn(N, H, W, C, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
    dz = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
    sc = (torch.rand(C, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(C, device="cuda", generator=g) * 0.3).float()
    return x, dz, w, sc, sh


def ops(N, H, W, C, k, t, act=True):
    x, dz, w, sc, sh = t
    z = torch.full_like(x, float("nan")); dx = torch.full_like(x, float("nan")); dw = torch.zeros_like(w)
    st = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    a_sc, a_sh = (P(sc), P(sh)) if act else (None, None)
    fwd = lambda: L.call("mnb_dw_fwd", P(x), a_sc, a_sh, P(w), None, P(z), P(st), N, H, W, C, k, 1, S())
    dgr = lambda: L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())
    wgr = lambda: L.call("mnb_dw_wgrad", P(x), a_sc, a_sh, P(dz), P(dw), N, H, W, C, k, 1, S())
    return z, st, dx, dw, fwd, dgr, wgr


def parity(N, H, W, C, k, act=True):
    t = make(N, H, W, C, k)
    x, dz, w, sc, sh = t
    z, st, dx, dw, fwd, dgr, wgr = ops(N, H, W, C, k, t, act)
    fwd(); dgr(); wgr(); torch.cuda.synchronize()
    A = x.float()
    if act: A = torch.relu(A * sc + sh).to(BF).float()
    A = A.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wb = w.to(BF).float().requires_grad_(True)
    zr = F.conv2d(A, wb, None, stride=1, padding=k // 2, groups=C)
    gA, gw = torch.autograd.grad(zr, [A, wb], dz.float().permute(0, 3, 1, 2))
    zs = z.double()
    out = {"shape": f"{N}x{H}x{W}x{C} k{k}", "act": act,
           "nan": int(torch.isnan(z.float()).sum().item() + torch.isnan(dx.float()).sum().item()),
           "fwd": rel(z.float().permute(0, 3, 1, 2), zr), "dgrad": rel(dx.float().permute(0, 3, 1, 2), gA), "wgrad": rel(dw, gw),
           "stats_sum": rel(st[:C], zs.sum(dim=(0, 1, 2))), "stats_sq": rel(st[C:], (zs * zs).sum(dim=(0, 1, 2)))}
    out["ok"] = out["nan"] == 0 and max(out["fwd"], out["dgrad"], out["wgrad"]) < 1e-2 and max(out["stats_sum"], out["stats_sq"]) < 1e-5
    print(json.dumps(out), flush=True)
    return out["ok"]


def timing(H, C, k, N=256, on=1):
    t = make(N, H, H, C, k)
    row = {"shape": f"{N}x{H}x{H}x{C} k{k}"}
    mb = 2 * N * H * H * C * 2 / 1e6
    for small, name in ((on, "small"), (0, "stream")):
        L.set_option("dw_small", small)
        z, st, dx, dw, fwd, dgr, wgr = ops(N, H, H, C, k, t)
        for kk, fn in (("fwd", fwd), ("dgrad", dgr), ("wgrad", wgr)):
            us = timeit(fn)
            row[f"{kk}_{name}_us"] = round(us, 1)
            if small: row[f"{kk}_frac_hbm"] = round(mb / us * 1e3 / 6546.6, 3)
    L.set_option("dw_small", 1)
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "time"]
    ok = True
    if "parity" in what:
        L.set_option("dw_small", 2)
        for (N, H, W, C, k) in [(3, 14, 14, 48, 5), (3, 14, 14, 48, 3), (2, 28, 28, 72, 5), (2, 28, 28, 40, 3), (2, 7, 7, 96, 5),
                                (2, 7, 7, 48, 3), (2, 16, 24, 32, 5), (2, 12, 16, 56, 3), (2, 24, 32, 24, 5), (3, 4, 4, 80, 5),
                                (2, 28, 20, 240, 5), (2, 14, 14, 576, 5), (2, 14, 14, 480, 3), (2, 20, 20, 88, 3), (1, 6, 8, 1152, 3)]:
            ok &= parity(N, H, W, C, k)
        ok &= parity(2, 14, 14, 48, 5, act=False)
        ok &= parity(2, 28, 28, 24, 3, act=False)
        L.set_option("dw_small", 1)
        print("DW_SMALL PARITY", "PASS" if ok else "FAIL")
    if "time" in what:
        for (H, C, k) in [(28, 240, 5), (14, 576, 5), (14, 480, 3), (28, 72, 5)]:
            timing(H, C, k)
        for (H, C, k) in [(7, 1152, 5), (7, 1152, 3)]:
            timing(H, C, k, on=2)


def bwd_timing(H, C, k, N=256):
    """fused mnb_dw_bwd_fused (dw_small) vs the unfused chain bn_bwd_apply_fused + dw_dgrad + dw_wgrad + bn_bwd_reduce(producer)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    M = N * H * H
    x = torch.randn(N, H, H, C, device="cuda", generator=g).to(BF)
    z = (torch.randn(N, H, H, C, device="cuda", generator=g) * 0.7 + 0.2).to(BF)
    dA = torch.randn(N, H, H, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
    sc, isc = ((torch.rand(C, device="cuda", generator=g) + 0.5).float() for _ in range(2))
    sh, ish = ((torch.randn(C, device="cuda", generator=g) * 0.3).float() for _ in range(2))
    mean = z.float().mean(dim=(0, 1, 2)).contiguous(); inv = (1 / torch.sqrt(z.float().var(dim=(0, 1, 2), unbiased=False) + 1e-5)).contiguous()
    sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, C, 1, S())
    dz = torch.empty_like(z); dx = torch.empty_like(x); dw = torch.zeros_like(w)
    ns = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dga, dbe, dbi = (torch.zeros(C, device="cuda") for _ in range(3))
    def unfused():
        L.call("mnb_bn_bwd_apply_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(inv), P(dga), P(dbe), P(dbi), P(dz), M, C, float(M), 1, S())
        L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, H, C, k, 1, S())
        L.call("mnb_dw_wgrad", P(x), P(isc), P(ish), P(dz), P(dw), N, H, H, C, k, 1, S())
        L.call("mnb_bn_bwd_reduce", P(dx), P(x), P(isc), P(ish), P(ns), M, C, 1, S())
    def fused():
        L.call("mnb_dw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(inv), P(dga), P(dbe), P(dbi), P(x), P(isc), P(ish),
               P(w), P(dx), P(dw), P(ns), N, H, H, C, k, float(M), 1, S())
    row = {"shape": f"{N}x{H}x{H}x{C} k{k}", "unfused_us": round(timeit(unfused), 1), "fused_small_us": round(timeit(fused), 1)}
    L.set_option("dw_small", 0)
    row["fused_stream_us"] = round(timeit(fused), 1)
    L.set_option("dw_small", 1)
    row["fused_frac_hbm"] = round(4 * M * C * 2 / row["fused_small_us"] / 1e3 / 6546.6 * 1e3 / 1e3, 3)
    print(json.dumps(row), flush=True)


if __name__ == "__main__" and "bwd" in sys.argv[1:]:
    for (H, C, k) in [(28, 240, 5), (14, 576, 5), (14, 480, 3), (28, 72, 5)]:
        bwd_timing(H, C, k)
