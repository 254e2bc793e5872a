"""GPU experiment for the TMA + mma.sync 3x3 kernels (csrc/c3_mma.cu): parity against torch fp32 math on the same bf16
operands and stand-alone timings at batch 256 against the tcgen05 pipeline (impl 2).
   python scripts/exp_c3.py [parity] [time]"""
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
torch.backends.cuda.matmul.allow_tf32 = False


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


def make(N, H, W, Cin, Cout, stride, seed=3):
    g = torch.Generator(device="cuda").manual_seed(seed + Cin + 7 * Cout + H)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    x = (torch.randn(N, H, W, Cin, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
    dz = torch.randn(N, Ho, Wo, Cout, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)).float()
    sc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
    wpf = torch.empty(w.numel(), device="cuda", dtype=BF); wpd = torch.empty(w.numel(), device="cuda", dtype=BF)
    L.call("mnb_pack_weights", P(w), P(wpf), P(wpd), Cout, Cin, 3, S())
    return x, dz, w, sc, sh, wpf, wpd, Ho, Wo


def run(N, H, W, Cin, Cout, stride, impl, t, packed=True, act=True):
    x, dz, w, sc, sh, wpf, wpd, Ho, Wo = t
    z = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda", dtype=BF)
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    dx = torch.full((N, H, W, Cin), float("nan"), device="cuda", dtype=BF)
    dw = torch.zeros_like(w)
    a_sc, a_sh = (P(sc), P(sh)) if act else (None, None)
    fwd = lambda: L.call("mnb_conv_fwd_packed", P(x), a_sc, a_sh, P(w), P(wpf) if packed else None, None, P(z), P(st), N, H, W, Cin, Cout, 3, stride, 1, 1, 0, impl, S())
    dgr = lambda: L.call("mnb_conv_dgrad_packed", P(dz), P(w), P(wpd) if packed else None, None, P(dx), None, None, None, None, N, H, W, Cin, Cout, 3, stride, 1, 1, impl, S())
    wgr = lambda: L.call("mnb_conv_wgrad", P(x), a_sc, a_sh, P(dz), P(dw), N, H, W, Cin, Cout, 3, stride, 1, 1, 0, impl, S())
    return z, st, dx, dw, fwd, dgr, wgr


def parity(N, H, W, Cin, Cout, stride, packed=True, act=True):
    t = make(N, H, W, Cin, Cout, stride)
    x, dz, w, sc, sh, *_ = t
    z, st, dx, dw, fwd, dgr, wgr = run(N, H, W, Cin, Cout, stride, 0, t, packed, act)
    fwd(); dgr(); wgr(); torch.cuda.synchronize()
    A = x.float()
    if act: A = torch.relu(A * sc + sh).to(BF).float()
    A = A.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wb = w.to(BF).float().requires_grad_(True)
    zr = F.conv2d(A, wb, None, stride=stride, padding=1)
    gA, gw = torch.autograd.grad(zr, [A, wb], dz.float().permute(0, 3, 1, 2))
    zs = z.double()
    out = {"shape": f"{N}x{H}x{W} {Cin}->{Cout} s{stride}", "packed": packed, "act": act,
           "nan": int(torch.isnan(z.float()).sum().item() + torch.isnan(dx.float()).sum().item()),
           "fwd": rel(z.float().permute(0, 3, 1, 2), zr), "dgrad": rel(dx.float().permute(0, 3, 1, 2), gA), "wgrad": rel(dw, gw),
           "stats_sum": rel(st[:Cout], zs.sum(dim=(0, 1, 2))), "stats_sq": rel(st[Cout:], (zs * zs).sum(dim=(0, 1, 2)))}
    out["ok"] = out["nan"] == 0 and max(out["fwd"], out["dgrad"], out["wgrad"]) < 1e-2 and max(out["stats_sum"], out["stats_sq"]) < 1e-5
    print(json.dumps(out), flush=True)
    return out["ok"]


def timing(H, Cin, Cout, stride, N=256):
    t = make(N, H, H, Cin, Cout, stride)
    row = {"shape": f"{N}x{H}x{H} {Cin}->{Cout} s{stride}"}
    Ho = (H - 1) // stride + 1
    mb = (N * H * H * Cin + N * Ho * Ho * Cout) * 2 / 1e6
    for impl, name in ((0, "c3"), (2, "tc")):
        z, st, dx, dw, fwd, dgr, wgr = run(N, H, H, Cin, Cout, stride, impl, t)
        for k, fn in (("fwd", fwd), ("dgrad", dgr), ("wgrad", wgr)):
            us = timeit(fn)
            row[f"{k}_{name}_us"] = round(us, 1)
            if impl == 0: row[f"{k}_GBps"] = round(mb / us * 1e3)
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "time"]
    ok = True
    if "parity" in what:
        for (N, H, W, ci, co, s) in [(3, 112, 112, 16, 24, 2), (3, 56, 56, 24, 40, 2), (3, 28, 28, 40, 80, 2), (2, 30, 28, 16, 24, 2),
                                     (2, 13, 12, 24, 40, 2), (2, 64, 96, 16, 24, 2), (5, 6, 6, 40, 80, 2), (2, 48, 64, 24, 40, 2)]:
            ok &= parity(N, H, W, ci, co, s)
        ok &= parity(2, 56, 56, 16, 24, 2, packed=False, act=False)
        ok &= parity(2, 28, 28, 40, 80, 2, packed=False, act=True)
        print("C3 PARITY", "PASS" if ok else "FAIL")
    if "time" in what:
        for (H, ci, co, s) in [(112, 16, 24, 2), (56, 24, 40, 2), (28, 40, 80, 2)]:
            timing(H, ci, co, s)
