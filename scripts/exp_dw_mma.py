"""GPU experiment for the TMA + mma.sync depthwise kernels (csrc/dw_mma.cu, dw_mma_bwd.cu; option "dw_mma"): parity
against torch fp32 math on the same bf16 operands and against the shared-memory tile kernels, on odd / ragged shapes and
on every depthwise layer shape of MNASNet-224 at batch 256 (plus the 112^2 and 192x256 configurations), then stand-alone
timings.  Writes gpurun_out/exp_dw_mma.json incrementally.

    python scripts/exp_dw_mma.py [parity] [time] [bwd]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "exp_dw_mma.json")
os.makedirs(os.path.dirname(OUT), exist_ok=True)
T0 = time.time()
RES = {"items": []}


def flush():
    RES["elapsed_s"] = round(time.time() - T0, 1)
    with open(OUT, "w") as f:
        json.dump(RES, f, indent=1)


def item(name, **kw):
    kw["name"] = name
    RES["items"].append(kw)
    print(json.dumps(kw), flush=True)
    flush()


flush()
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from mnb200 import _lib as L  # noqa: E402

BF = torch.bfloat16
dev = "cuda"
PEAK = 6546.6


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


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make(N, H, W, C, k, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed + H * 7 + C)
    x = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)
    dz = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device=dev, generator=g) / k).float()
    sc = (torch.rand(C, device=dev, generator=g) + 0.5).float()
    sh = (torch.randn(C, device=dev, generator=g) * 0.3).float()
    return x, dz, w, sc, sh


def fwd_case(N, H, W, C, k, xform=True, timing=False):
    x, dz, w, sc, sh = make(N, H, W, C, k)
    out = {"shape": f"{N}x{H}x{W}x{C} k{k}", "xform": xform}
    res = {}
    for tag, opt in (("tile", 0), ("mma", 2)):
        L.set_option("dw_mma", opt)
        z = torch.full_like(x, float("nan"))
        st = torch.zeros(2 * C, device=dev, dtype=torch.float64)
        dx = torch.full_like(x, float("nan"))

        def f_fwd(z=z, st=st):
            L.call("mnb_dw_fwd", P(x), P(sc) if xform else None, P(sh) if xform else None, P(w), None, P(z), P(st), N, H, W,
                   C, k, 1, S())

        def f_dg(dx=dx):
            L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())
        dwt = torch.zeros(C, 1, k, k, device=dev)

        def f_wg(dwt=dwt):
            L.call("mnb_dw_wgrad", P(x), P(sc) if xform else None, P(sh) if xform else None, P(dz), P(dwt), N, H, W, C, k, 1,
                   S())
        f_fwd(); f_dg(); f_wg()
        torch.cuda.synchronize()
        res[tag] = (z.clone(), st.clone(), dx.clone(), dwt.clone())
        if timing:
            out[f"fwd_us_{tag}"] = round(timeit(f_fwd), 1)
            out[f"dgrad_us_{tag}"] = round(timeit(f_dg), 1)
            out[f"wgrad_us_{tag}"] = round(timeit(f_wg), 1)
    L.set_option("dw_mma", 1)
    if timing:
        by = 2 * x.numel() * 2
        out["fwd_frac_hbm_mma"] = round(by / out["fwd_us_mma"] / 1e3 / PEAK, 3)
        out["dgrad_frac_hbm_mma"] = round(by / out["dgrad_us_mma"] / 1e3 / PEAK, 3)
    # torch fp32 math on the same bf16 operands (weights rounded to bf16 like the tensor-pipe operand)
    if N * H * W * C <= 64 * 1024 * 1024:
        a = x.float().permute(0, 3, 1, 2)
        if xform:
            a = torch.relu(a * sc[None, :, None, None] + sh[None, :, None, None]).to(BF).float()
        wb = w.to(BF).float()
        zr = F.conv2d(a, wb, None, padding=k // 2, groups=C)
        dxr = F.conv_transpose2d(dz.float().permute(0, 3, 1, 2), wb, None, padding=k // 2, groups=C)
        z1, st1, dx1, dw1 = res["mma"]
        zq = z1.float().permute(0, 3, 1, 2)
        dwr = torch.nn.grad.conv2d_weight(a, w.shape, dz.float().permute(0, 3, 1, 2), padding=k // 2, groups=C)
        out["rel_wgrad_vs_torch"] = float(f"{rel(dw1, dwr):.2e}")
        out["rel_fwd_vs_torch"] = float(f"{rel(zq, zr):.2e}")
        out["rel_dgrad_vs_torch"] = float(f"{rel(dx1.float().permute(0, 3, 1, 2), dxr):.2e}")
        out["rel_stats_sum"] = float(f"{rel(st1[:C], zq.double().sum(dim=(0, 2, 3))):.2e}")
        out["rel_stats_sq"] = float(f"{rel(st1[C:], (zq.double() ** 2).sum(dim=(0, 2, 3))):.2e}")
    z0, st0, dx0, dw0 = res["tile"]
    z1, st1, dx1, dw1 = res["mma"]
    out["rel_wgrad_vs_tile"] = float(f"{rel(dw1, dw0):.2e}")
    out["nan"] = int(torch.isnan(z1.float()).sum().item() + torch.isnan(dx1.float()).sum().item())
    out["rel_fwd_vs_tile"] = float(f"{rel(z1.float(), z0.float()):.2e}")
    out["rel_dgrad_vs_tile"] = float(f"{rel(dx1.float(), dx0.float()):.2e}")
    out["rel_stats_vs_tile"] = float(f"{rel(st1, st0):.2e}")
    ok = out["nan"] == 0 and out["rel_fwd_vs_tile"] < 8e-3 and out["rel_dgrad_vs_tile"] < 8e-3 and out["rel_stats_vs_tile"] < 5e-3
    ok = ok and out["rel_wgrad_vs_tile"] < 5e-3 and not bool(torch.isnan(dw1).any())
    for key in ("rel_fwd_vs_torch", "rel_dgrad_vs_torch", "rel_wgrad_vs_torch"):
        if key in out:
            ok = ok and out[key] < 5e-3
    for key in ("rel_stats_sum", "rel_stats_sq"):
        if key in out:
            ok = ok and out[key] < 1e-4
    out["ok"] = bool(ok)
    return out


def bwd_case(N, H, W, C, k, timing=False, act=True):
    """Fused depthwise ConvBlock backward vs the unfused chain (bn_bwd_reduce -> bn_bwd_apply_fused -> dw_dgrad + dw_wgrad
    -> bn_bwd_reduce of the producing block), all through the C ABI on the same bf16 operands."""
    g = torch.Generator(device=dev).manual_seed(3 + H * 7 + C)
    x = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)            # raw output of the producing block
    z = (torch.randn(N, H, W, C, device=dev, generator=g) * 0.7 + 0.2).to(BF)   # this block's raw conv output
    dA = torch.randn(N, H, W, C, device=dev, generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device=dev, generator=g) / k).float()
    sc = (torch.rand(C, device=dev, generator=g) + 0.5).float()
    sh = (torch.randn(C, device=dev, generator=g) * 0.3).float()
    isc = (torch.rand(C, device=dev, generator=g) + 0.5).float()
    ish = (torch.randn(C, device=dev, generator=g) * 0.3).float()
    zf = z.float()
    mean = zf.mean(dim=(0, 1, 2)).contiguous()
    invstd = (1.0 / torch.sqrt(zf.var(dim=(0, 1, 2), unbiased=False) + 1e-5)).contiguous()
    M = N * H * W
    out = {"shape": f"{N}x{H}x{W}x{C} k{k}"}
    # ---- unfused chain (the round-1 path) ----
    L.set_option("dw_mma", 0)
    sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, C, 1, S())
    dz = torch.empty_like(z)
    dga0, dbe0, dbi0 = (torch.zeros(C, device=dev) for _ in range(3))
    dx0 = torch.full_like(x, float("nan"))
    dw0 = torch.zeros(C, 1, k, k, device=dev)
    ns0 = torch.zeros(2 * C, device=dev, dtype=torch.float64)

    def unfused():
        L.call("mnb_bn_bwd_apply_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(invstd), P(dga0), P(dbe0), P(dbi0),
               P(dz), M, C, float(M), 1, S())
        L.call("mnb_dw_dgrad", P(dz), P(w), P(dx0), None, None, None, None, N, H, W, C, k, 1, S())
        L.call("mnb_dw_wgrad", P(x), P(isc) if act else None, P(ish) if act else None, P(dz), P(dw0), N, H, W, C, k, 1, S())
        if act:
            L.call("mnb_bn_bwd_reduce", P(dx0), P(x), P(isc), P(ish), P(ns0), M, C, 1, S())
    unfused()
    torch.cuda.synchronize()
    ref = [t.clone() for t in (dx0, dw0, ns0, dga0, dbe0)]
    L.set_option("dw_mma", 1)
    # ---- fused ----
    dx1 = torch.full_like(x, float("nan"))
    dw1 = torch.zeros(C, 1, k, k, device=dev)
    ns1 = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    dga1, dbe1, dbi1 = (torch.zeros(C, device=dev) for _ in range(3))

    def fused():
        L.call("mnb_dw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(invstd), P(dga1), P(dbe1), P(dbi1), P(x),
               P(isc) if act else None, P(ish) if act else None, P(w), P(dx1), P(dw1), P(ns1) if act else None, N, H, W, C, k,
               float(M), 1, S())
    fused()
    torch.cuda.synchronize()
    out["nan"] = int(torch.isnan(dx1.float()).sum().item() + torch.isnan(dw1).sum().item())
    out["rel_dx"] = float(f"{rel(dx1.float(), ref[0].float()):.2e}")
    out["rel_dw"] = float(f"{rel(dw1, ref[1]):.2e}")
    out["rel_nsums"] = float(f"{rel(ns1, ref[2]):.2e}") if act else 0.0
    out["rel_dgamma"] = float(f"{rel(dga1, ref[3]):.2e}")
    out["rel_dbeta"] = float(f"{rel(dbe1, ref[4]):.2e}")
    # torch fp32 math on the same operands: dZ in fp32 (the fused kernel rounds dZ to bf16 once, like the unfused chain)
    if N * H * W * C <= 32 * 1024 * 1024:
        sg, sgz = sums[:C].float(), sums[C:].float()
        dga = invstd * (sgz - mean * sg)
        b = -sc * invstd * dga / M
        c3 = -sc * sg / M - b * mean
        Gm = dA.float() * ((zf * sc + sh) > 0)
        dzr = (sc * Gm + b * zf + c3).to(BF).float().permute(0, 3, 1, 2)
        wb = w.to(BF).float()
        a = x.float().permute(0, 3, 1, 2)
        if act:
            a = torch.relu(a * isc[None, :, None, None] + ish[None, :, None, None]).to(BF).float()
        dxr = F.conv_transpose2d(dzr, wb, None, padding=k // 2, groups=C)
        ar = a.clone().requires_grad_(False)
        dwr = torch.nn.grad.conv2d_weight(ar, w.shape, dzr, padding=k // 2, groups=C)
        out["rel_dx_torch"] = float(f"{rel(dx1.float().permute(0, 3, 1, 2), dxr):.2e}")
        out["rel_dw_torch"] = float(f"{rel(dw1, dwr):.2e}")
    ok = out["nan"] == 0 and out["rel_dx"] < 8e-3 and out["rel_dw"] < 8e-3 and out["rel_nsums"] < 8e-3 and \
        out["rel_dgamma"] < 1e-5 and out["rel_dbeta"] < 1e-5
    for key in ("rel_dx_torch", "rel_dw_torch"):
        if key in out:
            ok = ok and out[key] < 6e-3
    out["ok"] = bool(ok)
    if timing:
        out["us_unfused"] = round(timeit(unfused), 1)
        out["us_fused"] = round(timeit(fused), 1)
        by = 4 * x.numel() * 2
        out["fused_frac_hbm"] = round(by / out["us_fused"] / 1e3 / PEAK, 3)
    return out


SMALL = [(2, 12, 10, 32, 3), (2, 9, 11, 72, 5), (3, 7, 7, 48, 3), (2, 4, 4, 240, 5), (1, 17, 5, 1152, 3), (2, 14, 14, 576, 5),
         (2, 30, 20, 16, 5), (4, 56, 56, 72, 5), (1, 37, 45, 8, 3), (2, 25, 33, 40, 5), (3, 13, 50, 56, 3), (1, 64, 64, 24, 5),
         (2, 6, 8, 1152, 3), (2, 12, 16, 576, 5), (1, 128, 96, 32, 3), (1, 96, 128, 48, 3)]
LAYERS = [(256, 112, 112, 32, 3), (256, 112, 112, 48, 3), (256, 56, 56, 72, 5), (256, 28, 28, 240, 5), (256, 14, 14, 480, 3),
          (256, 14, 14, 576, 5), (256, 7, 7, 1152, 3)]
OTHER = [(512, 56, 56, 48, 3), (512, 28, 28, 72, 5), (512, 4, 4, 1152, 3), (256, 96, 128, 48, 3), (256, 48, 64, 72, 5),
         (256, 6, 8, 1152, 3)]


def main():
    what = sys.argv[1:] or ["parity", "time"]
    allok = True
    if "parity" in what:
        for c in SMALL:
            for xf in (True, False):
                try:
                    r = fwd_case(*c, xform=xf)
                except Exception as e:          # noqa: BLE001
                    r = {"shape": str(c), "ok": False, "error": repr(e)[:300]}
                item("parity", **r)
                allok = allok and r["ok"]
                if "error" in r:
                    item("abort", reason="kernel error; context unusable")
                    print("EXP_DW_MMA FAIL")
                    return 1
    if "bwd" in what:
        for c in SMALL:
            for act in (True, False):
                try:
                    r = bwd_case(*c, act=act)
                except Exception as e:          # noqa: BLE001
                    r = {"shape": str(c), "ok": False, "error": repr(e)[:300]}
                r["act"] = act
                item("bwd_parity", **r)
                allok = allok and r["ok"]
                if "error" in r:
                    print("EXP_DW_MMA FAIL")
                    return 1
    if "bwdtime" in what:
        for c in LAYERS + OTHER:
            if c[1] < 12:
                continue
            try:
                r = bwd_case(*c, timing=True)
            except Exception as e:          # noqa: BLE001
                r = {"shape": str(c), "ok": False, "error": repr(e)[:300]}
            item("bwd_layer", **r)
            allok = allok and r["ok"]
            if "error" in r:
                break
    if "ncubwd" in what:
        for idx in [int(a) for a in what if a.isdigit()]:
            r = bwd_case(*(LAYERS + OTHER)[idx])
            print(r)
        return 0
    if "ncu" in what:              # a few launches of selected layers, for an ncu capture
        for idx in [int(a) for a in what if a.isdigit()]:
            c = (LAYERS + OTHER)[idx]
            x, dz, w, sc, sh = make(*c)
            N, H, W, C, k = c
            z = torch.empty_like(x)
            st = torch.zeros(2 * C, device=dev, dtype=torch.float64)
            for _ in range(2):
                L.call("mnb_dw_fwd", P(x), P(sc), P(sh), P(w), None, P(z), P(st), N, H, W, C, k, 1, S())
            torch.cuda.synchronize()
        return 0
    if "time" in what:
        for c in LAYERS + OTHER:
            try:
                r = fwd_case(*c, xform=True, timing=True)
            except Exception as e:          # noqa: BLE001
                r = {"shape": str(c), "ok": False, "error": repr(e)[:300]}
            item("layer", **r)
            allok = allok and r["ok"]
            if "error" in r:
                break
    print("EXP_DW_MMA", "PASS" if allok else "FAIL")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
