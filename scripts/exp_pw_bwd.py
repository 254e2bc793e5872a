"""GPU experiment for the fused pointwise ConvBlock backward (csrc/pw_bwd_fused.cu): parity against the unfused chain
(bn_bwd_reduce -> bn_bwd_apply_fused -> conv_dgrad + conv_wgrad -> bn_bwd_reduce of the producer) and torch fp32 math, and
stand-alone timings at batch 256.   python scripts/exp_pw_bwd.py [parity] [time]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from mnb200 import _lib as L
BF = torch.bfloat16
PEAK = 6546.6
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream
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

def case(N, H, W, Cin, Cout, act=True, add=True, timing=False):
    M = N * H * W
    g = torch.Generator(device="cuda").manual_seed(7 + Cin * 3 + Cout)
    x = torch.randn(M, Cin, device="cuda", generator=g).to(BF)
    z = (torch.randn(M, Cout, device="cuda", generator=g) * 0.7 + 0.2).to(BF)
    dA = torch.randn(M, Cout, device="cuda", generator=g).to(BF)
    sk = torch.randn(M, Cin, device="cuda", generator=g).to(BF) if add else None
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5).float()
    sc, isc = ((torch.rand(c, device="cuda", generator=g) + 0.5).float() for c in (Cout, Cin))
    sh, ish = ((torch.randn(c, device="cuda", generator=g) * 0.3).float() for c in (Cout, Cin))
    zf = z.float()
    mean = zf.mean(0).contiguous(); invstd = (1.0 / torch.sqrt(zf.var(0, unbiased=False) + 1e-5)).contiguous()
    out = {"shape": f"{N}x{H}x{W} {Cin}->{Cout}", "act": act, "add": add}
    sums = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, Cout, 1, S())
    dz = torch.empty_like(z); dx0 = torch.full_like(x, float("nan")); dw0 = torch.zeros_like(w)
    ns0 = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    dga0, dbe0, dbi0 = (torch.zeros(Cout, device="cuda") for _ in range(3))
    wpk_d = torch.empty(w.numel(), device="cuda", dtype=BF); wpk_f = torch.empty(w.numel(), device="cuda", dtype=BF)
    L.call("mnb_pack_weights", P(w), P(wpk_f), P(wpk_d), Cout, Cin, 1, S())
    def unfused():
        L.call("mnb_bn_bwd_apply_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(invstd), P(dga0), P(dbe0), P(dbi0), P(dz), M, Cout, float(M), 1, S())
        L.call("mnb_conv_dgrad_packed", P(dz), P(w), P(wpk_d), P(sk), P(dx0), None, None, None, None, N, H, W, Cin, Cout, 1, 1, 0, 1, 0, S())
        L.call("mnb_conv_wgrad", P(x), P(isc) if act else None, P(ish) if act else None, P(dz), P(dw0), N, H, W, Cin, Cout, 1, 1, 0, 1, 0, 0, S())
        if act:
            L.call("mnb_bn_bwd_reduce", P(dx0), P(x), P(isc), P(ish), P(ns0), M, Cin, 1, S())
    unfused(); torch.cuda.synchronize()
    ref = [t.clone() for t in (dx0, dw0, ns0, dga0, dbe0)]
    dx1 = torch.full_like(x, float("nan")); dw1 = torch.zeros_like(w)
    ns1 = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    dga1, dbe1, dbi1 = (torch.zeros(Cout, device="cuda") for _ in range(3))
    def fused():
        L.call("mnb_pw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(invstd), P(dga1), P(dbe1), P(dbi1), P(x),
               P(isc) if act else None, P(ish) if act else None, P(w), P(sk), P(dx1), P(dw1), P(ns1) if act else None, M, Cin, Cout, float(M), 1, S())
    fused(); torch.cuda.synchronize()
    out["nan"] = int(torch.isnan(dx1.float()).sum().item() + torch.isnan(dw1).sum().item())
    out["rel_dx"] = float(f"{rel(dx1.float(), ref[0].float()):.2e}")
    out["rel_dw"] = float(f"{rel(dw1, ref[1]):.2e}")
    out["rel_nsums"] = float(f"{rel(ns1, ref[2]):.2e}") if act else 0.0
    out["rel_dgamma"] = float(f"{rel(dga1, ref[3]):.2e}"); out["rel_dbeta"] = float(f"{rel(dbe1, ref[4]):.2e}")
    if M * max(Cin, Cout) <= 64 * 1024 * 1024:
        sg, sgz = sums[:Cout].float(), sums[Cout:].float()
        dga = invstd * (sgz - mean * sg); b = -sc * invstd * dga / M; c3 = -sc * sg / M - b * mean
        Gm = dA.float() * ((zf * sc + sh) > 0)
        dzr = (sc * Gm + b * zf + c3).to(BF).float()
        wb = w.view(Cout, Cin).to(BF).float()
        a = x.float()
        if act: a = torch.relu(a * isc + ish).to(BF).float()
        dxr = dzr @ wb + (sk.float() if add else 0)
        dwr = dzr.t() @ a
        out["rel_dx_torch"] = float(f"{rel(dx1.float(), dxr):.2e}")
        out["rel_dw_torch"] = float(f"{rel(dw1.view(Cout, Cin), dwr):.2e}")
        if act:
            msk = (x.float() * isc + ish) > 0
            out["rel_nsums_self"] = float(f"{rel(ns1[:Cin], (dx1.double() * msk).sum(0)):.2e}")
    ok = out["nan"] == 0 and out["rel_dx"] < 8e-3 and out["rel_dw"] < 8e-3 and out["rel_nsums"] < 8e-3 and out["rel_dgamma"] < 1e-5
    for k in ("rel_dx_torch", "rel_dw_torch"):
        if k in out: ok = ok and out[k] < 6e-3
    if "rel_nsums_self" in out: ok = ok and out["rel_nsums_self"] < 1e-5
    out["ok"] = bool(ok)
    if timing:
        out["us_unfused"] = round(timeit(unfused), 1); out["us_fused"] = round(timeit(fused), 1)
        by = (2 * Cout + (3 if add else 2) * Cin) * M * 2
        out["fused_frac_hbm"] = round(by / out["us_fused"] / 1e3 / PEAK, 3)
    return out

SHAPES = [(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)]
def main():
    what = sys.argv[1:] or ["parity", "time"]
    allok = True
    if "parity" in what:
        for cin, cout in SHAPES:
            for (n, h, w) in ((2, 12, 10), (1, 37, 5), (3, 28, 28), (1, 1, 2)):
                for act, add in ((True, True), (False, False), (True, False)):
                    try: r = case(n, h, w, cin, cout, act, add)
                    except Exception as e: r = {"shape": f"{n}x{h}x{w} {cin}->{cout}", "ok": False, "error": repr(e)[:300]}
                    print(json.dumps(r), flush=True); allok = allok and r["ok"]
                    if "error" in r: print("EXP_PW_BWD FAIL"); return 1
    if "time" in what:
        for (hw, cin, cout) in ((112, 16, 48), (112, 48, 16), (112, 32, 16), (56, 24, 72), (56, 72, 24)):
            try: r = case(256, hw, hw, cin, cout, True, cin < cout, timing=True)
            except Exception as e: r = {"shape": f"{hw} {cin}->{cout}", "ok": False, "error": repr(e)[:300]}
            print(json.dumps(r), flush=True); allok = allok and r["ok"]
            if "error" in r: break
    if "slice" in what:
        for sl in (48, 80):
            L.set_option("pwb_slice", sl)
            try: r = case(256, 28, 28, 240, 40, True, False, timing=True)
            except Exception as e: r = {"shape": "28 240->40", "ok": False, "error": repr(e)[:300]}
            r["slice"] = sl
            print(json.dumps(r), flush=True)
        L.set_option("pwb_slice", 0)
    print("EXP_PW_BWD", "PASS" if allok else "FAIL")
    return 0 if allok else 1
if __name__ == "__main__" and "proj" not in sys.argv[1:]:
    sys.exit(main())


def proj_timing(H, Cin, Cout, add=True, N=256):
    """mnb_pw_proj_bwd vs conv_dgrad + conv_wgrad + bn_bwd_reduce(producer) on the same dZ."""
    M = N * H * H
    g = torch.Generator(device="cuda").manual_seed(3)
    dz = torch.randn(M, Cout, device="cuda", generator=g).to(BF)
    x = torch.randn(M, Cin, device="cuda", generator=g).to(BF)
    sk = torch.randn(M, Cin, device="cuda", generator=g).to(BF) if add else None
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5).float()
    isc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float(); ish = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
    dx = torch.empty_like(x); dw = torch.zeros_like(w); ns = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    wpk_d = torch.empty(w.numel(), device="cuda", dtype=BF); wpk_f = torch.empty(w.numel(), device="cuda", dtype=BF)
    L.call("mnb_pack_weights", P(w), P(wpk_f), P(wpk_d), Cout, Cin, 1, S())
    def unfused():
        L.call("mnb_conv_dgrad_packed", P(dz), P(w), P(wpk_d), P(sk), P(dx), None, None, None, None, N, H, H, Cin, Cout, 1, 1, 0, 1, 0, S())
        L.call("mnb_conv_wgrad", P(x), P(isc), P(ish), P(dz), P(dw), N, H, H, Cin, Cout, 1, 1, 0, 1, 0, 0, S())
        L.call("mnb_bn_bwd_reduce", P(dx), P(x), P(isc), P(ish), P(ns), M, Cin, 1, S())
    def fused():
        L.call("mnb_pw_proj_bwd", P(dz), P(x), P(isc), P(ish), P(w), P(sk), P(dx), P(dw), P(ns), M, Cin, Cout, 1, S())
    print(json.dumps({"shape": f"{N}x{H}x{H} {Cin}->{Cout}", "add": add, "unfused_us": round(timeit(unfused), 1), "proj_fused_us": round(timeit(fused), 1)}), flush=True)


if __name__ == "__main__" and "proj" in sys.argv[1:]:
    for (H, ci, co) in ((14, 576, 96), (28, 240, 40), (14, 480, 80)):
        proj_timing(H, ci, co)
