"""Timing of the BatchNorm-backward streaming kernels (bn_bwd_reduce / bn_bwd_apply_fused) on the step's tensor shapes,
sweeping the CTAs-per-SM launch parameter ("bn_ctas")."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from mnb200 import _lib as L
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
SHAPES = [(112, 48), (112, 16), (56, 72), (56, 24), (28, 240), (28, 40), (14, 576), (14, 96), (7, 1152)]
for H, C in SHAPES:
    M = 256 * H * H
    g = torch.Generator(device="cuda").manual_seed(1)
    dA = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    z = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    dz = torch.empty_like(z)
    sc = torch.rand(C, device="cuda") + 0.5; sh = torch.randn(C, device="cuda") * 0.3
    mean = torch.zeros(C, device="cuda"); inv = torch.ones(C, device="cuda")
    sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dg, db, dbi = (torch.zeros(C, device="cuda") for _ in range(3))
    row = {"shape": f"{H}x{H}x{C}", "MB": round(M * C * 2 / 1e6, 1)}
    for ctas in (1, 2, 3):
        L.set_option("bn_ctas", ctas)
        t = timeit(lambda: L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, C, 1, S()))
        row[f"reduce_us_{ctas}"] = round(t, 1); row[f"reduce_GBps_{ctas}"] = round(2 * M * C * 2 / t / 1e3)
    L.set_option("bn_ctas", 0)
    t = timeit(lambda: L.call("mnb_bn_bwd_apply_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(mean), P(inv), P(dg), P(db), P(dbi), P(dz), M, C, float(M), 1, S()))
    row["apply_us"] = round(t, 1); row["apply_GBps"] = round(3 * M * C * 2 / t / 1e3)
    t = timeit(lambda: dz.copy_(z))
    row["copy_us"] = round(t, 1); row["copy_GBps"] = round(2 * M * C * 2 / t / 1e3)
    print(json.dumps(row), flush=True)
