"""How much of a forward kernel's time is the per-CTA fp64 flush of the BN statistics?  Same launch with / without stats."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from mnb200 import _lib as L
BF = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
N = 256
g = torch.Generator(device="cuda").manual_seed(1)
for (H, Cin, Cout, k, dw) in [(112, 16, 48, 1, 0), (112, 48, 16, 1, 0), (56, 24, 72, 1, 0), (28, 40, 240, 1, 0), (14, 96, 576, 1, 0), (14, 576, 96, 1, 0),
                              (112, 48, 48, 3, 1), (56, 72, 72, 5, 1), (28, 240, 240, 5, 1), (14, 576, 576, 5, 1)]:
    x = torch.randn(N, H, H, Cin, device="cuda", generator=g).to(BF)
    sc = torch.rand(Cin, device="cuda") + 0.5; sh = torch.randn(Cin, device="cuda") * 0.3
    z = torch.empty(N, H, H, Cout, device="cuda", dtype=BF)
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    if dw:
        w = torch.randn(Cout, 1, k, k, device="cuda") / k
        f = lambda s_: L.call("mnb_dw_fwd", P(x), P(sc), P(sh), P(w), None, P(z), s_, N, H, H, Cin, k, 1, S())
    else:
        w = torch.randn(Cout, Cin, 1, 1, device="cuda") / Cin ** 0.5
        f = lambda s_: L.call("mnb_conv_fwd", P(x), P(sc), P(sh), P(w), None, P(z), s_, N, H, H, Cin, Cout, 1, 1, 0, 1, 0, 0, S())
    a = timeit(lambda: f(P(st))); b = timeit(lambda: f(None))
    print(json.dumps({"layer": f"{H}x{H} {Cin}->{Cout} k{k}{' dw' if dw else ''}", "with_stats_us": round(a, 1), "no_stats_us": round(b, 1), "delta_us": round(a - b, 1)}), flush=True)
