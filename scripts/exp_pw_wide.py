"""Stand-alone timing of the wide 1x1 forward (csrc/pw_wide_fwd.cu, impl 0) against the tcgen05 pipeline (impl 2), batch 256."""
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
L.set_option("pw_wide", 1)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(64, 1024, 1024, device="cuda")       # warm the clocks
for _ in range(20): x = x * 1.0001
for (H, Cin, Cout) in [(14, 576, 96), (28, 240, 40), (14, 480, 80), (14, 96, 576), (28, 40, 240), (14, 80, 480)]:
    x = torch.randn(N, H, H, Cin, device="cuda", generator=g).to(BF)
    sc = torch.rand(Cin, device="cuda") + 0.5; sh = torch.randn(Cin, device="cuda") * 0.3
    w = torch.randn(Cout, Cin, 1, 1, device="cuda") / Cin ** 0.5
    wpf = torch.empty(w.numel(), device="cuda", dtype=BF); wpd = torch.empty(w.numel(), device="cuda", dtype=BF)
    L.call("mnb_pack_weights", P(w), P(wpf), P(wpd), Cout, Cin, 1, S())
    z = torch.empty(N, H, H, Cout, device="cuda", dtype=BF); st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    f = lambda impl: L.call("mnb_conv_fwd_packed", P(x), P(sc), P(sh), P(w), P(wpf), None, P(z), P(st), N, H, H, Cin, Cout, 1, 1, 0, 1, 0, impl, S())
    a = timeit(lambda: f(0)); b = timeit(lambda: f(2))
    mb = N * H * H * (Cin + Cout) * 2 / 1e6
    print(json.dumps({"layer": f"{H}x{H} {Cin}->{Cout}", "pw_wide_us": round(a, 1), "tcgen05_us": round(b, 1), "pw_wide_GBps": round(mb / a * 1e3)}), flush=True)
