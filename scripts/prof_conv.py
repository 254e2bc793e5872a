"""Run one conv layer of the step in isolation (for ncu captures).  usage: prof_conv.py MODE N H W Cin Cout k stride [xf]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mnasnet-pytorch_b200"))
import torch
from mnb200 import _lib as L
mode, N, H, W, Cin, Cout, k, stride = sys.argv[1], *map(int, sys.argv[2:9])
xf = len(sys.argv) > 9 and sys.argv[9] == "xf"
pad = k // 2
Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
dev = "cuda"
x = torch.randn(N, H, W, Cin, device=dev).bfloat16()
z = torch.empty(N, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
dz = torch.randn(N, Ho, Wo, Cout, device=dev).bfloat16()
dx = torch.empty_like(x)
w = torch.randn(Cout, Cin, k, k, device=dev) * 0.1
b = torch.zeros(Cout, device=dev)
s = torch.rand(Cin, device=dev) + 0.5 if xf else None
t = torch.randn(Cin, device=dev) * 0.1 if xf else None
stats = torch.zeros(2 * Cout, device=dev, dtype=torch.float64)
if os.environ.get("NOSTATS"): stats = None
dw = torch.zeros_like(w)
P = lambda q: None if q is None else q.data_ptr()
st = torch.cuda.current_stream().cuda_stream
reps = int(os.environ.get("REPS", "4"))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(reps + 3):
    if i == 3:
        ev[0].record()
    if mode == "fwd":
        L.call("mnb_conv_fwd", P(x), P(s), P(t), P(w), P(b), P(z), P(stats), N, H, W, Cin, Cout, k, stride, pad, 1, 0, 0, st)
    elif mode == "dgrad":
        L.call("mnb_conv_dgrad", P(dz), P(w), None, P(dx), None, None, None, None, N, H, W, Cin, Cout, k, stride, pad, 1, 0, st)
    elif mode == "wgrad":
        L.call("mnb_conv_wgrad", P(x), P(s), P(t), P(dz), P(dw), N, H, W, Cin, Cout, k, stride, pad, 1, 0, 0, st)
    elif mode.startswith("dw"):
        kk = int(mode[2])
        if mode.endswith("fwd"):
            L.call("mnb_dw_fwd", P(x), P(s), P(t), P(w), P(b), P(z), P(stats), N, H, W, Cin, kk, 1, st)
        elif mode.endswith("dgrad"):
            L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, Cin, kk, 1, st)
        else:
            L.call("mnb_dw_wgrad", P(x), P(s), P(t), P(dz), P(dw), N, H, W, Cin, kk, 1, st)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
gb = (x.numel() + z.numel()) * 2 / 1e9
print(f"{mode} {N}x{H}x{W} {Cin}->{Cout} k{k}s{stride}: {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s (x+z bytes)")
