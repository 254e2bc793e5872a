"""Stock-PyTorch GPU throughput of the reference architecture (oracle modules = same torch ops as the reference)
on the same B200: fp32 NCHW and bf16-autocast channels_last, batch 256, 224^2.  Context for DESIGN.md only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import mnasnet_oracle as O

def run(mode, n=256, steps=8, warm=3):
    torch.manual_seed(42)
    sd = O.init_state_dict()
    sd = {k: v.cuda() for k, v in sd.items()}
    # keep aliasing of shared blocks
    tr = O.Trainer(sd)
    x = torch.randn(n, 3, 224, 224, device="cuda")
    t = torch.randint(0, 1000, (n,), device="cuda")
    if mode == "bf16_cl":
        x = x.contiguous(memory_format=torch.channels_last)
    def step():
        if mode == "bf16_cl":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return tr.step(x, t, dropout_masks=None)
        return tr.step(x, t, dropout_masks=None)
    for _ in range(warm): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    print(f"torch-eager {mode}: {dt*1e3:.1f} ms/step  {n/dt:.0f} img/s")

if __name__ == "__main__":
    torch.backends.cudnn.benchmark = True
    for m in ("fp32_nchw", "bf16_cl"):
        try: run(m)
        except Exception as e: print(m, "failed:", repr(e)[:200])
