"""One launch of each c3_mma kernel at the 16->24 / 112x112 batch-256 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from exp_c3 import make, run
N, H, ci, co, s = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 112, int(sys.argv[2]) if len(sys.argv) > 2 else 16, int(sys.argv[3]) if len(sys.argv) > 3 else 24, 2
t = make(N, H, H, ci, co, s)
z, st, dx, dw, fwd, dgr, wgr = run(N, H, H, ci, co, s, 0, t)
for _ in range(2):
    fwd(); dgr(); wgr()
torch.cuda.synchronize()
