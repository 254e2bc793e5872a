"""One launch set of the dw_small kernels at a given shape (for ncu): python scripts/exp_dw_small_one.py H C k"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from exp_dw_small import make, ops
H, C, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
t = make(256, H, H, C, k)
z, st, dx, dw, fwd, dgr, wgr = ops(256, H, H, C, k, t)
for _ in range(2):
    fwd(); dgr(); wgr()
torch.cuda.synchronize()
