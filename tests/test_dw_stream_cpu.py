"""CPU: the depthwise row-stream kernel's per-lane program (csrc/dwconv_stream.cu) executed ON THE HOST, lane by lane
and warp by warp, against torch's depthwise convolution.  The lane program has no warp collectives, so the same source
compiles for the host (-DMNB_DW_STREAM_EMUL); this checks the ring-buffer arithmetic, the zero padding after the fused
BN+ReLU, the task decomposition (channel blocks x strips x row segments) and the flush, without a GPU.  The device
kernel itself is covered by the GPU tests."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "mnasnet-pytorch_b200", "csrc", "dwconv_stream.cu")
OUT = os.path.join(ROOT, "tests", "_build", "libdw_stream_emul.so")


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.fixture(scope="module")
def emul():
    nvcc = _nvcc()
    if nvcc is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC",
               "-DMNB_DW_STREAM_EMUL", "-shared", SRC, "-o", OUT]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    lib = ctypes.CDLL(OUT)
    fn = lib.mnb_emul_dw_stream
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 9 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    fn.set_pd = lib.mnb_emul_dw_stream_set_pd
    fn.set_pd.argtypes = [ctypes.c_int]
    fn.set_pd.restype = None
    fn.set_tw8 = lib.mnb_emul_dw_stream_set_tw8
    fn.set_tw8.argtypes = [ctypes.c_int]
    fn.set_tw8.restype = None
    return fn


def P(t):
    return None if t is None else t.data_ptr()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


CASES = [
    # N, H, W, C, k, xform, warps
    (2, 12, 10, 32, 3, True, 8),
    (1, 9, 11, 72, 5, True, 12),        # 3 strips x 9 pairs = 27 lanes, 4 channel blocks
    (2, 7, 7, 48, 3, False, 6),
    (1, 30, 9, 48, 5, True, 3),
    (1, 6, 5, 240, 5, True, 8),         # 30 pairs per lane group
    (1, 33, 17, 16, 3, True, 40),       # more warps than tasks -> short row segments, idle warps
    (3, 14, 14, 64, 5, False, 5),
    (1, 5, 6, 1152, 3, True, 36),
    (2, 2, 2, 1152, 3, True, 592),      # tiny maps of the progressive-resize shapes, full-size grid
    (2, 4, 4, 576, 5, True, 444),
    (1, 3, 4, 480, 3, False, 592),
    (1, 1, 1, 240, 5, True, 444),
    (2, 64, 48, 72, 5, True, 444),      # rectangular cluster crop at the 5x5 / 72-channel layer
    (1, 9, 70, 32, 3, True, 16),        # wide maps: interior (unpredicated) strips between the two border strips
    (2, 20, 100, 48, 5, True, 24),
    (1, 12, 66, 16, 3, False, 10),
]


@pytest.mark.parametrize("pd,tw8", [(1, 0), (2, 0), (3, 0), (1, 1), (3, 1)])
@pytest.mark.parametrize("case", CASES)
def test_lane_program_matches_torch(emul, case, pd, tw8):
    N, H, W, C, k, xform, warps = case
    if tw8 and k != 3:
        pytest.skip("8-column strips exist for 3x3 only")
    emul.set_pd(pd)                                    # rows of prefetch kept in flight (option "dw_stream_pd")
    emul.set_tw8(tw8)                                  # 8 output columns per lane (option "dw_stream_tw8")
    g = torch.Generator().manual_seed(N * 100 + H + C)
    x = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16)
    w = (torch.randn(C, 1, k, k, generator=g) / k).float()
    b = (torch.randn(C, generator=g) * 0.1).float()
    s = (torch.rand(C, generator=g) + 0.5).float() if xform else None
    t = (torch.randn(C, generator=g) * 0.3).float() if xform else None
    dz = torch.randn(N, H, W, C, generator=g).to(torch.bfloat16)
    geo = (ctypes.c_int * 6)()
    # ---- forward ----
    z = torch.full((N, H, W, C), float("nan")).to(torch.bfloat16)
    stats = torch.zeros(2 * C, dtype=torch.float64)
    rc = emul(0, P(x), P(s), P(t), P(w), P(b), None, P(z), None, P(stats), N, H, W, C, k, warps, ctypes.addressof(geo))
    assert rc == 0
    PL, G, NB, HS, nws, nhs = list(geo)
    assert PL * G <= 32 and PL * NB * 2 == C and nhs * HS >= H and nws * (8 if tw8 else 4) * G >= W
    xa = x.float().permute(0, 3, 1, 2).double()
    a = torch.relu(xa * s.double()[None, :, None, None] + t.double()[None, :, None, None]) if xform else xa
    a_ = a.clone().requires_grad_(True)
    w_ = w.double().clone().requires_grad_(True)
    zref = F.conv2d(a_, w_, b.double(), padding=k // 2, groups=C)
    zf = z.float().permute(0, 3, 1, 2)
    assert torch.isfinite(zf).all()
    assert rel(zf, zref) < 4e-3                                   # bf16 output rounding
    zs = zf.double()
    torch.testing.assert_close(stats[:C], zs.sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(stats[C:], (zs * zs).sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)
    # ---- backward-data (input = dz, flipped kernel, no transform) and backward-weight ----
    zref.backward(dz.float().permute(0, 3, 1, 2).double())
    dx = torch.full((N, H, W, C), float("nan")).to(torch.bfloat16)
    rc = emul(1, P(dz), None, None, P(w), None, None, P(dx), None, None, N, H, W, C, k, warps, None)
    assert rc == 0
    dxf = dx.float().permute(0, 3, 1, 2)
    assert torch.isfinite(dxf).all()
    assert rel(dxf, a_.grad) < 4e-3
    dw = torch.zeros(C, 1, k, k)
    for _ in range(2):                                             # accumulates INTO dw
        rc = emul(2, P(x), P(s), P(t), None, None, P(dz), None, P(dw), None, N, H, W, C, k, warps, None)
        assert rc == 0
    assert rel(dw, 2 * w_.grad) < 1e-5
