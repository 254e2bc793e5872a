"""GPU: the whole-tile depthwise kernels for small maps (csrc/dw_small.cu) behind mnb_dw_fwd / mnb_dw_dgrad / mnb_dw_wgrad,
against torch fp32 math on the same bf16 operands: forward + BN statistics of the stored values, backward-data,
backward-weight.  Option dw_small = 2 routes every shape there (default 1: maps of 12..64 rows, the 56 x 56 .. 14 x 14
stages of MnasNet: nn.Conv2d(groups=C) at src/models/mnasnet.py:76-81,120-125).  Shapes: one and two row tiles,
ragged heights / widths, one and two 16-column strips, 24- and 40-channel groups with partial last groups, 7 x 7 and 4 x 4
maps (TH = 7 tiles), with / without the fused BN-apply+ReLU of the producing block.  Gate 1e-2 rel-L2, statistics 1e-5."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def forced_small():
    from mnb200 import _lib as L
    old = L.get_option("dw_small")
    L.set_option("dw_small", 2)
    yield L
    L.set_option("dw_small", old)


CASES = [(3, 14, 14, 48, 5), (3, 14, 14, 48, 3), (2, 28, 28, 72, 5), (2, 28, 28, 40, 3), (2, 7, 7, 96, 5), (2, 7, 7, 48, 3),
         (2, 16, 24, 32, 5), (2, 12, 16, 56, 3), (2, 24, 32, 24, 5), (3, 4, 4, 80, 5), (2, 28, 20, 240, 5), (2, 14, 14, 576, 5),
         (2, 14, 14, 480, 3), (2, 20, 20, 88, 3), (1, 6, 8, 1152, 3), (2, 27, 17, 16, 5), (2, 15, 33, 8, 3),
         (2, 56, 56, 72, 5), (1, 112, 112, 48, 3), (1, 40, 64, 24, 5), (1, 64, 48, 32, 3), (1, 113, 50, 16, 3)]


@pytest.mark.parametrize("act", [True, False])
@pytest.mark.parametrize("N,H,W,C,k", CASES)
def test_dw_small_matches_torch(N, H, W, C, k, act, forced_small):
    L = forced_small
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(3 + C + 7 * H + k)
        x = (torch.randn(N, H, W, C, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
        dz = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
        w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
        sc = (torch.rand(C, device="cuda", generator=g) + 0.5).float()
        sh = (torch.randn(C, device="cuda", generator=g) * 0.3).float()
        z = torch.full_like(x, float("nan"))
        dx = torch.full_like(x, float("nan"))
        dw = torch.zeros_like(w)
        st = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
        a_sc, a_sh = (P(sc), P(sh)) if act else (None, None)
        L.call("mnb_dw_fwd", P(x), a_sc, a_sh, P(w), None, P(z), P(st), N, H, W, C, k, 1, S())
        L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())
        L.call("mnb_dw_wgrad", P(x), a_sc, a_sh, P(dz), P(dw), N, H, W, C, k, 1, S())
        torch.cuda.synchronize()
        A = x.float()
        if act:
            A = torch.relu(A * sc + sh).to(BF).float()
        A = A.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        wb = w.to(BF).float().requires_grad_(True)
        zr = F.conv2d(A, wb, None, stride=1, padding=k // 2, groups=C)
        gA, gw = torch.autograd.grad(zr, [A, wb], dz.float().permute(0, 3, 1, 2))
        assert not torch.isnan(z.float()).any() and not torch.isnan(dx.float()).any()
        assert rel(z.float().permute(0, 3, 1, 2), zr) < 1e-2
        assert rel(dx.float().permute(0, 3, 1, 2), gA) < 1e-2
        assert rel(dw, gw) < 1e-2
        zs = z.double()
        torch.testing.assert_close(st[:C], zs.sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(st[C:], (zs * zs).sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_dw_small_agrees_with_the_row_streaming_kernels():
    """Same operands through dw_small (default on a 14 x 14 map) and dw_mma (dw_small = 0): equal to bf16 rounding."""
    from mnb200 import _lib as L
    N, H, W, C, k = 4, 14, 14, 96, 5
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    dz = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
    sc = (torch.rand(C, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(C, device="cuda", generator=g) * 0.3).float()
    res = []
    old = L.get_option("dw_small")
    try:
        for small in (1, 0):
            L.set_option("dw_small", small)
            z = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.zeros_like(w)
            st = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
            L.call("mnb_dw_fwd", P(x), P(sc), P(sh), P(w), None, P(z), P(st), N, H, W, C, k, 1, S())
            L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())
            L.call("mnb_dw_wgrad", P(x), P(sc), P(sh), P(dz), P(dw), N, H, W, C, k, 1, S())
            torch.cuda.synchronize()
            res.append((z.float(), st, dx.float(), dw))
    finally:
        L.set_option("dw_small", old)
    for a, b in zip(*res):
        assert rel(a, b) < 5e-3
