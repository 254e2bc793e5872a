"""GPU: forward of the wide 1x1 ConvBlocks of the 28x28 / 14x14 stages on cp.async + mma.sync (csrc/pw_wide_fwd.cu behind
mnb_conv_fwd, option "pw_wide") against fp64 math on the same bf16 operands: output, BN statistics of the stored values;
expand (40->240, 80->480, 96->576) and project (240->40, 480->80, 576->96) shapes, ragged row counts, packed / unpacked
weights, with / without the fused BN-apply+ReLU of the producing block; and against the tcgen05 path (impl 2).
Reference call sites: src/models/mnasnet.py:58-62,116-128."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


SHAPES = [(576, 96), (480, 80), (240, 40), (96, 576), (80, 480), (40, 240)]


@pytest.mark.parametrize("packed,act", [(True, True), (False, False)])
@pytest.mark.parametrize("N,H,W", [(3, 14, 14), (2, 28, 28), (1, 5, 7), (1, 1, 1), (7, 14, 14)])
@pytest.mark.parametrize("Cin,Cout", SHAPES)
def test_pw_wide_forward(Cin, Cout, N, H, W, packed, act):
    from mnb200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(Cin + 3 * H)
    M = N * H * W
    x = (torch.randn(N, H, W, Cin, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5).float()
    sc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
    wpf = torch.empty(w.numel(), device="cuda", dtype=BF)
    wpd = torch.empty(w.numel(), device="cuda", dtype=BF)
    L.call("mnb_pack_weights", P(w), P(wpf), P(wpd), Cout, Cin, 1, S())
    res = []
    old = L.get_option("pw_wide")
    L.set_option("pw_wide", 1)          # not the default: measured slower than the tcgen05 pipeline (scripts/exp_pw_wide.py)
    for impl in (0, 2):
        z = torch.full((N, H, W, Cout), float("nan"), device="cuda", dtype=BF)
        st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
        L.call("mnb_conv_fwd_packed", P(x), P(sc) if act else None, P(sh) if act else None, P(w), P(wpf) if packed else None, None,
               P(z), P(st), N, H, W, Cin, Cout, 1, 1, 0, 1, 0, impl, S())
        torch.cuda.synchronize()
        res.append((z, st))
    L.set_option("pw_wide", old)
    z, st = res[0]
    a = x.float().view(M, Cin)
    if act:
        a = torch.relu(a * sc + sh).to(BF).float()
    zr = a.double() @ w.view(Cout, Cin).to(BF).double().t()
    assert torch.isfinite(z.float()).all()
    assert rel(z.view(M, Cout), zr) < 5e-3
    zs = z.double().view(M, Cout)
    torch.testing.assert_close(st[:Cout], zs.sum(0), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(st[Cout:], (zs * zs).sum(0), rtol=1e-5, atol=1e-4)
    assert rel(z.float(), res[1][0].float()) < 5e-3 and rel(st, res[1][1]) < 1e-3       # tcgen05 path: same to bf16 rounding
