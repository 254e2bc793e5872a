"""GPU: backward of the project 1x1 ConvBlocks after their BN pass (csrc/pw_proj_bwd.cu, mnb_pw_proj_bwd): backward-data
(+ residual skip gradient), backward-weight and the producer's BatchNorm-backward reductions in one kernel, against fp64
math on the same bf16 operands (src/models/mnasnet.py:58-62,120-128 under autograd).  Shapes: the
three instantiated slice geometries (576->96, 480->80, 240->40), ragged row counts (partial last 96-row tile, fewer rows
than one tile), with / without skip gradient, activation, weight gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
P = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("act,add,wgrad", [(True, True, True), (True, False, True), (False, False, True), (True, True, False)])
@pytest.mark.parametrize("M,Cin,Cout", [(3 * 14 * 14, 576, 96), (2 * 28 * 28, 240, 40), (5 * 14 * 14, 480, 80), (1000, 192, 96),
                                        (50, 160, 40), (96 * 7 + 1, 576, 96), (4097, 480, 80)])
def test_pw_proj_bwd_matches_fp64(M, Cin, Cout, act, add, wgrad):
    from mnb200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M + Cin)
    dz = torch.randn(M, Cout, device="cuda", generator=g).to(BF)
    x = (torch.randn(M, Cin, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
    sk = torch.randn(M, Cin, device="cuda", generator=g).to(BF) if add else None
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5).float()
    isc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float()
    ish = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
    dx = torch.full((M, Cin), float("nan"), device="cuda", dtype=BF)
    dw = torch.zeros_like(w)
    ns = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    L.call("mnb_pw_proj_bwd", P(dz), P(x), P(isc) if act else None, P(ish) if act else None, P(w), P(sk), P(dx),
           P(dw) if wgrad else None, P(ns) if act else None, M, Cin, Cout, 1, S())
    torch.cuda.synchronize()
    wb = w.view(Cout, Cin).to(BF).double()
    dxr = dz.double() @ wb
    if add:
        dxr = dxr + sk.double()
    assert torch.isfinite(dx.float()).all()
    assert rel(dx, dxr) < 5e-3
    if wgrad:
        a = x.double()
        if act:
            a = torch.relu(x.float() * isc + ish).to(BF).double()
        assert rel(dw.view(Cout, Cin), dz.double().t() @ a) < 1e-3
    else:
        assert dw.abs().max().item() == 0
    if act:
        dxs = dx.double()
        msk = (x.float() * isc + ish) > 0
        torch.testing.assert_close(ns[:Cin], (dxs * msk).sum(0), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(ns[Cin:], (dxs * msk * x.double()).sum(0), rtol=1e-5, atol=1e-4)


def test_pw_proj_bwd_rejects_other_shapes():
    from mnb200 import _lib as L
    t = torch.zeros(64, 96, device="cuda", dtype=BF)
    x = torch.zeros(64, 100, device="cuda", dtype=BF)
    w = torch.zeros(96, 104, device="cuda")
    with pytest.raises(L.MnbError):
        L.call("mnb_pw_proj_bwd", P(t), P(x), None, None, P(w), None, P(x), None, None, 64, 104, 96, 1, S())
