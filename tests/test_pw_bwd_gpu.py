"""GPU: the fused pointwise (1x1) ConvBlock backward (csrc/pw_bwd_fused.cu, mnb_pw_bwd_fused) through the C ABI against
the SURVEY appendix-F math in fp64 on the same bf16 operands -- BatchNorm-backward elementwise pass, conv
backward-data (+ residual skip gradient), conv backward-weight, dgamma / dbeta, and the producing block's BatchNorm
reductions -- for every instantiated (Cin, Cout), ragged row counts (tiles that end mid-way, single rows) and all
combinations of input activation / skip gradient.  Replaces what autograd runs for the MBConv expand / project blocks
(src/models/mnasnet.py:116-129) in the 112x112 / 56x56 stages."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2
SHAPES = [(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)]
ROWS = [1, 37, 96, 97, 2 * 28 * 28 + 5, 4 * 56 * 56]


def P(t):
    return None if t is None else t.data_ptr()


def S():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("act,add", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("M", ROWS)
@pytest.mark.parametrize("cin,cout", SHAPES)
def test_pw_fused_backward_matches_torch(cin, cout, M, act, add):
    from mnb200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(3 + 7 * cin + cout + M)
    x = torch.randn(M, cin, device="cuda", generator=g).to(BF)
    z = (torch.randn(M, cout, device="cuda", generator=g) * 0.7 + 0.2).to(BF)
    dA = torch.randn(M, cout, device="cuda", generator=g).to(BF)
    sk = torch.randn(M, cin, device="cuda", generator=g).to(BF) if add else None
    w = (torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5).float()
    sc, isc = ((torch.rand(c, device="cuda", generator=g) + 0.5).float() for c in (cout, cin))
    sh, ish = ((torch.randn(c, device="cuda", generator=g) * 0.3).float() for c in (cout, cin))
    z64 = z.double()
    meanf = z64.mean(0).float().contiguous()
    invf = (1.0 / torch.sqrt(z64.var(0, unbiased=False) + 1e-5)).float().contiguous()
    sums = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, cout, 1, S())
    dx = torch.full_like(x, float("nan"))
    dw = torch.zeros_like(w)
    ns = torch.zeros(2 * cin, device="cuda", dtype=torch.float64)
    dga, dbe, dbi = (torch.zeros(cout, device="cuda") for _ in range(3))
    for _ in range(2):      # dw / dgamma / dbeta / in_sums accumulate
        L.call("mnb_pw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(meanf), P(invf), P(dga), P(dbe), P(dbi), P(x),
               P(isc) if act else None, P(ish) if act else None, P(w), P(sk), P(dx), P(dw), P(ns) if act else None, M, cin,
               cout, float(M), 1, S())
    torch.cuda.synchronize()
    Gm = dA.double() * ((z64 * sc.double() + sh.double()) > 0)
    sg, sgz = Gm.sum(0), (Gm * z64).sum(0)
    dgr = invf.double() * (sgz - meanf.double() * sg)
    b = -sc.double() * invf.double() * dgr / M
    c3 = -sc.double() * sg / M - b * meanf.double()
    dzr = sc.double() * Gm + b * z64 + c3
    a = x.double()
    if act:
        a = torch.relu(a * isc.double() + ish.double())
    w64 = w.double().view(cout, cin)
    dxr = dzr @ w64 + (sk.double() if add else 0)
    dwr = dzr.t() @ a
    assert torch.isfinite(dx.float()).all()
    if M > 1:               # with one row BatchNorm backward is identically zero: only absolute checks make sense
        assert rel(dx, dxr) < TOL
        assert rel(dw.view(cout, cin), 2 * dwr) < TOL
        assert rel(dga, 2 * dgr) < 1e-4 and rel(dbe, 2 * sg) < 1e-4
    else:
        assert (dx.double() - dxr).abs().max().item() < 1e-2
    if act:
        msk = (x.double() * isc.double() + ish.double()) > 0
        torch.testing.assert_close(ns[:cin], 2 * (dx.double() * msk).sum(0), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(ns[cin:], 2 * (dx.double() * msk * x.double()).sum(0), rtol=1e-5, atol=1e-4)


def test_pw_fused_backward_unsupported_shapes_are_reported():
    from mnb200 import _lib as L
    t40 = torch.zeros(8, 40, device="cuda", dtype=BF)
    t240 = torch.zeros(8, 240, device="cuda", dtype=BF)
    f = torch.zeros(240, device="cuda")
    d = torch.zeros(480, device="cuda", dtype=torch.float64)
    w = torch.zeros(240, 40, 1, 1, device="cuda")
    rc = L.lib.mnb_pw_bwd_fused(P(t240), P(t240), P(f), P(f), P(d), P(f), P(f), None, None, None, P(t40), None, None, P(w),
                                None, P(t40), None, None, 8, 40, 240, 8.0, 1, S())
    assert rc == -2 or rc != 0          # MNB_ERR_UNSUPPORTED: the engine keeps the unfused chain for this layer


@pytest.mark.parametrize("n,h,w", [(4, 224, 224), (3, 96, 128)])
def test_engine_fused_pointwise_backward_matches_unfused(n, h, w):
    """Whole bf16 network, ONE forward, the backward program twice on the same saved activations: unfused chain vs the
    fused pointwise (+ fused depthwise) backward kernels.  Same masks, same dlogits -> gradients agree to bf16 rounding."""
    from mnb200 import engine
    from oracle import mnasnet_oracle as O
    from test_net_gpu import build
    x, t = O.synthetic_batch(n, h, w)
    m = build("bf16")
    eng = engine.engine_for(m)
    eng.fuse_dw_bwd = 0
    eng.fuse_pw_bwd = 0
    eng.wgrad_slack = 0
    out = m(x.cuda())
    torch.nn.CrossEntropyLoss()(out, t.cuda()).backward()
    torch.cuda.synchronize()
    plan = eng.plan(n, h, w)
    assert not any(getattr(op, "label", "").endswith("_bwd_fused") for op in plan.bwd)
    g0 = eng.store.grad.clone()
    eng.fuse_dw_bwd, eng.fuse_pw_bwd = 1, 1
    for a in plan.apps:
        a.reduce_fused = False
    plan.bwd = []
    plan.last_write = {}
    plan._emit_backward()
    labels = [getattr(op, "label", "") for op in plan.bwd]
    assert sum(l == "pw1x1_bwd_fused" for l in labels) == 13          # 3x(16->48, 48->16) + 32->16 + 3x(24->72, 72->24)
    plan.dstats.zero_()
    eng.backward(plan)
    torch.cuda.synchronize()
    g2 = eng.store.grad.clone()
    assert torch.isfinite(g2).all()
    cos = (g0 @ g2 / (g0.norm() * g2.norm())).item()
    err = ((g0 - g2).norm() / g0.norm()).item()
    print(f"fused-vs-unfused pointwise backward on one forward: gradient cosine {cos:.5f}, rel-L2 {err:.2e}")
    assert cos > 0.999 and err < 3e-2
