"""GPU: the TMA + mma.sync depthwise kernels (csrc/dw_mma.cu) through the C ABI against torch fp64 math on the same
bf16 operands: forward (+BN-apply/ReLU prologue, +BN statistics), backward-data, backward-weight, and the fused
ConvBlock backward (BN-backward elementwise pass + backward-data + backward-weight + the producing block's BN-backward
reductions).  Replaces nn.Conv2d(groups=C) / native_batch_norm_backward of src/models/mnasnet.py:48-62,76-81,120-125.
Shapes: odd / ragged maps (borders in every direction, partial strips, 1-3 row blocks), channel counts that do not
fill a 24 / 40-channel group, and the MNASNet maps of the 224^2, 112^2 and 192x256 configurations."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2          # north_star bf16 gate; measured 1.7e-3 (one bf16 rounding of the outputs)

CASES = [(2, 12, 10, 32, 3), (2, 9, 11, 72, 5), (3, 7, 7, 48, 3), (2, 4, 4, 240, 5), (1, 17, 5, 1152, 3),
         (2, 14, 14, 576, 5), (2, 30, 20, 16, 5), (2, 56, 56, 72, 5), (1, 37, 45, 8, 3), (2, 25, 33, 40, 5),
         (3, 13, 50, 56, 3), (1, 64, 64, 24, 5), (2, 6, 8, 1152, 3), (2, 12, 16, 576, 5), (1, 128, 96, 32, 3),
         (1, 96, 128, 48, 3), (2, 112, 112, 48, 3), (3, 28, 28, 240, 5), (3, 1, 1, 24, 3), (1, 2, 40, 24, 5)]


def P(t):
    return None if t is None else t.data_ptr()


def S():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(params=[0, 2], ids=["strips-default", "strips-2"])
def forced_mma(request):
    from mnb200 import _lib as L
    old, old_small, old_tws = L.get_option("dw_mma"), L.get_option("dw_small"), L.get_option("dw_mma_tws")
    L.set_option("dw_mma", 2)           # 2 = every bf16 shape, not only the ones where it is the fastest kernel
    L.set_option("dw_small", 0)         # the whole-tile kernels (dw_small.cu, tests/test_dw_small_gpu.py) take the small maps otherwise
    L.set_option("dw_mma_tws", request.param)   # 0 = default geometry (one 16-column strip per CTA), 2 = two strips on wide maps
    yield L
    L.set_option("dw_mma", old)
    L.set_option("dw_small", old_small)
    L.set_option("dw_mma_tws", old_tws)


def _operands(N, H, W, C, k, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + 31 * H + C)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    dz = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
    sc = (torch.rand(C, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(C, device="cuda", generator=g) * 0.3).float()
    return x, dz, w, sc, sh


def _nchw64(t):
    return t.double().permute(0, 3, 1, 2)


@pytest.mark.parametrize("xform", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_dw_mma_forward_dgrad_wgrad(case, xform, forced_mma):
    L = forced_mma
    N, H, W, C, k = case
    x, dz, w, sc, sh = _operands(*case)
    z = torch.full_like(x, float("nan"))
    dx = torch.full_like(x, float("nan"))
    st = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dw = torch.zeros(C, 1, k, k, device="cuda")
    s_, t_ = (P(sc), P(sh)) if xform else (None, None)
    L.call("mnb_dw_fwd", P(x), s_, t_, P(w), None, P(z), P(st), N, H, W, C, k, 1, S())
    L.call("mnb_dw_dgrad", P(dz), P(w), P(dx), None, None, None, None, N, H, W, C, k, 1, S())
    for _ in range(2):                  # accumulates INTO dw
        L.call("mnb_dw_wgrad", P(x), s_, t_, P(dz), P(dw), N, H, W, C, k, 1, S())
    torch.cuda.synchronize()
    a = _nchw64(x)
    if xform:
        a = torch.relu(a * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
    a = a.requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    zr = F.conv2d(a, w64, None, padding=k // 2, groups=C)
    zr.backward(_nchw64(dz))
    assert torch.isfinite(z.float()).all() and torch.isfinite(dx.float()).all()
    assert rel(_nchw64(z), zr.detach()) < TOL
    zq = _nchw64(z)
    torch.testing.assert_close(st[:C], zq.sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)          # of the STORED values
    torch.testing.assert_close(st[C:], (zq * zq).sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)
    assert rel(_nchw64(dx), a.grad) < TOL
    assert rel(dw, 2 * w64.grad) < TOL


@pytest.mark.parametrize("act", [True, False])
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("small", [0, 2])
def test_dw_fused_backward_matches_torch(case, act, small):
    """mnb_dw_bwd_fused vs the SURVEY appendix-F math in fp64: G = dA*[scale*z+shift>0]; dZ = a*G + b*z + c;
    dx = conv_dgrad(dZ); dw += conv_wgrad(A, dZ); dgamma/dbeta; reductions of dx for the producing block.
    small = 0: the row-streaming kernel (dw_mma.cu) on every map; 2: the whole-tile kernel (dw_small.cu) on every map."""
    from mnb200 import _lib as L
    old_small = L.get_option("dw_small")
    L.set_option("dw_small", small)
    try:
        _fused_backward_case(L, case, act)
    finally:
        L.set_option("dw_small", old_small)


def _fused_backward_case(L, case, act):
    N, H, W, C, k = case
    g = torch.Generator(device="cuda").manual_seed(5 + 31 * H + C)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    z = (torch.randn(N, H, W, C, device="cuda", generator=g) * 0.7 + 0.2).to(BF)
    dA = torch.randn(N, H, W, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(C, 1, k, k, device="cuda", generator=g) / k).float()
    sc, isc = ((torch.rand(C, device="cuda", generator=g) + 0.5).float() for _ in range(2))
    sh, ish = ((torch.randn(C, device="cuda", generator=g) * 0.3).float() for _ in range(2))
    M = N * H * W
    z64 = z.double()
    mean = z64.mean(dim=(0, 1, 2))
    invstd = 1.0 / torch.sqrt(z64.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
    meanf, invf = mean.float().contiguous(), invstd.float().contiguous()
    sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dA), P(z), P(sc), P(sh), P(sums), M, C, 1, S())
    dx = torch.full_like(x, float("nan"))
    dw = torch.zeros(C, 1, k, k, device="cuda")
    ns = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    dga, dbe, dbi = (torch.zeros(C, device="cuda") for _ in range(3))
    L.call("mnb_dw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(meanf), P(invf), P(dga), P(dbe), P(dbi), P(x),
           P(isc) if act else None, P(ish) if act else None, P(w), P(dx), P(dw), P(ns) if act else None, N, H, W, C, k,
           float(M), 1, S())
    torch.cuda.synchronize()
    # fp64 reference from the same bf16 operands and the same (fp32) mean / invstd / scale
    Gm = dA.double() * ((z64 * sc.double() + sh.double()) > 0)
    sg, sgz = Gm.sum(dim=(0, 1, 2)), (Gm * z64).sum(dim=(0, 1, 2))
    torch.testing.assert_close(sums[:C], sg, rtol=1e-6, atol=1e-6)
    dgr = invf.double() * (sgz - meanf.double() * sg)
    b = -sc.double() * invf.double() * dgr / M
    c3 = -sc.double() * sg / M - b * meanf.double()
    dzr = (sc.double() * Gm + b * z64 + c3).permute(0, 3, 1, 2)
    a = _nchw64(x)
    if act:
        a = torch.relu(a * isc.double()[None, :, None, None] + ish.double()[None, :, None, None])
    a = a.requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    F.conv2d(a, w64, None, padding=k // 2, groups=C).backward(dzr)
    assert torch.isfinite(dx.float()).all()
    assert rel(_nchw64(dx), a.grad) < TOL
    assert rel(dw, w64.grad) < TOL
    assert rel(dga, dgr) < 1e-4 and rel(dbe, sg) < 1e-4
    assert dbi.abs().max().item() <= 1e-3 * max(1.0, sg.abs().max().item())       # analytically zero
    if act:
        # reductions of the producing block: of the STORED dx, masked by that block's ReLU, against its raw output
        dxs = dx.double()
        msk = (x.double() * isc.double() + ish.double()) > 0
        torch.testing.assert_close(ns[:C], (dxs * msk).sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(ns[C:], (dxs * msk * x.double()).sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
    # frozen weight / BN parameters: NULL gradient slots are skipped, dx unchanged
    dx2 = torch.full_like(x, float("nan"))
    L.call("mnb_dw_bwd_fused", P(dA), P(z), P(sc), P(sh), P(sums), P(meanf), P(invf), None, None, None, P(x),
           P(isc) if act else None, P(ish) if act else None, P(w), P(dx2), None, None, N, H, W, C, k, float(M), 1, S())
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2)


def test_dw_fused_backward_rejects_bad_arguments():
    from mnb200 import _lib as L
    t = torch.zeros(1, 4, 4, 8, device="cuda", dtype=BF)
    f = torch.zeros(8, device="cuda")
    d = torch.zeros(16, device="cuda", dtype=torch.float64)
    w = torch.zeros(8, 1, 3, 3, device="cuda")
    args = [P(t), P(t), P(f), P(f), P(d), P(f), P(f), None, None, None, P(t), None, None, P(w), P(t), None]
    with pytest.raises(L.MnbError):      # reductions for the producing block need its scale / shift
        L.call("mnb_dw_bwd_fused", *args, P(d), 1, 4, 4, 8, 3, 16.0, 1, S())
    with pytest.raises(L.MnbError):      # fp32 activations: unsupported, the caller keeps the unfused chain
        L.call("mnb_dw_bwd_fused", *args, None, 1, 4, 4, 8, 3, 16.0, 0, S())
    with pytest.raises(L.MnbError):
        L.call("mnb_dw_bwd_fused", *args, None, 1, 4, 4, 8, 7, 16.0, 1, S())


@pytest.mark.parametrize("n,h,w", [(4, 224, 224), (3, 96, 128)])
def test_engine_fused_depthwise_backward_matches_unfused(n, h, w):
    """Whole bf16 network: ONE forward, then the backward program twice from the same saved activations (same ReLU
    masks, same dlogits) -- the unfused chain, then the fused depthwise backward forced on every depthwise block.
    Backward is linear in dlogits once the masks are fixed, so the two gradient vectors must agree to bf16 rounding
    (a second forward would not: bf16 forwards differ run to run by the order of the BN atomics)."""
    from mnb200 import engine
    from oracle import mnasnet_oracle as O
    from test_net_gpu import build
    x, t = O.synthetic_batch(n, h, w)
    m = build("bf16")
    eng = engine.engine_for(m)
    eng.fuse_dw_bwd = 0
    eng.fuse_pw_bwd = 0                  # (the fused pointwise backward has its own test, tests/test_pw_bwd_gpu.py)
    eng.wgrad_slack = 0                  # single stream: the second program may reuse the first one's scratch
    out = m(x.cuda())
    loss = torch.nn.CrossEntropyLoss()(out, t.cuda())
    loss.backward()
    torch.cuda.synchronize()
    plan = eng.plan(n, h, w)
    assert not any(getattr(op, "label", "").endswith("_bwd_fused") for op in plan.bwd)
    g0 = eng.store.grad.clone()
    # rebuild only the backward program with the fused kernel everywhere, rerun it on the same forward state
    eng.fuse_dw_bwd = 2
    for a in plan.apps:
        a.reduce_fused = False
    plan.bwd = []
    plan.last_write = {}
    plan._emit_backward()
    n_fused = sum(getattr(op, "label", "").endswith("_bwd_fused") for op in plan.bwd)
    assert n_fused == 17                 # every depthwise application
    assert not any(getattr(op, "name", "") == "mnb_dw_wgrad" for op in plan.bwd)
    plan.dstats.zero_()                  # the BN-backward sums accumulate; the forward statistics are already consumed
    eng.backward(plan)
    torch.cuda.synchronize()
    g2 = eng.store.grad.clone()
    assert torch.isfinite(g2).all()
    cos = (g0 @ g2 / (g0.norm() * g2.norm())).item()
    err = ((g0 - g2).norm() / g0.norm()).item()
    print(f"fused-vs-unfused depthwise backward on one forward: gradient cosine {cos:.5f}, rel-L2 {err:.2e}")
    assert cos > 0.999 and err < 2e-2
    # per stage (features.N / classifier): both programs round dZ to bf16 at different places, and the BN scale / shift
    # gradients of the early layers are small residuals of large cancelling sums -- the stage totals agree, single
    # 32-element vectors near the stem need not
    views0, views2 = eng.store.grad_views(g0), eng.store.grad_views(g2)
    stages = {}
    for k, p in m.named_parameters():
        if k.endswith("conv.bias"):
            continue                     # analytically zero
        key = ".".join(k.split(".")[:2]) if k.startswith("features") else "classifier"
        d = stages.setdefault(key, [0.0, 0.0])
        d[0] += (views0[id(p)] - views2[id(p)]).double().pow(2).sum().item()
        d[1] += views0[id(p)].double().pow(2).sum().item()
    for key, (num, den) in stages.items():
        e = (num / max(den, 1e-30)) ** 0.5
        assert e < (1e-5 if key == "classifier" else 0.1), (key, e)
