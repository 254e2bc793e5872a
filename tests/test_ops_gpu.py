"""GPU: every C-ABI kernel against a plain torch fp64 CPU reference of the same op (teacher-forced, T1).
Tolerances: fp32 mode <= 1e-4 rel-L2 (north_star), bf16 mode <= 2e-2 rel-L2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _lib():
    from mnb200 import _lib
    return _lib


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def tdt(dtype):
    return torch.float32 if dtype == "fp32" else torch.bfloat16


def code(dtype):
    return 0 if dtype == "fp32" else 1


def nhwc(x, dtype):   # NCHW cpu fp64 -> NHWC cuda T
    return x.permute(0, 2, 3, 1).contiguous().to(device="cuda", dtype=tdt(dtype))


def nchw(y):          # NHWC cuda -> NCHW cpu fp64
    return y.permute(0, 3, 1, 2).double().cpu()


def stream():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def act(x, s, t):
    if s is None:
        return x
    return torch.relu(x * s[None, :, None, None] + t[None, :, None, None])


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, xform
    (2, 12, 10, 16, 48, 1, 1, 0, False),
    (2, 12, 10, 48, 16, 1, 1, 0, True),
    (3, 9, 7, 24, 72, 1, 1, 0, True),
    (2, 14, 14, 16, 24, 3, 2, 1, False),
    (2, 7, 9, 80, 96, 3, 1, 1, False),
    (1, 7, 7, 192, 320, 3, 1, 1, False),
    (2, 13, 11, 40, 80, 3, 2, 1, True),
]


# (Cin, Cout) of the 1x1 layers the warp-streaming kernels are instantiated for (csrc/pw_stream.cu)
STREAM_SHAPES = {(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)}


def _conv_setup(case, dtype, seed=0):
    N, H, W, Cin, Cout, k, stride, pad, xform = case
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g, dtype=torch.float64) * 0.1
    s = (torch.rand(Cin, generator=g, dtype=torch.float64) + 0.5) if xform else None
    t = (torch.randn(Cin, generator=g, dtype=torch.float64) * 0.3) if xform else None
    xd = nhwc(x, dtype)
    x = nchw(xd)                      # the values the kernel really sees (bf16-rounded in bf16 mode)
    return x, w, b, s, t, xd


@pytest.mark.parametrize("impl", ["simt", "auto"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, dtype, impl):
    L = _lib()
    N, H, W, Cin, Cout, k, stride, pad, xform = case
    im = {"simt": 1, "auto": 0}[impl]
    x, w, b, s, t, xd = _conv_setup(case, dtype)
    wd, bd = w.float().cuda(), b.float().cuda()
    sd = s.float().cuda() if xform else None
    td = t.float().cuda() if xform else None
    a = act(x, s.float().double() if xform else None, t.float().double() if xform else None)
    zref = F.conv2d(a, w.float().double(), b.float().double(), stride=stride, padding=pad)
    Ho, Wo = zref.shape[2], zref.shape[3]
    z = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=tdt(dtype))
    stats = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    L.call("mnb_conv_fwd", P(xd), P(sd), P(td), P(wd), P(bd), P(z), P(stats), N, H, W, Cin, Cout, k, stride, pad,
           code(dtype), 0, im, stream())
    torch.cuda.synchronize()
    assert rel(nchw(z), zref) < TOL[dtype]
    zs = nchw(z)
    m = N * Ho * Wo
    torch.testing.assert_close(stats[:Cout].cpu(), zs.sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-5 * m ** 0.5)
    torch.testing.assert_close(stats[Cout:].cpu(), (zs * zs).sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-5)
    # backward
    g = torch.Generator().manual_seed(5)
    dz = torch.randn(N, Cout, Ho, Wo, generator=g, dtype=torch.float64)
    dzd = nhwc(dz, dtype)
    dz = nchw(dzd)
    a_ = a.clone().requires_grad_(True)
    w_ = w.float().double().requires_grad_(True)
    F.conv2d(a_, w_, None, stride=stride, padding=pad).backward(dz)
    add = torch.randn(N, Cin, H, W, generator=g, dtype=torch.float64)
    addd = nhwc(add, dtype)
    dx = torch.empty(N, H, W, Cin, device="cuda", dtype=tdt(dtype))
    # fused BN-backward reduction of the producer block in the dgrad epilogue == mnb_bn_bwd_reduce on dx
    gq = torch.Generator().manual_seed(11)
    bz = nhwc(torch.randn(N, Cin, H, W, generator=gq, dtype=torch.float64), dtype)
    bsc = (torch.rand(Cin, generator=gq) + 0.5).cuda()
    bsh = (torch.randn(Cin, generator=gq) * 0.3).cuda()
    fused = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    L.call("mnb_conv_dgrad", P(dzd), P(wd), P(addd), P(dx), P(bz), P(bsc), P(bsh), P(fused), N, H, W, Cin, Cout, k,
           stride, pad, code(dtype), im, stream())
    sep = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dx), P(bz), P(bsc), P(bsh), P(sep), N * H * W, Cin, code(dtype), stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(fused, sep, rtol=1e-4, atol=1e-4 * sep.abs().max().item())
    dw = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.float32)
    L.call("mnb_conv_wgrad", P(xd), P(sd), P(td), P(dzd), P(dw), N, H, W, Cin, Cout, k, stride, pad, code(dtype), 0,
           im, stream())
    torch.cuda.synchronize()
    assert rel(nchw(dx), a_.grad + nchw(addd)) < TOL[dtype]
    assert rel(dw, w_.grad) < TOL[dtype]
    # plain dgrad (no residual / no fused reduction): stride-2 3x3 takes the parity-decomposed tcgen05 path
    dx3 = torch.full((N, H, W, Cin), 7.0, device="cuda", dtype=tdt(dtype))
    L.call("mnb_conv_dgrad", P(dzd), P(wd), None, P(dx3), None, None, None, None, N, H, W, Cin, Cout, k, stride, pad,
           code(dtype), im, stream())
    torch.cuda.synchronize()
    assert rel(nchw(dx3), a_.grad) < TOL[dtype]
    if dtype == "bf16" and im == 0:
        wpf = torch.empty(Cout * Cin * k * k, device="cuda", dtype=torch.bfloat16)
        wpd = torch.empty_like(wpf)
        L.call("mnb_pack_weights", P(wd), P(wpf), P(wpd), Cout, Cin, k, stream())
        z2 = torch.empty_like(z)
        L.call("mnb_conv_fwd_packed", P(xd), P(sd), P(td), P(wd), P(wpf), P(bd), P(z2), None, N, H, W, Cin, Cout, k,
               stride, pad, code(dtype), 0, im, stream())
        dx4 = torch.empty_like(dx3)
        L.call("mnb_conv_dgrad_packed", P(dzd), P(wd), P(wpd), None, P(dx4), None, None, None, None, N, H, W, Cin, Cout,
               k, stride, pad, code(dtype), im, stream())
        torch.cuda.synchronize()
        if k == 1 and (Cin, Cout) in STREAM_SHAPES:
            # auto routes packed 1x1 problems of these shapes to the warp-streaming kernels (pw_stream.cu): same
            # operands, different fp32 summation order -> equal up to single bf16 roundings
            assert rel(z2.float(), z.float()) < 1e-3 and rel(dx4.float(), dx3.float()) < 1e-3
            # ... and forcing tcgen05 with the packed weight reproduces the unpacked call bit for bit
            z3, dx5 = torch.empty_like(z), torch.empty_like(dx3)
            L.call("mnb_conv_fwd_packed", P(xd), P(sd), P(td), P(wd), P(wpf), P(bd), P(z3), None, N, H, W, Cin, Cout, k,
                   stride, pad, code(dtype), 0, 2, stream())
            L.call("mnb_conv_dgrad_packed", P(dzd), P(wd), P(wpd), None, P(dx5), None, None, None, None, N, H, W, Cin,
                   Cout, k, stride, pad, code(dtype), 2, stream())
            torch.cuda.synchronize()
            assert torch.equal(z3, z) and torch.equal(dx5, dx3)
        else:
            assert torch.equal(z2, z) and torch.equal(dx4, dx3)
    # accumulate semantics
    L.call("mnb_conv_wgrad", P(xd), P(sd), P(td), P(dzd), P(dw), N, H, W, Cin, Cout, k, stride, pad, code(dtype), 0,
           im, stream())
    torch.cuda.synchronize()
    assert rel(dw, 2 * w_.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_stem_conv_nchw_input(dtype):
    L = _lib()
    N, H, W, Cin, Cout = 2, 18, 14, 3, 32
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 5
    b = torch.randn(Cout, generator=g) * 0.1
    zref = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1)
    Ho, Wo = zref.shape[2:]
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    z = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=tdt(dtype))
    stats = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    L.call("mnb_conv_fwd", P(xd), None, None, P(wd), P(bd), P(z), P(stats), N, H, W, Cin, Cout, 3, 2, 1, code(dtype),
           1, 0, stream())
    torch.cuda.synchronize()
    assert rel(nchw(z), zref) < TOL[dtype]
    dz = torch.randn(N, Cout, Ho, Wo, generator=g, dtype=torch.float64)
    dzd = nhwc(dz, dtype)
    w_ = w.double().requires_grad_(True)
    F.conv2d(x.double(), w_, None, stride=2, padding=1).backward(nchw(dzd))
    dw = torch.zeros(Cout, Cin, 3, 3, device="cuda")
    L.call("mnb_conv_wgrad", P(xd), None, None, P(dzd), P(dw), N, H, W, Cin, Cout, 3, 2, 1, code(dtype), 1, 0,
           stream())
    torch.cuda.synchronize()
    assert rel(dw, w_.grad) < TOL[dtype]


DW_CASES = [(2, 12, 10, 32, 3, True), (2, 9, 11, 72, 5, True), (3, 7, 7, 48, 3, False), (2, 4, 4, 240, 5, True),
            (1, 17, 5, 1152, 3, True), (2, 14, 14, 576, 5, True), (1, 6, 8, 24, 5, False)]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", DW_CASES)
def test_depthwise_fwd_dgrad_wgrad(case, dtype):
    L = _lib()
    N, H, W, C, k, xform = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    w = (torch.randn(C, 1, k, k, generator=g, dtype=torch.float64) / k).float()
    b = (torch.randn(C, generator=g, dtype=torch.float64) * 0.1).float()
    s = (torch.rand(C, generator=g) + 0.5) if xform else None
    t = (torch.randn(C, generator=g) * 0.3) if xform else None
    xd = nhwc(x, dtype)
    x = nchw(xd)
    a = act(x, s.double() if xform else None, t.double() if xform else None)
    zref = F.conv2d(a, w.double(), b.double(), padding=k // 2, groups=C)
    wd, bd = w.cuda(), b.cuda()
    sd = s.cuda() if xform else None
    td = t.cuda() if xform else None
    z = torch.empty(N, H, W, C, device="cuda", dtype=tdt(dtype))
    stats = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_dw_fwd", P(xd), P(sd), P(td), P(wd), P(bd), P(z), P(stats), N, H, W, C, k, code(dtype), stream())
    torch.cuda.synchronize()
    assert rel(nchw(z), zref) < TOL[dtype]
    zs = nchw(z)
    torch.testing.assert_close(stats[:C].cpu(), zs.sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(stats[C:].cpu(), (zs * zs).sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-4)
    dz = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    dzd = nhwc(dz, dtype)
    dz = nchw(dzd)
    a_ = a.clone().requires_grad_(True)
    w_ = w.double().requires_grad_(True)
    F.conv2d(a_, w_, None, padding=k // 2, groups=C).backward(dz)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=tdt(dtype))
    gq = torch.Generator().manual_seed(12)
    bz = nhwc(torch.randn(N, C, H, W, generator=gq, dtype=torch.float64), dtype)
    bsc = (torch.rand(C, generator=gq) + 0.5).cuda()
    bsh = (torch.randn(C, generator=gq) * 0.3).cuda()
    fused = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_dw_dgrad", P(dzd), P(wd), P(dx), P(bz), P(bsc), P(bsh), P(fused), N, H, W, C, k, code(dtype), stream())
    sep = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dx), P(bz), P(bsc), P(bsh), P(sep), N * H * W, C, code(dtype), stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(fused, sep, rtol=1e-4, atol=1e-4 * sep.abs().max().item())
    dx2 = torch.empty_like(dx)
    L.call("mnb_dw_dgrad", P(dzd), P(wd), P(dx2), None, None, None, None, N, H, W, C, k, code(dtype), stream())
    torch.cuda.synchronize()
    if dtype == "fp32":
        assert torch.equal(dx, dx2)
    else:   # without the fused reduction bf16 maps of >= 12 rows take the tensor-pipe kernel (bf16 weights): same math,
        # one more rounding of the taps
        assert rel(dx2.float(), dx.float()) < 5e-3
    dw = torch.zeros(C, 1, k, k, device="cuda")
    L.call("mnb_dw_wgrad", P(xd), P(sd), P(td), P(dzd), P(dw), N, H, W, C, k, code(dtype), stream())
    torch.cuda.synchronize()
    assert rel(nchw(dx), a_.grad) < TOL[dtype]
    assert rel(dw, w_.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(4, 9, 7, 16), (2, 14, 14, 576), (3, 5, 5, 72), (1, 3, 4, 320)])
def test_batchnorm_forward_backward(shape, dtype):
    """stats -> finalize -> apply(+residual); reduce -> finalize -> apply (backward) vs torch autograd."""
    L = _lib()
    N, H, W, C = shape
    g = torch.Generator().manual_seed(7)
    z = torch.randn(N, C, H, W, generator=g, dtype=torch.float64) * 0.5 + 0.7
    zd = nhwc(z, dtype)
    z = nchw(zd)
    gamma = (torch.rand(C, generator=g) + 0.5)
    beta = torch.randn(C, generator=g) * 0.2
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    res = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    resd = nhwc(res, dtype)
    M = N * H * W
    stats = torch.stack([z.sum(dim=(0, 2, 3)), (z * z).sum(dim=(0, 2, 3))]).reshape(-1).cuda()
    gd, bd, rmd, rvd = gamma.cuda(), beta.cuda(), rm.clone().cuda(), rv.clone().cuda()
    nbt = torch.zeros(1, dtype=torch.long, device="cuda")
    vec = torch.zeros(4, C, device="cuda")
    L.call("mnb_bn_finalize", P(stats), P(gd), P(bd), P(rmd), P(rvd), P(nbt), P(vec[0]), P(vec[1]), P(vec[2]),
           P(vec[3]), C, float(M), 1e-5, 0.1, stream())
    y = torch.empty(N, H, W, C, device="cuda", dtype=tdt(dtype))
    L.call("mnb_bn_relu_apply", P(zd), P(vec[0]), P(vec[1]), P(resd), P(y), M, C, code(dtype), stream())
    torch.cuda.synchronize()
    z_ = z.clone().requires_grad_(True)
    g_ = gamma.double().requires_grad_(True)
    b_ = beta.double().requires_grad_(True)
    rm_, rv_ = rm.double().clone(), rv.double().clone()
    a = torch.relu(F.batch_norm(z_, rm_, rv_, g_, b_, True, 0.1, 1e-5))
    yref = a + nchw(resd)
    assert rel(nchw(y), yref) < TOL[dtype]
    torch.testing.assert_close(rmd.cpu().double(), rm_, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rvd.cpu().double(), rv_, rtol=1e-5, atol=1e-6)
    assert int(nbt) == 1
    # backward
    dA = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    dAd = nhwc(dA, dtype)
    a.backward(nchw(dAd))
    sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    L.call("mnb_bn_bwd_reduce", P(dAd), P(zd), P(vec[0]), P(vec[1]), P(sums), M, C, code(dtype), stream())
    dg, db, dbias = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    coef = torch.zeros(3, C, device="cuda")
    L.call("mnb_bn_bwd_finalize", P(sums), P(vec[0]), P(vec[2]), P(vec[3]), P(dg), P(db), P(dbias), P(coef), C,
           float(M), stream())
    dZ = torch.empty(N, H, W, C, device="cuda", dtype=tdt(dtype))
    L.call("mnb_bn_bwd_apply", P(dAd), P(zd), P(vec[0]), P(vec[1]), P(coef), P(dZ), M, C, code(dtype), stream())
    torch.cuda.synchronize()
    assert rel(dg, g_.grad) < 1e-4
    assert rel(db, b_.grad) < 1e-4
    assert rel(nchw(dZ), z_.grad) < TOL[dtype]
    assert dbias.abs().max().item() <= 1e-5 * max(1.0, z_.grad.abs().sum().item())
    # finalize + apply in one launch
    dg2, db2, dbias2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dZ2 = torch.empty_like(dZ)
    L.call("mnb_bn_bwd_apply_fused", P(dAd), P(zd), P(vec[0]), P(vec[1]), P(sums), P(vec[2]), P(vec[3]), P(dg2), P(db2),
           P(dbias2), P(dZ2), M, C, float(M), code(dtype), stream())
    torch.cuda.synchronize()
    assert torch.equal(dZ2, dZ) and torch.equal(dg2, dg) and torch.equal(db2, db)
    # folded conv bias: statistics of z - b with the bias handed to finalize give the same activation / buffers
    cbias = (torch.randn(C, generator=g) * 0.3)
    zb = nhwc(z - cbias.double()[None, :, None, None], "fp32")
    zbn = nchw(zb)
    stats_b = torch.stack([zbn.sum(dim=(0, 2, 3)), (zbn * zbn).sum(dim=(0, 2, 3))]).reshape(-1).cuda()
    rm2, rv2 = rm.clone().cuda(), rv.clone().cuda()
    vec2 = torch.zeros(4, C, device="cuda")
    L.call("mnb_bn_finalize_fb", P(stats_b), P(gd), P(bd), P(cbias.cuda()), P(rm2), P(rv2), None, P(vec2[0]),
           P(vec2[1]), P(vec2[2]), P(vec2[3]), C, float(M), 1e-5, 0.1, stream())
    torch.cuda.synchronize()
    act_a = torch.relu(vec[0].cpu().double()[None, :, None, None] * z + vec[1].cpu().double()[None, :, None, None])
    act_b = torch.relu(vec2[0].cpu().double()[None, :, None, None] * zbn + vec2[1].cpu().double()[None, :, None, None])
    assert rel(act_b, act_a) < 1e-4
    torch.testing.assert_close(rm2, rmd, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv2, rvd, rtol=1e-4, atol=1e-5)
    # eval coefficients
    sc, sh = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    L.call("mnb_bn_eval_coeffs", P(gd), P(bd), P(rmd), P(rvd), P(sc), P(sh), C, 1e-5, stream())
    torch.cuda.synchronize()
    ref_s = gamma.double() / torch.sqrt(rvd.cpu().double() + 1e-5)
    assert rel(sc, ref_s) < 1e-6
    assert rel(sh, beta.double() - rmd.cpu().double() * ref_s) < 1e-5


def test_head_kernels():
    L = _lib()
    N, HW, C, Hd, O = 5, 12, 320, 512, 1000
    g = torch.Generator().manual_seed(9)
    z = torch.randn(N, HW, C, generator=g)
    s, t = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    zd, sd, td = z.cuda(), s.cuda(), t.cuda()
    f = torch.empty(N, C, device="cuda")
    L.call("mnb_gap_fwd", P(zd), P(sd), P(td), P(f), N, HW, C, 0, stream())
    fref = torch.relu(z.double() * s.double() + t.double()).mean(dim=1)
    torch.cuda.synchronize()
    assert rel(f, fref) < 1e-6
    dA = torch.empty(N, HW, C, device="cuda")
    df = torch.randn(N, C, generator=g).cuda()
    L.call("mnb_gap_bwd", P(df), P(dA), N, HW, C, 0, stream())
    torch.cuda.synchronize()
    assert rel(dA, (df.cpu().double() / HW)[:, None, :].expand(N, HW, C)) < 1e-6
    # dropout mask statistics + determinism
    n = 1 << 20
    m1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    m2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    step = torch.zeros(1, dtype=torch.long, device="cuda")
    L.call("mnb_dropout_mask", P(m1), n, 0.5, 123, 0, P(step), stream())
    L.call("mnb_dropout_mask", P(m2), n, 0.5, 123, 0, P(step), stream())
    assert torch.equal(m1, m2)
    assert abs(m1.float().mean().item() - 0.5) < 5e-3
    step += 1
    L.call("mnb_dropout_mask", P(m2), n, 0.5, 123, 0, P(step), stream())
    assert not torch.equal(m1, m2)
    L.call("mnb_dropout_mask", P(m2), n, 0.2, 7, 0, None, stream())
    assert abs(m2.float().mean().item() - 0.8) < 5e-3
    # fc
    x = torch.randn(N, C, generator=g)
    w1, b1 = torch.randn(Hd, C, generator=g) / 18, torch.randn(Hd, generator=g) * 0.1
    mask = (torch.rand(N, C, generator=g) > 0.5).to(torch.uint8)
    xd, w1d, b1d, md = x.cuda(), w1.cuda(), b1.cuda(), mask.cuda()
    y = torch.empty(N, Hd, device="cuda")
    L.call("mnb_fc_fwd", P(xd), P(md), 2.0, P(w1d), P(b1d), P(y), 1, N, C, Hd, stream())
    x_ = x.double().requires_grad_(True)
    w_ = w1.double().requires_grad_(True)
    b_ = b1.double().requires_grad_(True)
    h = torch.relu(F.linear(x_ * mask.double() * 2.0, w_, b_))
    torch.cuda.synchronize()
    assert rel(y, h) < 1e-5
    dy = torch.randn(N, Hd, generator=g, dtype=torch.float64)
    dpre = dy * (h.detach() > 0)
    h.backward(dy)
    dw, dbb = torch.zeros(Hd, C, device="cuda"), torch.zeros(Hd, device="cuda")
    dpd = dpre.float().cuda()
    L.call("mnb_fc_wgrad", P(xd), P(md), 2.0, P(dpd), P(dw), P(dbb), N, C, Hd, stream())
    dx = torch.empty(N, C, device="cuda")
    L.call("mnb_fc_dgrad", P(dpd), P(w1d), P(md), 2.0, None, P(dx), N, C, Hd, stream())
    torch.cuda.synchronize()
    assert rel(dw, w_.grad) < 1e-5 and rel(dbb, b_.grad) < 1e-5 and rel(dx, x_.grad) < 1e-5
    # relu gating in dgrad
    ref = torch.randn(N, C, generator=g)
    L.call("mnb_fc_dgrad", P(dpd), P(w1d), None, 1.0, P(ref.cuda()), P(dx), N, C, Hd, stream())
    torch.cuda.synchronize()
    assert rel(dx, (dpre @ w1.double()) * (ref.double() > 0)) < 1e-5
    # cross entropy
    logits = torch.randn(N, O, generator=g) * 3
    tgt = torch.randint(0, O, (N,), generator=g)
    ld, tg = logits.cuda(), tgt.cuda()
    loss = torch.zeros(1, device="cuda")
    dl = torch.empty(N, O, device="cuda")
    L.call("mnb_xent_fwd_bwd", P(ld), P(tg), P(loss), P(dl), N, O, 1.0, stream())
    l_ = logits.double().requires_grad_(True)
    lref = F.cross_entropy(l_, tgt)
    lref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - lref.item()) < 1e-5
    assert rel(dl, l_.grad) < 1e-5


def test_adam_matches_torch_optim():
    L = _lib()
    n = 100003
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(n, generator=g)
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([q], lr=1e-3)
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.long, device="cuda")
    for i in range(5):
        gr = torch.randn(n, generator=g) * (10.0 ** (i - 3))
        q.grad = gr.clone()
        opt.step()
        L.call("mnb_counter_inc", P(step), stream())
        L.call("mnb_adam_step", P(p), P(gr.cuda()), P(m), P(v), n, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, None, P(step),
               stream())
    torch.cuda.synchronize()
    assert int(step) == 5
    assert (p.cpu() - q.data).abs().max().item() < 2e-6


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_layout_roundtrip(dtype):
    L = _lib()
    N, C, H, W = 2, 24, 5, 7
    x = torch.randn(N, C, H, W, device="cuda")
    y = torch.empty(N, H, W, C, device="cuda", dtype=tdt(dtype))
    L.call("mnb_nchw_f32_to_nhwc", P(x), P(y), N, H, W, C, code(dtype), stream())
    back = torch.empty_like(x)
    L.call("mnb_nhwc_to_nchw_f32", P(y), P(back), N, H, W, C, code(dtype), stream())
    torch.cuda.synchronize()
    assert torch.equal(y.float(), x.permute(0, 2, 3, 1).to(tdt(dtype)).float())
    assert torch.equal(back, x.to(tdt(dtype)).float())


def test_argument_errors_are_reported():
    L = _lib()
    x = torch.zeros(64, device="cuda")
    rc = L.lib.mnb_dw_fwd(P(x), None, None, P(x), None, P(x), None, 1, 4, 4, 12, 3, 0, stream())
    assert rc == -1 and b"multiple of 8" in L.lib.mnb_last_error()
    rc = L.lib.mnb_conv_fwd(P(x), None, None, P(x), None, P(x), None, 1, 4, 4, 8, 8, 5, 1, 2, 0, 0, 0, stream())
    assert rc == -1
    with pytest.raises(L.MnbError):
        L.call("mnb_bn_relu_apply", P(x), P(x), P(x), None, P(x), 4, 12, 0, stream())


@pytest.mark.parametrize("N,K,O,relu", [(37, 320, 512, 1), (256, 512, 1000, 0), (5, 320, 1003, 0)])
def test_fc_on_tcgen05(N, K, O, relu):
    """Classifier GEMMs on the tensor pipe (bf16 operands, fp32 accumulate/output) vs fp64."""
    L = _lib()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(N, K, generator=g)
    w = torch.randn(O, K, generator=g) / math.sqrt(K)
    b = torch.randn(O, generator=g) * 0.1
    mask = (torch.rand(N, K, generator=g) > 0.5).to(torch.uint8)
    xd, wd, bd, md = x.cuda(), w.cuda(), b.cuda(), mask.cuda()
    xb = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    pf = torch.empty(O * K, device="cuda", dtype=torch.bfloat16)
    pd = torch.empty(O * K, device="cuda", dtype=torch.bfloat16)
    L.call("mnb_pack_weights", P(wd), P(pf), P(pd), O, K, 1, stream())
    L.call("mnb_fc_prep_bf16", P(xd), P(md), 2.0, P(xb), N * K, stream())
    y = torch.empty(N, O, device="cuda")
    L.call("mnb_fc_fwd_tc", P(xb), P(wd), P(pf), P(bd), P(y), relu, N, K, O, stream())
    torch.cuda.synchronize()
    xm = xb.float().cpu().double()
    assert rel(xm, x.double() * mask.double() * 2.0) < 5e-3
    ref = xm @ w.to(torch.bfloat16).double().t() + b.double()
    ref = torch.relu(ref) if relu else ref
    assert rel(y, ref) < 1e-4                     # same bf16 operands, fp32 accumulate
    if O % 8 == 0:
        dy = torch.randn(N, O, generator=g)
        dyb = torch.empty(N, O, device="cuda", dtype=torch.bfloat16)
        L.call("mnb_fc_prep_bf16", P(dy.cuda()), None, 1.0, P(dyb), N * O, stream())
        dx = torch.empty(N, K, device="cuda")
        L.call("mnb_fc_dgrad_tc", P(dyb), P(wd), P(pd), P(dx), N, K, O, stream())
        ref_dx = dyb.float().cpu().double() @ w.to(torch.bfloat16).double()
        torch.cuda.synchronize()
        assert rel(dx, ref_dx) < 1e-4
        rr = torch.randn(N, K, generator=g)
        L.call("mnb_fc_gate", P(dx), P(md), 2.0, P(rr.cuda()), N * K, stream())
        torch.cuda.synchronize()
        assert rel(dx, ref_dx * mask.double() * 2.0 * (rr.double() > 0)) < 1e-4
        dw = torch.zeros(O, K, device="cuda")
        L.call("mnb_conv_wgrad", P(xb), None, None, P(dyb), P(dw), N, 1, 1, K, O, 1, 1, 0, 1, 0, 0, stream())
        db = torch.zeros(O, device="cuda")
        L.call("mnb_fc_bias_grad", P(dy.cuda()), P(db), N, O, stream())
        torch.cuda.synchronize()
        assert rel(dw, dyb.float().cpu().double().t() @ xm) < 1e-4
        assert rel(db, dy.double().sum(0)) < 1e-5
