"""GPU, SURVEY.md section 4 level T1 at BASELINE shapes: every one of the 27 unique ConvBlocks of the bf16 (benchmarked)
path, teacher-forced.  One bf16 forward at N=32, 224x224 leaves every block's real operands in the plan (the raw output
of the producing block + its BN scale / shift, the weights, this block's raw output and statistics); each block is then
re-checked in isolation against torch fp32 math on those same operands:
  * forward: the stored conv output, the fused BN statistics (of the stored values), the finalized scale / shift;
  * backward-data and backward-weight (and the fused depthwise backward where the plan uses it) with a bf16 dZ;
  * the BN-backward reduction + elementwise pass.
This exercises the persistent multi-tile loops, stage-ring wraps and big-tensor index paths of the kernels the bench
runs (conv_tc_k, pw_stream_k, dw_mma_*_k, dw_tile_k, stem kernels), which toy-shape tests cannot reach.  Gate: 2e-2
rel-L2 per op (north_star bf16), statistics 1e-4.  Reference call sites: src/models/mnasnet.py:48-62 (ConvBlock)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def P(t):
    return None if t is None else t.data_ptr()


def S():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _activation(plan, a):
    """fp32 NCHW value of the block's input as the kernels see it."""
    r = a.inp
    if r.nchw:
        return plan.cur_input.float()
    x = r.t.float().permute(0, 3, 1, 2)
    if r.scale is not None:
        x = torch.relu(x * r.scale[None, :, None, None] + r.shift[None, :, None, None])
    return x.contiguous()


@pytest.fixture(scope="module")
def forwarded():
    from mnb200 import engine
    from oracle import mnasnet_oracle as O
    from test_net_gpu import build
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = build("bf16")
    eng = engine.engine_for(m)
    x, t = O.synthetic_batch(32, 224, 224)
    out = m(x.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    yield eng, eng.plan(32, 224, 224)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _unique_apps(plan):
    seen, out = set(), []
    for a in plan.apps:
        if id(a.cb) not in seen:
            seen.add(id(a.cb))
            out.append(a)
    return out


def test_27_convblocks_forward_at_baseline_shapes(forwarded):
    eng, plan = forwarded
    apps = _unique_apps(plan)
    assert len(apps) == 27
    worst = 0.0
    for a in apps:
        conv, C = a.cb.conv, a.Cout
        A = _activation(plan, a)
        zr = F.conv2d(A, conv.weight.float(), None, stride=a.stride, padding=a.pad, groups=conv.groups)
        z = a.z.float().permute(0, 3, 1, 2)
        e = rel(z, zr)
        worst = max(worst, e)
        assert e < 2e-2, (a.label, a.inp.H, a.inp.C, C, e)
        off = a.stats.off // 8
        st = plan.dstats[off:off + 2 * C]
        z64 = z.double()
        torch.testing.assert_close(st[:C], z64.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(st[C:], (z64 * z64).sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
        # bn_finalize: scale = gamma / sqrt(var + eps), shift = beta - mean * scale from those statistics
        mean = z64.mean(dim=(0, 2, 3))
        var = z64.var(dim=(0, 2, 3), unbiased=False)
        bn = a.cb.bn
        sc = bn.weight.double() / torch.sqrt(var + bn.eps)
        assert rel(a.scale, sc) < 1e-4 and rel(a.invstd, 1 / torch.sqrt(var + bn.eps)) < 1e-4
        torch.testing.assert_close(a.shift.double(), bn.bias.double() - mean * sc, rtol=1e-3, atol=1e-4)
    print(f"T1 forward, 27 ConvBlocks at N=32 224^2: worst rel-L2 {worst:.2e}")


def test_27_convblocks_backward_at_baseline_shapes(forwarded):
    from mnb200 import _lib as ML
    L = ML
    eng, plan = forwarded
    worst = {"dgrad": 0.0, "wgrad": 0.0, "bn": 0.0}
    gen = torch.Generator(device="cuda").manual_seed(11)
    for a in _unique_apps(plan):
        conv, bn, r, C = a.cb.conv, a.cb.bn, a.inp, a.Cout
        M = int(a.m)
        # ---- BN backward on the real z: reduce + fused elementwise pass vs the appendix-F formulas ----
        dA = torch.randn(a.z.shape, device="cuda", generator=gen).to(BF)
        sums = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
        L.call("mnb_bn_bwd_reduce", P(dA), P(a.z), P(a.scale), P(a.shift), P(sums), M, C, 1, S())
        dz = torch.empty_like(a.z)
        dga, dbe, dbi = (torch.zeros(C, device="cuda") for _ in range(3))
        L.call("mnb_bn_bwd_apply_fused", P(dA), P(a.z), P(a.scale), P(a.shift), P(sums), P(a.mean), P(a.invstd), P(dga),
               P(dbe), P(dbi), P(dz), M, C, float(M), 1, S())
        z64 = a.z.double()
        Gm = dA.double() * ((a.z.float() * a.scale + a.shift) > 0)
        sg, sgz = Gm.sum(dim=(0, 1, 2)), (Gm * z64).sum(dim=(0, 1, 2))
        torch.testing.assert_close(sums[:C], sg, rtol=1e-5, atol=1e-3)
        dgr = a.invstd.double() * (sgz - a.mean.double() * sg)
        b = -a.scale.double() * a.invstd.double() * dgr / M
        c3 = -a.scale.double() * sg / M - b * a.mean.double()
        dzr = a.scale.double() * Gm + b * z64 + c3
        e = rel(dz, dzr)
        worst["bn"] = max(worst["bn"], e)
        assert e < 2e-2 and rel(dga, dgr) < 1e-3 and rel(dbe, sg) < 1e-3, (a.label, e)
        # ---- conv backward with that dZ ----
        A = _activation(plan, a).requires_grad_(not r.nchw)
        wt = conv.weight.detach().float().requires_grad_(True)
        zr = F.conv2d(A, wt, None, stride=a.stride, padding=a.pad, groups=conv.groups)
        gin = [wt] + ([A] if not r.nchw else [])
        grads = torch.autograd.grad(zr, gin, dz.float().permute(0, 3, 1, 2))
        dw = torch.zeros_like(conv.weight, dtype=torch.float32)
        x_t = plan.cur_input if r.nchw else r.t
        layout = ML.LAYOUT_NCHW_F32 if r.nchw else ML.LAYOUT_NHWC
        dx = None
        if a.kind == "dense":
            L.call("mnb_conv_wgrad", P(x_t), P(r.scale), P(r.shift), P(dz), P(dw), r.N, r.H, r.W, r.C, C, a.k, a.stride,
                   a.pad, 1, layout, eng.impl, S())
            if not r.nchw:
                dx = torch.full((r.N, r.H, r.W, r.C), float("nan"), device="cuda", dtype=BF)
                wpk = plan.packed[id(a.cb)][1] if id(a.cb) in plan.packed else None
                L.call("mnb_conv_dgrad_packed", P(dz), P(conv.weight), P(wpk), None, P(dx), None, None, None, None, r.N,
                       r.H, r.W, r.C, C, a.k, a.stride, a.pad, 1, eng.impl, S())
        else:
            L.call("mnb_dw_wgrad", P(x_t), P(r.scale), P(r.shift), P(dz), P(dw), r.N, r.H, r.W, r.C, a.k, 1, S())
            dx = torch.full((r.N, r.H, r.W, r.C), float("nan"), device="cuda", dtype=BF)
            L.call("mnb_dw_dgrad", P(dz), P(conv.weight), P(dx), None, None, None, None, r.N, r.H, r.W, r.C, a.k, 1, S())
            # the fused depthwise backward (what the plan runs on the big 3x3 layers) from dA directly
            dxf = torch.full_like(dx, float("nan"))
            dwf = torch.zeros_like(dw)
            ns = torch.zeros(2 * r.C, device="cuda", dtype=torch.float64)
            L.call("mnb_dw_bwd_fused", P(dA), P(a.z), P(a.scale), P(a.shift), P(sums), P(a.mean), P(a.invstd), None, None,
                   None, P(r.t), P(r.scale), P(r.shift), P(conv.weight), P(dxf), P(dwf), P(ns), r.N, r.H, r.W, r.C, a.k,
                   float(M), 1, S())
            torch.cuda.synchronize()
            assert rel(dxf.float().permute(0, 3, 1, 2), grads[1]) < 2e-2, (a.label, "fused dx")
            assert rel(dwf, grads[0]) < 2e-2, (a.label, "fused dw")
            msk = (r.t.float() * r.scale + r.shift) > 0
            torch.testing.assert_close(ns[:r.C], (dxf.double() * msk).sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-3)
        torch.cuda.synchronize()
        e = rel(dw, grads[0])
        worst["wgrad"] = max(worst["wgrad"], e)
        assert e < 2e-2, (a.label, r.H, r.C, C, "wgrad", e)
        if dx is not None:
            e = rel(dx.float().permute(0, 3, 1, 2), grads[1])
            worst["dgrad"] = max(worst["dgrad"], e)
            assert e < 2e-2, (a.label, r.H, r.C, C, "dgrad", e)
    print("T1 backward, 27 ConvBlocks at N=32 224^2: worst rel-L2", {k: f"{v:.2e}" for k, v in worst.items()})


@pytest.mark.parametrize("n,h,w", [(512, 112, 112), (256, 192, 256), (256, 256, 192), (2048, 224, 224)])
def test_full_size_configurations_run_clean(n, h, w):
    """BASELINE configs 3-5 at their full per-GPU sizes (N=2048 at 224^2 = the strong-scaling shard on ONE GPU: 2.4 GB
    tensors, offsets beyond 2^31 bytes; 112^2 ends in 4x4 maps; 192x256 / 256x192 in 6x8 / 8x6): two fused bf16 steps
    stay finite, start at chance (ln 1000) and make progress; every parameter, BN buffer and gradient is finite."""
    import math
    from mnb200 import engine
    from test_net_gpu import build
    m = build("bf16")
    eng = engine.engine_for(m)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, 3, h, w, device="cuda", generator=g)
    t = torch.randint(0, 1000, (n,), device="cuda", generator=g)
    before = eng.store.flat.clone()
    l1 = eng.train_step(x, t, lr=1e-3).item()
    l2 = eng.train_step(x, t, lr=1e-3).item()
    torch.cuda.synchronize()
    assert math.isfinite(l1) and math.isfinite(l2)
    assert abs(l1 - math.log(1000.0)) < 0.15 and l2 < l1          # random init starts at chance; Adam makes progress
    assert torch.isfinite(eng.store.flat).all() and torch.isfinite(eng.store.fbuf).all()
    assert not torch.equal(before, eng.store.flat)
    assert torch.isfinite(eng.store.grad).all() and eng.store.grad.abs().max().item() > 0
    engine.release(m)
    del eng, m
    torch.cuda.empty_cache()
