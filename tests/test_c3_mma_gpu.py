"""GPU: the bulk-copy + mma.sync 3x3 kernels (csrc/c3_mma.cu) behind mnb_conv_fwd / mnb_conv_dgrad / mnb_conv_wgrad for the
stride-2 stage transitions (src/models/mnasnet.py:130-161 via ConvBlock :48-62), against torch fp32 math on the same bf16
operands: forward + BN statistics of the stored values, backward-data, backward-weight.  Shapes cover full-size maps,
ragged last row blocks, odd heights, tiny maps, non-square maps, packed / unpacked weights, with / without the fused
BN-apply+ReLU of the producing block.  Gate 1e-2 rel-L2 (bf16 operands, fp32 accumulate; north_star bf16 tolerance is 2e-2)."""
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


CASES = [(3, 112, 112, 16, 24), (3, 56, 56, 24, 40), (3, 28, 28, 40, 80), (2, 30, 28, 16, 24), (2, 13, 12, 24, 40),
         (2, 64, 96, 16, 24), (5, 6, 6, 40, 80), (2, 48, 64, 24, 40), (1, 2, 2, 16, 24), (2, 96, 128, 16, 24)]


@pytest.mark.parametrize("N,H,W,Cin,Cout", CASES)
@pytest.mark.parametrize("packed,act", [(True, True), (False, False)])
def test_c3_mma_matches_torch(N, H, W, Cin, Cout, packed, act):
    from mnb200 import _lib as L
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device="cuda").manual_seed(3 + Cin + 7 * Cout + H)
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        x = (torch.randn(N, H, W, Cin, device="cuda", generator=g) * 0.8 + 0.1).to(BF)
        dz = torch.randn(N, Ho, Wo, Cout, device="cuda", generator=g).to(BF)
        w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)).float()
        sc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float()
        sh = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
        wpf = torch.empty(w.numel(), device="cuda", dtype=BF)
        wpd = torch.empty(w.numel(), device="cuda", dtype=BF)
        L.call("mnb_pack_weights", P(w), P(wpf), P(wpd), Cout, Cin, 3, S())
        z = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda", dtype=BF)
        st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
        dx = torch.full((N, H, W, Cin), float("nan"), device="cuda", dtype=BF)
        dw = torch.zeros_like(w)
        a_sc, a_sh = (P(sc), P(sh)) if act else (None, None)
        assert L.get_option("c3_mma") == 1
        L.call("mnb_conv_fwd_packed", P(x), a_sc, a_sh, P(w), P(wpf) if packed else None, None, P(z), P(st), N, H, W, Cin,
               Cout, 3, 2, 1, 1, 0, 0, S())
        L.call("mnb_conv_dgrad_packed", P(dz), P(w), P(wpd) if packed else None, None, P(dx), None, None, None, None, N, H, W,
               Cin, Cout, 3, 2, 1, 1, 0, S())
        L.call("mnb_conv_wgrad", P(x), a_sc, a_sh, P(dz), P(dw), N, H, W, Cin, Cout, 3, 2, 1, 1, 0, 0, S())
        torch.cuda.synchronize()
        A = x.float()
        if act:
            A = torch.relu(A * sc + sh).to(BF).float()
        A = A.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        wb = w.to(BF).float().requires_grad_(True)
        zr = F.conv2d(A, wb, None, stride=2, padding=1)
        gA, gw = torch.autograd.grad(zr, [A, wb], dz.float().permute(0, 3, 1, 2))
        assert not torch.isnan(z.float()).any() and not torch.isnan(dx.float()).any()
        assert rel(z.float().permute(0, 3, 1, 2), zr) < 1e-2
        assert rel(dx.float().permute(0, 3, 1, 2), gA) < 1e-2
        assert rel(dw, gw) < 1e-2
        zs = z.double()
        torch.testing.assert_close(st[:Cout], zs.sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(st[Cout:], (zs * zs).sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_c3_mma_agrees_with_the_tcgen05_path():
    """Same operands through impl 0 (c3_mma) and impl 2 (tcgen05): the two kernels agree to bf16 rounding."""
    from mnb200 import _lib as L
    N, H, W, Cin, Cout = 4, 56, 56, 16, 24
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(BF)
    dz = torch.randn(N, H // 2, W // 2, Cout, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / 12).float()
    sc = (torch.rand(Cin, device="cuda", generator=g) + 0.5).float()
    sh = (torch.randn(Cin, device="cuda", generator=g) * 0.3).float()
    res = []
    for impl in (0, 2):
        z = torch.empty(N, H // 2, W // 2, Cout, device="cuda", dtype=BF)
        st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
        dx = torch.empty(N, H, W, Cin, device="cuda", dtype=BF)
        dw = torch.zeros_like(w)
        L.call("mnb_conv_fwd", P(x), P(sc), P(sh), P(w), None, P(z), P(st), N, H, W, Cin, Cout, 3, 2, 1, 1, 0, impl, S())
        L.call("mnb_conv_dgrad", P(dz), P(w), None, P(dx), None, None, None, None, N, H, W, Cin, Cout, 3, 2, 1, 1, impl, S())
        L.call("mnb_conv_wgrad", P(x), P(sc), P(sh), P(dz), P(dw), N, H, W, Cin, Cout, 3, 2, 1, 1, 0, impl, S())
        torch.cuda.synchronize()
        res.append((z.float(), st, dx.float(), dw))
    for a, b in zip(*res):
        assert rel(a, b) < 5e-3


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("N,H,W", [(2, 224, 224), (2, 64, 96), (1, 32, 48), (3, 30, 28), (2, 17, 16), (1, 448, 448)])
def test_stem_forward_on_the_tensor_pipe(N, H, W, u8):
    """features.0 forward in bf16 mode (csrc/stem_mma.cu: bulk-copied rows, im2col tile, mma.sync) for fp32 NCHW and uint8
    NHWC input against torch fp32 math on the bf16-rounded operands; shapes whose rows cannot be bulk-copied (W*3 % 16)
    or that are too wide (448) fall back to the CUDA-core kernel, which must agree to bf16 rounding (src/models/mnasnet.py:179)."""
    from mnb200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(H + W)
    w = (torch.randn(32, 3, 3, 3, device="cuda", generator=g) / 5).float()
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda")
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if u8:
        x8 = torch.randint(0, 256, (N, H, W, 3), device="cuda", generator=g, dtype=torch.uint8)
        xn = ((x8.permute(0, 3, 1, 2).float() / 255) - mean[None, :, None, None]) / std[None, :, None, None]
        args = (P(x8), P(mean), P(std))
        layout = L.LAYOUT_NHWC_U8
    else:
        xn = torch.randn(N, 3, H, W, device="cuda", generator=g)
        args = (P(xn), None, None)
        layout = L.LAYOUT_NCHW_F32
    res = []
    old = L.get_option("stem_mma")
    try:
        for opt in (1, 0):
            L.set_option("stem_mma", opt)
            z = torch.full((N, Ho, Wo, 32), float("nan"), device="cuda", dtype=BF)
            st = torch.zeros(64, device="cuda", dtype=torch.float64)
            L.call("mnb_conv_fwd", args[0], args[1], args[2], P(w), None, P(z), P(st), N, H, W, 3, 32, 3, 2, 1, 1, layout, 0, S())
            torch.cuda.synchronize()
            res.append((z, st))
    finally:
        L.set_option("stem_mma", old)
    z, st = res[0]
    zr = F.conv2d(xn.to(BF).float(), w.to(BF).float(), None, stride=2, padding=1)
    assert not torch.isnan(z.float()).any()
    assert rel(z.float().permute(0, 3, 1, 2), zr) < 5e-3
    zs = z.double()
    torch.testing.assert_close(st[:32], zs.sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(st[32:], (zs * zs).sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)
    assert rel(z.float(), res[1][0].float()) < 1e-2            # the CUDA-core kernel keeps fp32 operands
