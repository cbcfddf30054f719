"""Parity of the HBM-bound glue kernels against plain PyTorch fp32 (same ops as the reference's
GroupNorm/SiLU/resample/concat: networks/ncsnpp_utils/layerspp.py:242-263)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from buddy_b200 import ops
    return ops


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def nchw(x):
    return x.permute(0, 3, 1, 2)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def resample(x, mode):  # NCHW
    if mode == 1:
        return x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    if mode == 2:
        return F.avg_pool2d(x, 2)
    return x


CASES = [(2, 8, 12, 128, 0, 0, True), (2, 8, 12, 256, 0, 1, True), (1, 16, 20, 256, 0, 2, True),
         (2, 8, 12, 256, 128, 0, True), (1, 8, 12, 256, 256, 0, True), (2, 6, 10, 256, 0, 0, False),
         (1, 8, 8, 256, 128, 1, True)]


@pytest.mark.parametrize("B,H,W,Ca,Cb,mode,silu", CASES)
def test_gn_apply_and_bwd(B, H, W, Ca, Cb, mode, silu):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    C = Ca + Cb
    xa = torch.randn(B, H, W, Ca, device="cuda", generator=g) * 1.5 + 0.3
    xb = (torch.randn(B, H, W, Cb, device="cuda", generator=g) * 0.7 - 0.2) if Cb else None
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    sa = ops.gn_stats(xa)
    sb = ops.gn_stats(xb) if Cb else None
    Ho, Wo = (H * 2, W * 2) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
    out = torch.empty(B, Ho, Wo, C, device="cuda", dtype=torch.float16)
    raw = torch.empty_like(out)
    ops.gn_apply(xa, sa, gamma, beta, out, xb=xb, sb=sb, silu=silu, mode=mode, out_raw=raw)

    x = torch.cat([xa, xb], -1) if Cb else xa
    xr = nchw(x).clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gamma, beta, eps=1e-6)
    y = F.silu(y) if silu else y
    y = resample(y, mode)
    assert rel(out.float(), nhwc(y.detach())) < 1e-3
    assert rel(raw.float(), nhwc(resample(nchw(x), mode))) < 1e-3

    # backward
    da = torch.randn(B, Ho, Wo, C, device="cuda", generator=g)
    dskip = torch.randn(B, Ho, Wo, C, device="cuda", generator=g)
    extra_a = torch.randn(B, H, W, Ca, device="cuda", generator=g)
    extra_b = torch.randn(B, H, W, Cb, device="cuda", generator=g) if Cb else None
    (gx,) = torch.autograd.grad(y, xr, nchw(da))
    ref = nhwc(gx) + nhwc(_pull(nchw(dskip), mode)) * 0.7
    ref_a = ref[..., :Ca] + extra_a
    gsum = torch.empty(B, 32, 2, device="cuda", dtype=torch.float64)
    dxa = torch.empty(B, H, W, Ca, device="cuda")
    g16a = torch.empty(B, H, W, Ca, device="cuda", dtype=torch.float16)
    dxb = torch.empty(B, H, W, Cb, device="cuda") if Cb else None
    g16b = torch.empty(B, H, W, Cb, device="cuda", dtype=torch.float16) if Cb else None
    ops.gn_bwd(xa, sa, gamma, beta, da, gsum, xb=xb, sb=sb, silu=silu, mode=mode, dskip=dskip, skip_scale=0.7,
               extra_a=extra_a, extra_b=extra_b, dxa=dxa, dxb=dxb, g16a=g16a, g16b=g16b, g16_scale=0.5)
    assert rel(dxa, ref_a) < 2e-5, rel(dxa, ref_a)
    assert rel(g16a.float(), ref_a * 0.5) < 1e-3
    if Cb:
        ref_b = ref[..., Ca:] + extra_b
        assert rel(dxb, ref_b) < 2e-5
        assert rel(g16b.float(), ref_b * 0.5) < 1e-3


def _pull(t, mode):  # adjoint of resample, NCHW
    if mode == 1:
        return F.avg_pool2d(t, 2) * 4
    if mode == 2:
        return t.repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.25
    return t


def test_im2col_col2im_c2():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(8)
    B, H, W = 2, 12, 20
    x = torch.randn(B, H, W, 2, device="cuda", generator=g)
    col = torch.empty(B, H, W, 64, device="cuda", dtype=torch.float16)
    ops.im2col_c2(x, col)
    ref = F.unfold(nchw(x), 3, padding=1).view(B, 2, 9, H, W).permute(0, 3, 4, 2, 1).reshape(B, H, W, 18)
    assert rel(col[..., :18].float(), ref) < 1e-3 and col[..., 18:].abs().max().item() == 0
    dcol = torch.randn(B, H, W, 32, device="cuda", generator=g)
    dx = torch.empty(B, H, W, 2, device="cuda")
    ops.col2im_c2(dcol, dx)
    xr = nchw(x).clone().requires_grad_(True)
    cols = F.unfold(xr, 3, padding=1).view(B, 2, 9, H, W).permute(0, 3, 4, 2, 1).reshape(B, H, W, 18)
    (gx,) = torch.autograd.grad(cols, xr, dcol[..., :18].contiguous())
    assert rel(dx, nhwc(gx)) < 1e-5


def test_resample_combine_affine():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, W = 2, 8, 12
    x = torch.randn(B, H, W, 2, device="cuda", generator=g)
    o = torch.empty(B, H // 2, W // 2, 2, device="cuda")
    assert rel(ops.resample_c2(x, 0, o), nhwc(F.avg_pool2d(nchw(x), 2))) < 1e-6
    add = torch.randn(B, 2 * H, 2 * W, 2, device="cuda", generator=g)
    o2 = torch.empty(B, 2 * H, 2 * W, 2, device="cuda")
    assert rel(ops.resample_c2(x, 1, o2, add=add), nhwc(resample(nchw(x), 1)) + add) < 1e-6
    assert rel(ops.resample_c2(x, 2, o2), nhwc(_pull(nchw(x), 2))) < 1e-6
    assert rel(ops.resample_c2(x, 3, o), nhwc(_pull(nchw(x), 1))) < 1e-6
    C = 256
    h = torch.randn(B, H, W, C, device="cuda", generator=g)
    w = torch.randn(C, 2, device="cuda", generator=g)
    bias = torch.randn(C, device="cuda", generator=g)
    out = torch.empty_like(h)
    ops.combine_fwd(h, x, w, bias, out)
    assert rel(out, h + x @ w.t() + bias) < 1e-6
    dp = torch.empty(B, H, W, 2, device="cuda")
    ops.combine_bwd(h, w, dp)
    assert rel(dp, h @ w) < 1e-5
    y = torch.empty_like(x)
    ops.affine_c2(x, [1.0, 2.0, -0.5, 0.25], [0.1, -0.2], y)
    m = torch.tensor([[1.0, 2.0], [-0.5, 0.25]], device="cuda")
    assert rel(y, x @ m.t() + torch.tensor([0.1, -0.2], device="cuda")) < 1e-6


def test_softmax_transpose_cast():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(10)
    B, n = 2, 300
    s = torch.randn(B, n, n, device="cuda", generator=g) * 3
    p = torch.empty(B, n, n, device="cuda", dtype=torch.float16)
    ops.softmax_fwd(s, p)
    pr = torch.softmax(s, -1)
    assert rel(p.float(), pr) < 1e-3
    dp = torch.randn(B, n, n, device="cuda", generator=g)
    ds = torch.empty(B, n, n, device="cuda", dtype=torch.float16)
    ops.softmax_bwd(p, dp, 0.0625, ds)
    pf = p.float()
    ref = pf * (dp - (dp * pf).sum(-1, keepdim=True)) * 0.0625
    assert rel(ds.float(), ref) < 2e-3
    x = torch.randn(B, 100, 768, device="cuda", generator=g).half()
    out = torch.empty(B, 256, 100, device="cuda", dtype=torch.float16)
    ops.transpose_h(x[..., 256:512], out)
    assert torch.equal(out, x[..., 256:512].transpose(1, 2).contiguous())
    xf = torch.randn(1024, device="cuda", generator=g)
    yh = torch.empty(1024, device="cuda", dtype=torch.float16)
    ops.cast_scale_h(xf, 0.5, yh)
    assert rel(yh.float(), xf * 0.5) < 1e-3


def test_gn_act32():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(13)
    for C, silu in ((128, True), (256, False), (384, True)):
        x = torch.randn(2, 6, 10, C, device="cuda", generator=g) * 1.3 + 0.2
        gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
        beta = 0.1 * torch.randn(C, device="cuda", generator=g)
        out = ops.gn_act32(x, ops.gn_stats(x), gamma, beta, torch.empty_like(x), silu=silu)
        y = F.group_norm(nchw(x), 32, gamma, beta, eps=1e-6)
        y = F.silu(y) if silu else y
        assert rel(out, nhwc(y)) < 1e-5


def test_cast_operand():
    """Raw fp32 tensor -> tensor-core operand (fp16, fp16 hi|lo, fp16 + e4m3 pair), optional nearest x2 upsampling."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(12)
    B, H, W, C = 2, 6, 10, 128
    x = torch.randn(B, H, W, C, device="cuda", generator=g)
    for up in (False, True):
        ref = nhwc(resample(nchw(x), 1)) * 0.5 if up else x * 0.5
        Ho, Wo = ref.shape[1:3]
        o = ops.cast_operand(x, torch.empty(B, Ho, Wo, C, device="cuda", dtype=torch.float16), scale=0.5, upsample=up)
        assert torch.equal(o, ref.half())
        o2 = ops.cast_operand(x, torch.empty(B, Ho, Wo, 2 * C, device="cuda", dtype=torch.float16), scale=0.5,
                              upsample=up, split=1)
        assert torch.equal(o2[..., :C], ref.half()) and rel(o2[..., :C].float() + o2[..., C:].float(), ref) < 1e-6
        o8 = torch.empty(B, Ho, Wo, 2 * C, device="cuda", dtype=torch.uint8)
        o3 = ops.cast_operand(x, torch.empty(B, Ho, Wo, C, device="cuda", dtype=torch.float16), o8, scale=0.5,
                              upsample=up, split=2)
        assert torch.equal(o3, ref.half())
        # e4m3 pair = [e4m3(lo * 2^9) | e4m3(hi)], the layout gn_apply writes (elementwise.cu store_op4)
        hi = ref.half().float()
        f8 = o8.view(torch.float8_e4m3fn).float()
        assert torch.equal(f8[..., C:], hi.to(torch.float8_e4m3fn).float())
        assert torch.equal(f8[..., :C], ((ref - hi) * 512.0).to(torch.float8_e4m3fn).float())


def test_upfirdn2d_vs_reference_fixture_and_oracle():
    """sm_100a upfirdn2d (the reference's only native operator, op/upfirdn2d_kernel.cu:107-207) against outputs of the
    reference's own CPU branch (tests/golden/upfirdn2d.pt), its data-gradient against autograd through the oracle, the
    channels-last ([major][h][w][minor > 1]) form of the C entry point, and the FIR resampling helpers."""
    import os
    from buddy_b200 import upfirdn2d as bu
    from oracle import upfirdn as ou
    rn = lambda seed, *s: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "upfirdn2d.pt"), weights_only=False)
    for c in gold:
        x, k = rn(c["x_seed"], *c["shape"]).cuda(), rn(c["k_seed"], c["taps"], c["taps"]).cuda()
        xr = x.clone().requires_grad_(True)
        y = bu.UpFirDn2d.apply(xr, k, c["up"], c["down"], c["pad"])
        assert y.shape == c["out"].shape and rel(y.detach().cpu(), c["out"]) < 1e-6
        cot = torch.randn_like(y)
        (gx,) = torch.autograd.grad((y * cot).sum(), xr)
        xo = x.detach().cpu().double().requires_grad_(True)         # oracle in fp64 on the CPU (no TF32 convolutions)
        (go,) = torch.autograd.grad((ou.upfirdn2d(xo, k.cpu().double(), c["up"], c["down"], c["pad"])
                                     * cot.cpu().double()).sum(), xo)
        assert rel(gx.cpu(), go) < 1e-6
    # channels-last data, the layout of the engine's activations: [major = B][H][W][minor = C]
    x4 = rn(1, 2, 12, 10, 16).cuda()
    k = rn(2, 4, 4).cuda()
    y4 = bu._launch(x4.contiguous(), k, (2, 2), (1, 1), (2, 1, 2, 1))
    ref = ou.upfirdn2d(x4.cpu().double().permute(0, 3, 1, 2), k.cpu().double(), (2, 2), (1, 1),
                       (2, 1, 2, 1)).permute(0, 2, 3, 1)
    assert rel(y4.cpu(), ref) < 1e-6
    # StyleGAN2 resampling helpers (up_or_down_sampling.py:195-256) with the reference's FIR [1, 3, 3, 1]
    x = torch.ones(1, 4, 16, 20, device="cuda")
    up, dn = bu.upsample_2d(x, [1, 3, 3, 1]), bu.downsample_2d(x, [1, 3, 3, 1])
    assert up.shape == (1, 4, 32, 40) and dn.shape == (1, 4, 8, 10)
    assert (up[..., 2:-2, 2:-2] - 1).abs().max() < 1e-6 and (dn[..., 1:-1, 1:-1] - 1).abs().max() < 1e-6   # unit DC gain
