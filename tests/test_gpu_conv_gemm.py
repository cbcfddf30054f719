"""Parity of the tcgen05 implicit-GEMM conv / batched GEMM kernel against plain PyTorch fp32
(the same op the reference runs: nn.Conv2d, networks/ncsnpp_utils/layers.py:100-126)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from buddy_b200 import ops
    return ops


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def ref_conv(a16, w16, taps):
    """a16: [B,H,W,C] fp16, w16: [T,N,K] fp16 -> [B,H,W,N] fp32 (exact products of the fp16 operands)."""
    x = a16.float().permute(0, 3, 1, 2).double()
    T, N, K = w16.shape
    if taps == 9:
        w = w16.double().view(3, 3, N, K).permute(2, 3, 0, 1)
        y = F.conv2d(x, w, padding=1)
    else:
        w = w16.double().view(N, K, 1, 1)
        y = F.conv2d(x, w)
    return y.permute(0, 2, 3, 1).float()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("M,K,N", [(128, 64, 16), (1000, 128, 256), (2112, 256, 768), (4096, 512, 384)])
def test_plain_gemm(M, K, N):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(1, 1, M, K, device="cuda", generator=g).half()
    w = torch.randn(1, N, K, device="cuda", generator=g).half()
    out = torch.full((1, 1, M, N), float("nan"), device="cuda")
    ops.conv_gemm(a, w, out, taps=1, n_total=N)
    torch.cuda.synchronize()
    ref = ref_conv(a, w, 1)
    assert rel(out, ref) < 1e-5, rel(out, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 16), (2, 32, 66, 64, 128), (1, 64, 132, 128, 256),
                                            (2, 16, 40, 256, 2), (1, 24, 24, 64, 384)])
def test_conv3x3(B, H, W, Cin, Cout):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(B, H, W, Cin, device="cuda", generator=g).half()
    rows = max(Cout, 16)
    w = torch.zeros(9, rows, Cin, device="cuda", dtype=torch.float16)
    w[:, :Cout] = (torch.randn(9, Cout, Cin, device="cuda", generator=g) * 0.1).half()
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    ops.conv_gemm(a, w, out, taps=9, n_total=Cout)
    torch.cuda.synchronize()
    ref = ref_conv(a, w[:, :Cout].contiguous(), 9)
    assert rel(out, ref) < 1e-5, rel(out, ref)


def test_fused_epilogue_and_skip_conv():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, C1, C2, N = 2, 16, 24, 128, 64, 128
    a = torch.randn(B, H, W, C1, device="cuda", generator=g).half()
    a2 = torch.randn(B, H, W, C2, device="cuda", generator=g).half()
    w = (torch.randn(9, N, C1, device="cuda", generator=g) * 0.05).half()
    w2 = (torch.randn(N, C2, device="cuda", generator=g) * 0.1).half()
    bias = torch.randn(N, device="cuda", generator=g)
    bias_b = torch.randn(B, N, device="cuda", generator=g)
    resid = torch.randn(B, H, W, N, device="cuda", generator=g)
    out = torch.empty(B, H, W, N, device="cuda")
    stats = torch.zeros(B, N // 4, 2, device="cuda", dtype=torch.float64)
    ops.conv_gemm(a, w, out, taps=9, n_total=N, a2=a2, w2=w2, bias=bias, bias_b=bias_b, resid=resid,
                  scale=0.70710678, stats=stats)
    torch.cuda.synchronize()
    ref = ref_conv(a, w, 9) + ref_conv(a2, w2[None], 1) + bias + bias_b[:, None, None, :] + resid
    ref = ref * 0.70710678
    assert rel(out, ref) < 1e-5, rel(out, ref)
    o = out.double().view(B, H * W, N // 4, 4)
    s_ref = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1)
    assert rel(stats, s_ref) < 1e-6, rel(stats, s_ref)


def test_batched_b_fp16_out_strided():
    """Attention-style: per-image B operand, fp16 output written into a column window of a wider buffer."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    B, M, K, N = 3, 300, 128, 200
    qkv = torch.randn(B, 1, M, 3 * K, device="cuda", generator=g).half()
    q = qkv[..., :K]
    kmat = torch.randn(B, N, K, device="cuda", generator=g).half()
    out = torch.zeros(B, 1, M, 512, device="cuda", dtype=torch.float16)
    ops.conv_gemm(q, kmat, out, taps=1, n_total=N, b_batched=True, col_off=64, ldc=512, scale=0.125)
    torch.cuda.synchronize()
    ref = torch.einsum("bmk,bnk->bmn", q[:, 0].double(), kmat.double()) * 0.125
    got = out[:, 0, :, 64:64 + N].double()
    assert rel(got, ref) < 1e-3
    assert out[..., :64].abs().max().item() == 0 and out[..., 64 + N:].abs().max().item() == 0


def test_big_conv_speed():
    """256->256 3x3 at the full 256x528 resolution (the single largest layer)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, C = 4, 256, 528, 256
    a = torch.randn(B, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(9, C, C, device="cuda", generator=g) * 0.02).half()
    out = torch.empty(B, H, W, C, device="cuda")
    for _ in range(2):
        ops.conv_gemm(a, w, out, taps=9, n_total=C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 5
    for _ in range(iters):
        ops.conv_gemm(a, w, out, taps=9, n_total=C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * B * H * W * C * C * 9 / (ms * 1e-3) / 1e12
    print(f"\n[conv 256->256 @256x528 B={B}] {ms:.3f} ms  {tf:.1f} TFLOP/s")
    ref = ref_conv(a[:1, :64], w, 9)
    # interior rows only (the reference slab was cut at row 64, so its last row sees a different halo)
    assert rel(out[:1, :63], ref[:, :63]) < 1e-5


def test_fp8_corrected_products():
    """fp16 main pass + e4m3 correction passes (scale-input-d fold) reproduce fp32 products to ~1e-5."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    B, H, W, C, N = 2, 16, 24, 128, 256
    a = torch.randn(B, H, W, C, device="cuda", generator=g) * 1.3
    w = torch.randn(9, N, C, device="cuda", generator=g) * 0.04
    a_hi = a.half()
    a_lo = a - a_hi.float()
    w_hi = w.half()
    w_lo = w - w_hi.float()
    e4 = lambda t: t.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    a8 = torch.cat([e4(a_lo * 512.0), e4(a_hi.float())], dim=-1).contiguous()
    w8 = torch.cat([e4(w_hi.float() * 32.0), e4(w_lo * 16384.0)], dim=-1).contiguous()
    out = torch.empty(B, H, W, N, device="cuda")
    ops.conv_gemm(a_hi, w_hi, out, taps=9, n_total=N, a8=a8, w8=w8)
    x = a.permute(0, 3, 1, 2).double()
    ref = F.conv2d(x, w.double().view(3, 3, N, C).permute(2, 3, 0, 1), padding=1).permute(0, 2, 3, 1).float()
    single = torch.empty_like(out)
    ops.conv_gemm(a_hi, w_hi, single, taps=9, n_total=N)
    e_c8, e_1 = rel(out, ref), rel(single, ref)
    print(f"\n[conv precision] fp16 single pass {e_1:.2e}   fp16 + e4m3 corrections {e_c8:.2e}")
    assert e_1 > 1e-4 and e_c8 < 3e-5


@pytest.mark.parametrize("B,H,W,C,N,with_res", [(2, 16, 24, 64, 128, True), (3, 19, 37, 64, 256, True),
                                                (1, 33, 50, 128, 64, False), (2, 8, 130, 64, 384, True),
                                                (1, 5, 7, 64, 32, True)])
def test_staged_epilogue_matches_direct(B, H, W, C, N, with_res):
    """TMA-store epilogue (smem-staged chunks, column-wise statistics, TMA-loaded residual) against the direct
    register->global epilogue on ragged image sizes (tiles that hang over the border) — must agree to fp32 rounding in
    the output and to accumulation order in the statistics."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(B, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(9, N, C, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    bias_b = torch.randn(B, N, device="cuda", generator=g)
    resid = torch.randn(B, H, W, N, device="cuda", generator=g) if with_res else None
    outs, sts = [], []
    for direct in (True, False):
        out = torch.full((B, H, W, N), float("nan"), device="cuda")
        stats = torch.zeros(B, N // 4, 2, device="cuda", dtype=torch.float64)
        for _ in range(2):      # twice: the second launch checks barrier phases / buffer parity survive a relaunch
            stats.zero_()
            ops.conv_gemm(a, w, out, taps=9, n_total=N, bias=bias, bias_b=bias_b, resid=resid, scale=0.5, stats=stats,
                          direct_epilogue=direct)
        outs.append(out)
        sts.append(stats)
    torch.cuda.synchronize()
    # same products; the bias terms are summed in a different order (bias row is pre-added in the staged path)
    assert not torch.isnan(outs[1]).any()
    assert (outs[0] - outs[1]).abs().max().item() < 2e-6 * outs[0].abs().max().item()
    assert rel(sts[1], sts[0]) < 1e-6
    ref = (ref_conv(a, w, 9) + bias + bias_b[:, None, None, :] + (resid if with_res else 0.0)) * 0.5
    assert rel(outs[1], ref) < 1e-5


@pytest.mark.parametrize("B,H,W,C,N,c8", [(2, 16, 24, 64, 128, False), (3, 19, 37, 64, 256, True),
                                          (1, 8, 8, 128, 64, False), (1, 40, 50, 128, 384, True),
                                          (5, 8, 16, 64, 32, False), (1, 256, 528, 64, 128, True)])
def test_cta_pairs_match_single_cta(B, H, W, C, N, c8):
    """cta_group::2 (two CTAs share one 256-row tile pair, each loads half the weight tile) against the single-CTA
    path: same products in the same order -> identical output; odd pixel-tile counts exercise the padding tile."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(8)
    a = torch.randn(B, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(9, N, C, device="cuda", generator=g) * 0.05).half()
    a2 = torch.randn(B, H, W, 64, device="cuda", generator=g).half()
    w2 = (torch.randn(N, 64, device="cuda", generator=g) * 0.1).half()
    bias = torch.randn(N, device="cuda", generator=g)
    bias_b = torch.randn(B, N, device="cuda", generator=g)
    kw = {}
    if c8:
        e4 = lambda t: t.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
        kw = dict(a8=e4(torch.randn(B, H, W, 2 * C, device="cuda", generator=g)),
                  w8=e4(torch.randn(9, N, 2 * C, device="cuda", generator=g)),
                  a8_2=e4(torch.randn(B, H, W, 128, device="cuda", generator=g)),
                  w8_2=e4(torch.randn(N, 128, device="cuda", generator=g)))
    outs, sts = [], []
    # single CTA; CTA pairs (default: a kernel row of three weight tiles per pipeline stage where it fits); pairs
    # with one tile per stage — the same MMAs in the same order every time
    for no_pairs, one_tap in ((True, False), (False, False), (False, True)):
        out = torch.full((B, H, W, N), float("nan"), device="cuda")
        stats = torch.zeros(B, N // 4, 2, device="cuda", dtype=torch.float64)
        for _ in range(2):
            stats.zero_()
            ops.conv_gemm(a, w, out, taps=9, n_total=N, a2=a2, w2=w2, bias=bias, bias_b=bias_b, scale=0.5,
                          stats=stats, no_pairs=no_pairs, one_tap_per_stage=one_tap, **kw)
        torch.cuda.synchronize()
        outs.append(out)
        sts.append(stats)
    assert not torch.isnan(outs[1]).any()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel(sts[1], sts[0]) < 1e-9 and rel(sts[2], sts[0]) < 1e-9


@pytest.mark.parametrize("B,H,W,C,N", [(2, 16, 24, 64, 128), (3, 19, 37, 128, 256), (1, 40, 50, 64, 384)])
def test_fused_groupnorm_backward_statistics(B, H, W, C, N):
    """The dgrad epilogue's fused GroupNorm-backward statistics (pass 0 of buddy_gn_bwd) give the same input gradient as
    the two-pass kernel: conv -> da, then dx = GN/SiLU backward of x under da, with and without the fused pass."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(B, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(9, N, C, device="cuda", generator=g) * 0.05).half()
    x = torch.randn(B, H, W, N, device="cuda", generator=g) * 1.5 + 0.3
    gamma = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    sx = ops.gn_stats(x)
    outs = []
    for fused in (False, True):
        da = torch.empty(B, H, W, N, device="cuda")
        gsum = torch.zeros(B, 32, 2, device="cuda", dtype=torch.float64)
        ops.conv_gemm(a, w, da, taps=9, n_total=N, scale=0.37,
                      gnb=(x, sx, gamma, beta, gsum, 32, 1e-6, 1) if fused else None)
        dx = torch.empty_like(x)
        ops.gn_bwd(x, sx, gamma, beta, da, gsum, silu=True, dxa=dx, pass0_done=fused)
        outs.append((da, gsum.clone(), dx))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel(outs[1][1], outs[0][1]) < 1e-5, rel(outs[1][1], outs[0][1])
    assert rel(outs[1][2], outs[0][2]) < 1e-6, rel(outs[1][2], outs[0][2])


@pytest.mark.parametrize("B,H,W,C,N,c8,with_res", [(2, 32, 24, 64, 128, False, True), (3, 19, 37, 64, 128, True, True),
                                                   (1, 33, 50, 128, 64, False, False), (2, 50, 9, 64, 32, True, True),
                                                   (1, 256, 528, 128, 128, True, False)])
def test_stacked_tiles_match_single_tile(B, H, W, C, N, c8, with_res):
    """Plain 3x3 launches with N <= 128 stack two 16x8-pixel tiles per CTA (every weight tile feeds two accumulators
    from one (32+2)-row halo patch) — same products in the same order as one tile per CTA: identical output, including
    heights that are not a multiple of 32 (the lower half hangs over the border) and the residual / statistics path."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(B, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(9, N, C, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    bias_b = torch.randn(B, N, device="cuda", generator=g)
    res = torch.randn(B, H, W, N, device="cuda", generator=g) if with_res else None
    kw = {}
    if c8:
        e4 = lambda t: t.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
        kw = dict(a8=e4(torch.randn(B, H, W, 2 * C, device="cuda", generator=g)),
                  w8=e4(torch.randn(9, N, 2 * C, device="cuda", generator=g)))
    outs = []
    for single in (True, False):
        out = torch.full((B, H, W, N), float("nan"), device="cuda")
        stats = torch.zeros(B, N // 4, 2, device="cuda", dtype=torch.float64)
        ops.conv_gemm(a, w, out, taps=9, n_total=N, bias=bias, bias_b=bias_b, resid=res, scale=0.7, stats=stats,
                      single_tile=single, **kw)       # True: never stacked, False: stacked whenever legal
        torch.cuda.synchronize()
        outs.append((out, stats))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel(outs[1][1], outs[0][1]) < 1e-12
    if not c8:
        ref = (ref_conv(a, w, 9) + bias + bias_b[:, None, None, :] + (res if with_res else 0)) * 0.7
        assert rel(outs[1][0], ref) < 1e-5
