"""Parity of the spectral kernels against torch.stft/istft/fft (the reference's own ops) and the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("N", [3000, 8192, 65536])
def test_net_stft_roundtrip_and_adjoints(N):
    from buddy_b200.spectral import NetSTFT
    from oracle import net as onet
    st = NetSTFT("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    B = 2
    x = torch.randn(B, N, device="cuda", generator=g)
    sc = torch.tensor([0.5, 2.0], device="cuda")
    spec = st.forward(x, scale_b=sc)
    ref = onet.net_stft((x * sc[:, None])[:, None])[:, 0]
    assert spec.shape == (B, 256, ref.shape[-1], 2)  # exact frame count incl. zero padding
    assert rel(spec, torch.view_as_real(ref)) < 1e-5
    assert spec[:, :, 1 + N // 128:].abs().max().item() == 0.0
    # inverse on an arbitrary (non-consistent) spectrogram incl. non-zero padded frames
    S = torch.randn(B, 256, spec.shape[2], 2, device="cuda", generator=g)
    y = st.inverse(S, N)
    yref = onet.net_istft(torch.view_as_complex(S.contiguous())[:, None], N)[:, 0]
    assert rel(y, yref) < 1e-5
    # adjoints: <A x, S> == <x, A^T S>
    gsig = torch.randn(B, N, device="cuda", generator=g)
    lhs = (st.inverse(S, N).double() * gsig.double()).sum()
    rhs = (S.double() * st.inverse_adjoint(gsig).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    lhs = (st.forward(x).double() * S.double()).sum()
    rhs = (x.double() * st.forward_adjoint(S, N).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


@pytest.mark.parametrize("N", [4096, 65536])
def test_loss_stft_and_comp_loss(N):
    from buddy_b200 import ops
    from buddy_b200.spectral import LossSTFT
    from oracle import operators as oop
    st = LossSTFT("cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    B = 3
    y = torch.randn(B, N, device="cuda", generator=g) * 0.05
    x = (y + 0.02 * torch.randn(B, N, device="cuda", generator=g)).requires_grad_(True)
    Y, X = st.forward(y), st.forward(x.detach())
    assert X.shape == (B, 513, 1 + (N + 512) // 128, 2)
    assert rel(X, torch.view_as_real(oop.loss_stft(x.detach()))) < 1e-5
    loss = torch.empty(B, device="cuda", dtype=torch.float64)
    G = torch.empty_like(X)
    ops.comp_loss(Y, X, X.shape[2], 0.667, 512.0, loss, G)
    gx = st.adjoint(G, N)
    lref = oop.comp_loss(y, x, 512.0)
    (gref,) = torch.autograd.grad(lref.sum(), x)
    assert rel(loss.float(), lref.detach()) < 1e-4
    assert rel(gx, gref) < 1e-3, rel(gx, gref)


@pytest.mark.parametrize("N,M,per_utt", [(8192, 2000, False), (65536, 16000, False), (65536, 40000, True),
                                         (480000, 16000, True), (300001, 37710, False)])
def test_rir_fftconv(N, M, per_utt):
    """The last two cases (30 s long-form, BASELINE configs[4]) exceed one 2^17-point FFT: block-wise overlap-add."""
    from buddy_b200.spectral import RirConv
    from oracle import operators as oop
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 2
    x = torch.randn(B, N, device="cuda", generator=g)
    h = torch.randn((B, M) if per_utt else (M,), device="cuda", generator=g) * torch.exp(
        -torch.arange(M, device="cuda") / (M / 6))
    rc = RirConv(h, N, "cuda")
    y = rc.forward(x)
    if per_utt:
        ref = torch.cat([oop.fast_apply_rir(x[i:i + 1].double(), h[i].double()) for i in range(B)])
    else:
        ref = oop.fast_apply_rir(x.double(), h.double())
    assert rel(y, ref) < 1e-5, rel(y, ref)
    gy = torch.randn(B, N, device="cuda", generator=g)
    lhs = (y.double() * gy.double()).sum()
    rhs = (x.double() * rc.adjoint(gy).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_temb_philox_lincomb_rowstats():
    from buddy_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    B = 3
    t = torch.tensor([-0.3, -1.1, 0.2], device="cuda")
    Wf = torch.randn(128, device="cuda", generator=g) * 16
    emb = torch.empty(B, 256, device="cuda")
    ops.fourier_features(t, Wf, emb)
    xp = t[:, None] * Wf[None] * 2 * torch.pi
    assert rel(emb, torch.cat([xp.sin(), xp.cos()], -1)) < 1e-5
    W = torch.randn(512, 256, device="cuda", generator=g) * 0.05
    bias = torch.randn(512, device="cuda", generator=g)
    out = torch.empty(B, 512, device="cuda")
    ops.dense(emb, W, bias, out, act_in=True, act_out=True)
    ref = torch.nn.functional.silu(torch.nn.functional.linear(torch.nn.functional.silu(emb), W, bias))
    assert rel(out, ref) < 1e-5
    seeds = torch.tensor([3000, 3001, 3000], device="cuda", dtype=torch.int64)
    z = torch.empty(B, 65536, device="cuda")
    ops.philox_normal(seeds, 7, z)
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1) < 0.01
    assert torch.equal(z[0], z[2]) and not torch.equal(z[0], z[1])
    z2 = torch.empty(1, 65536, device="cuda")
    ops.philox_normal(seeds[1:2], 7, z2)
    assert torch.equal(z2[0], z[1])  # stream depends only on (seed, draw): invariant to batch sharding
    ops.philox_normal(seeds[1:2], 8, z2)
    assert not torch.equal(z2[0], z[1])
    kurt = ((z - z.mean()) ** 4).mean() / z.var() ** 2
    assert abs(kurt.item() - 3) < 0.05
    x, y_, w = (torch.randn(B, 1000, device="cuda", generator=g) for _ in range(3))
    ca, cb, cc = (torch.randn(B, device="cuda", generator=g) for _ in range(3))
    o = torch.empty_like(x)
    ops.lincomb3(o, x, ca, y_, cb, w, cc)
    assert rel(o, ca[:, None] * x + cb[:, None] * y_ + cc[:, None] * w) < 1e-6
    ops.lincomb3(o, x, ca, y_, cb)
    assert rel(o, ca[:, None] * x + cb[:, None] * y_) < 1e-6
    st = ops.row_stats(x)
    assert rel(st[:, 0], x.double().sum(1)) < 1e-9 and rel(st[:, 1], (x.double() ** 2).sum(1)) < 1e-9


@pytest.mark.parametrize("frames,K", [(517, 512), (113, 512), (7, 512), (33, 1024)])
def test_fft_stft_kernels_match_dft_matrix_form(frames, K):
    """1024-point FFT analysis / synthesis == the DFT-matrix kernels with mat = a[f] * w[n] * (cos, -sin) (and both ==
    fp64 torch): arbitrary per-bin weights a, window w of length K, ragged frame counts (16 frames per CTA)."""
    import math
    from buddy_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    B, bins, hop = 3, 513, 128
    a = torch.rand(bins, device="cuda", generator=g).double() + 0.5
    w = torch.rand(K, device="cuda", generator=g).double()
    n = torch.arange(K, device="cuda", dtype=torch.float64)
    f = torch.arange(bins, device="cuda", dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(f, n) / 1024
    mat = torch.empty(2 * bins, K, device="cuda", dtype=torch.float64)
    mat[0::2] = a[:, None] * torch.cos(ang) * w
    mat[1::2] = -a[:, None] * torch.sin(ang) * w
    fm = ops.FftMat(a.cpu(), w.cpu(), "cuda")
    L = (frames - 1) * hop + K
    sig = torch.randn(B, L, device="cuda", generator=g)
    out_f = ops.fft_analysis(sig, fm, hop, frames, frames, torch.empty(B, bins, frames, 2, device="cuda"))
    out_d = ops.dft_analysis(sig, mat.float().contiguous(), hop, frames, frames, torch.empty(B, bins, frames, 2, device="cuda"))
    fr64 = sig.double().unfold(1, K, hop)                                # [B, frames, K]
    ref = torch.einsum("mk,btk->bmt", mat, fr64).reshape(B, bins, 2, frames).permute(0, 1, 3, 2)
    assert rel(out_f, ref) < 2e-6 and rel(out_d, ref) < 1e-5, (rel(out_f, ref), rel(out_d, ref))
    S = torch.randn(B, bins, frames, 2, device="cuda", generator=g)
    fr_f = ops.fft_synthesis(S, fm, frames, torch.empty(B, frames, K, device="cuda"))
    fr_d = ops.dft_synthesis(S, mat.float().contiguous(), frames, torch.empty(B, frames, K, device="cuda"))
    ref_s = torch.einsum("bmt,mk->btk", S.double().permute(0, 1, 3, 2).reshape(B, 2 * bins, frames), mat)
    assert rel(fr_f, ref_s) < 2e-6 and rel(fr_d, ref_s) < 1e-5, (rel(fr_f, ref_s), rel(fr_d, ref_s))


def test_wpe_warm_start_vs_oracle():
    """`wpe_scaled` warm start (EulerHeunSamplerDPS.py:32-54): STFT(512/128, Blackman, fading) -> WPE (50 taps, delay 2,
    5 iterations, per bin, fp64) -> iSTFT on the GPU vs the numpy restatement of nara_wpe (oracle/wpe.py — parity
    unpinned: the package is absent, the algorithm is restated from its publication).  Ragged length, B = 2."""
    import numpy as np
    from buddy_b200 import ops
    from buddy_b200.wpe import WpeDereverb
    from oracle import wpe as ow
    n = 20517
    rng = np.random.default_rng(3)
    ys = []
    for b in range(2):
        s = np.convolve(rng.standard_normal(n), np.ones(6) / 6)[:n]
        h = rng.standard_normal(4000) * np.exp(-np.arange(4000) / (400.0 + 300 * b))
        h[0] = 1.0
        ys.append(0.05 * np.convolve(s, h)[:n])
    y = torch.tensor(np.stack(ys), dtype=torch.float32).cuda()
    wd = WpeDereverb("cuda")
    Y = wd.stft(y)
    Yo = np.stack([ow.stft(y[b].cpu().double().numpy()).T for b in range(2)])        # (B, 257, T)
    assert Y.shape[1:3] == Yo.shape[1:3]
    assert rel(Y.cpu(), torch.view_as_real(torch.from_numpy(Yo))) < 1e-5
    # the kernel alone on the oracle's own spectra (fp32 in / out, fp64 inside)
    Yin = torch.view_as_real(torch.from_numpy(Yo)).float().cuda().contiguous()
    Z = ops.wpe(Yin, 50, 2, 5)
    Zo = np.stack([ow.wpe(Yin[b].cpu().double().numpy().view(np.complex128)[..., 0], 50, 2, 5) for b in range(2)])
    e_k = rel(Z.cpu(), torch.view_as_real(torch.from_numpy(Zo)))
    x = wd(y)
    xo = torch.tensor(np.stack([ow.wpe_dereverb(y[b].cpu().double().numpy()) for b in range(2)]))
    e = rel(x.cpu(), xo)
    print(f"\n[WPE] kernel vs numpy on identical spectra {e_k:.2e}; end to end (fp32 STFT) {e:.2e}")
    assert e_k < 1e-5 and e < 1e-4 and x.shape == (2, n)
    # a silent bin / utterance must come back as zeros (singular system -> minimum-norm solution), not NaN
    Zz = ops.wpe(torch.zeros(1, 3, 64, 2, device="cuda"), 50, 2, 5)
    assert torch.equal(Zz, torch.zeros_like(Zz))


def test_wpe_scaled_initialisation():
    """initialize_x(mode = wpe_scaled) = scaling_factor * wpe(y) / std + sigma_max * noise, per utterance."""
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from buddy_b200.wpe import WpeDereverb
    from oracle import ref_harness as rh

    class _Net(torch.nn.Module):
        pass
    args = rh.make_args("blind", 2, warm="wpe_scaled")
    from buddy_b200.edm import EDM
    smp = EulerHeunSamplerDPS(_Net(), EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)), args)
    y = (torch.randn(2, 8192, generator=torch.Generator().manual_seed(5)) * 0.05).cuda()
    z = torch.randn(2, 8192, generator=torch.Generator().manual_seed(6)).cuda()
    smp.y, smp.noise_source = y, iter([z])
    t = smp.create_schedule()
    x = smp.initialize_x((2, 8192), "cuda", t)
    xp = WpeDereverb("cuda")(y)
    want = 0.05 * xp / xp.std(dim=1, keepdim=True) + float(t[0]) * z
    assert rel(x, want) < 1e-5
