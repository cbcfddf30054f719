"""Host-side composition of the spectral path, checked on the CPU against the oracle.

The product's spectral classes (`buddy_b200.spectral`, `.operators.RIROperator`, `.functional`) are Python compositions
of C-ABI calls: which padding, how many frames, which overlap-add offset / envelope, which blocks of a long RIR
convolution.  These tests swap plain-torch stand-ins for those entry points into `buddy_b200.ops`
(tests/emulated_kernels.py — test infrastructure, each restating the contract documented in include/buddy_b200.h) and
hold the compositions to the oracle (`oracle/`, pinned to the reference) — so an indexing slip in the host code shows
up in the `-m "not gpu"` suite, not only on the GPU box.  The kernels themselves are covered by the `-m gpu` tests.
"""

import pytest
import torch

from oracle import net as onet
from oracle import operators as oop


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.fixture()
def emu(monkeypatch):
    from buddy_b200 import ops
    import emulated_kernels as ek          # tests/ is on sys.path (pytest rootdir import mode)
    for name, fn in ek.ALL.items():
        monkeypatch.setattr(ops, name, fn)
    return ops


@pytest.mark.parametrize("n", [8192, 3000, 20517])
def test_network_stft_pair_and_adjoints(emu, n):
    """NCSNppTime.stft / .istft (networks/ncsnpp.py:473-496): frame count, zero frames up to a multiple of 16, inverse
    over ALL padded frames, and the two adjoints the data-gradient uses."""
    from buddy_b200.spectral import NetSTFT
    st = NetSTFT("cpu")
    x = randn(n, 2, n) * 0.1
    want = torch.view_as_real(onet.net_stft(x[:, None])[:, 0])              # [B, 256, Tp, 2]
    got = st.forward(x)
    assert got.shape == want.shape and got.shape[2] == st.padded_frames(n) and got.shape[2] % 16 == 0
    assert rel(got, want) < 1e-5
    assert torch.equal(got[:, :, st.frames(n):], torch.zeros_like(got[:, :, st.frames(n):]))   # padded frames: exact 0
    spec = randn(n + 1, *want.shape) * 0.1
    want_x = onet.net_istft(torch.view_as_complex(spec)[:, None], n)[:, 0]
    assert rel(st.inverse(spec, n), want_x) < 1e-5
    # <A x, S> = <x, A^T S> and <A^-1 S, g> = <S, (A^-1)^T g>
    g = randn(n + 2, 2, n)
    assert abs(dot(st.forward(x), spec) - dot(x, st.forward_adjoint(spec, n))) < 1e-4 * abs(dot(st.forward(x), spec))
    assert abs(dot(st.inverse(spec, n), g) - dot(spec, st.inverse_adjoint(g))) < 1e-4 * abs(dot(st.inverse(spec, n), g))
    # per-utterance scales (EDM c_in / c_out) ride on the padding / gather kernels
    sc = torch.tensor([0.5, 3.0])
    assert rel(st.forward(x, scale_b=sc), want * sc[:, None, None, None]) < 1e-5


@pytest.mark.parametrize("n", [8192, 5000])
def test_likelihood_stft_and_operator_transforms(emu, n):
    """operator.apply_stft / apply_istft / stft / istft (reverb.py:54-84 == subband_filtering.py:41-80)."""
    from buddy_b200.spectral import LossSTFT, OperatorSTFT
    x = randn(n, 2, n) * 0.1
    ls, tf = LossSTFT("cpu"), OperatorSTFT("cpu")
    want = torch.view_as_real(oop.loss_stft(x))
    for got in (ls.forward(x), tf.apply_stft(x)):
        assert got.shape == want.shape and got.shape[2] == ls.frames(n) and rel(got, want) < 1e-5
    G = randn(n + 1, *want.shape)
    assert abs(dot(ls.forward(x), G) - dot(x, ls.adjoint(G, n))) < 1e-4 * abs(dot(ls.forward(x), G))
    X = oop._stft1024(x)
    assert rel(tf.stft(x), torch.view_as_real(X)) < 1e-5
    full = 128 * (X.shape[-1] - 1)
    for length in (None, full, full - 300):
        want_x = oop._istft1024(X, length)
        got_x = tf.istft(torch.view_as_real(X).contiguous(), length)
        # the last samples of a full-length inverse are divided by w^2[511] = 1.4e-9: compare away from them
        assert got_x.shape == want_x.shape and rel(got_x[:, :full - 300], want_x[:, :full - 300]) < 1e-5
    with pytest.raises(RuntimeError):                    # torch.istft refuses too (no overlap-add envelope out there)
        tf.istft(torch.view_as_real(X).contiguous(), full + 1)
    Xa = oop.loss_stft(x)
    assert rel(tf.apply_istft(torch.view_as_real(Xa).contiguous(), n), oop.loss_istft(Xa, n)) < 1e-5


def test_rir_convolution_single_fft_block_ola_and_adjoint(emu):
    """fast_apply_RIR (utils/reverb_utils.py:25-60): one FFT when N + M - 1 fits 2^17 points, otherwise block-wise
    overlap-add (30 s utterances); shared and per-utterance RIRs; the adjoint (correlation) of both paths."""
    from buddy_b200.spectral import RirConv
    for n, m in ((8192, 2000), (300000, 16000)):
        h = randn(1, m) * torch.exp(-6.908 * torch.arange(m) / m)
        x, g = randn(2, 3, n), randn(3, 3, n)
        rc = RirConv(h, n, "cpu")
        assert (rc.block is None) == (n + m - 1 <= 1 << 17)
        y = rc.forward(x)
        assert rel(y, oop.fast_apply_rir(x, h)) < 1e-5
        assert abs(dot(y, g) - dot(x, rc.adjoint(g))) < 1e-4 * abs(dot(y, g))
        hb = torch.stack([h, h.flip(0), 0.5 * h])
        rb = RirConv(hb, n, "cpu")
        want = torch.stack([oop.fast_apply_rir(x[b:b + 1], hb[b])[0] for b in range(3)])
        assert rel(rb.forward(x), want) < 1e-5
        assert rel(rb.forward(x[1:], first=1), want[1:]) < 1e-5          # micro-batch offset into the RIR table
        assert abs(dot(want, g) - dot(x, rb.adjoint(g))) < 1e-4 * abs(dot(want, g))
        with pytest.raises(ValueError):
            rb.forward(x, first=1)                                       # utterances [1, 4) of 3 RIRs
    with pytest.raises(ValueError):                                      # RIR longer than a block
        RirConv(torch.zeros(70000), 300000, "cpu")


def test_rir_operator_and_function_mirrors(emu):
    """buddy_b200.operators.RIROperator (reverb.py:8-87) and buddy_b200.functional (reverb_utils.py:3-60,
    losses.py:17-95) on top of the same compositions."""
    from buddy_b200 import functional as F
    from buddy_b200.operators import RIROperator
    from oracle.ref_harness import AD, op_hp
    n = 8192
    h = randn(10, 2000) * torch.exp(-6.908 * torch.arange(2000) / 6000.0)
    h[37] = 3.0
    x = randn(11, 2, n) * 0.05
    op = RIROperator(op_hp(), time_kernel_size=2000, sample_rate=16000)
    op.update_params(h)
    y = oop.fast_apply_rir(x, h)
    assert rel(op.degradation(x), y) < 1e-5 and rel(op.degradation(x[0]), y[0]) < 1e-5
    y_cut = oop.fast_apply_rir(x, h[37:])
    assert rel(op.degradation(x, rm_delay=True), y_cut) < 1e-5           # reverb_utils.py:27-28
    assert rel(op.degradation(x), y) < 1e-5                              # and back: the cached plan follows the flag
    assert rel(F.fast_apply_RIR(x, h), y) < 1e-5 and rel(F.fast_apply_RIR(x, h, rm_delay=True, zero_pad=True), y_cut) < 1e-5
    assert abs(float(op.optim_fwd(x, 0.9 * y)) - float(((y - 0.9 * y) ** 2).sum())) < 1e-5 * float(((0.1 * y) ** 2).sum())
    assert rel(torch.view_as_real(op.apply_stft(x)), torch.view_as_real(oop.loss_stft(x))) < 1e-5
    assert rel(torch.view_as_real(op.stft(x[0])), torch.view_as_real(oop._stft1024(x[0]))) < 1e-5
    Xa = oop.loss_stft(x)
    assert rel(op.apply_istft(Xa, length=n), oop.loss_istft(Xa, n)) < 1e-5
    # loss factory: the reference reduces over the batch axis too (losses.py:48-67)
    x_hat = x + 0.01 * randn(12, 2, n)
    per = lambda w: oop.comp_loss(x, xr, w)                              # per-utterance `summean`, weight w
    bins, frames = 513, 1 + (n + 512) // 128
    for name, w, red in (("l2_comp_stft_summean", 512.0, lambda v: v.mean()), ("l2_comp_stft_sum", 3.0, lambda v: v.sum() * frames),
                         ("l2_comp_stft_mean", 7.0, lambda v: v.mean() / bins)):
        xr = x_hat.clone().requires_grad_(True)
        want = red(per(w))
        (gw,) = torch.autograd.grad(want, xr)
        xo = x_hat.clone().requires_grad_(True)
        got = F.get_loss(AD(name=name, weight=w, compression_factor=0.667), operator=op)(x, xo)
        (go,) = torch.autograd.grad(got, xo)
        assert abs(got.item() - want.item()) < 1e-4 * abs(want.item()) and rel(go, gw) < 1e-4, name
    # minimum-phase helpers at the blind operator's size
    hm = randn(13, 2, 12928) * torch.exp(-torch.arange(12928) / 2000.0)
    got = F.minimum_phase_version(hm)
    for b in range(2):
        assert rel(got[b], oop.minimum_phase(hm[b])) < 1e-5
    z = randn(14, 25856)
    assert rel(torch.view_as_real(F.hilbert(z)), torch.view_as_real(oop.hilbert(z))) < 1e-5
