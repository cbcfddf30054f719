"""ORACLE (test infrastructure): restatement of the reference's degradation operators and likelihood loss.

  * loss_stft / loss_istft  — testing/operators/subband_filtering.py:41-65,76-80 (== reverb.py:54-84)
  * comp_loss               — utils/losses.py:27-31,59-64,74-76 ("l2_comp_stft_summean"; the frequency weighting
                              is all-ones because the code reads the key `freq_weighting`, never set: losses.py:31)
  * fast_apply_rir          — utils/reverb_utils.py:25-60 (power-of-two complex-FFT convolution)
  * subband_fir, blind_degradation, time_rir, design_H, project_params
                            — testing/operators/subband_filtering.py:67-113,206-251,298-351
  * minimum_phase           — utils/reverb_utils.py:3-23
  * linear_interp_knots     — stands in for torchcde (absent, un-pinned): parity UNPINNED at this one call
                              (subband_filtering.py:233-235)
All per-utterance: a leading batch dim means B independent problems.
"""
import math

import torch
import torch.nn.functional as F

NFFT, WIN, HOP, NF_FRAMES, SR = 1024, 512, 128, 100, 16000
EQ_FREQS = [0, 125, 250, 375, 500, 625, 750, 875, 1000, 1250, 1500, 1750, 2000, 2250, 2500, 2750, 3000, 3500, 4000,
            4500, 5000, 5500, 6000, 6500, 7000, 7500, 8000]
MAX_DECAY = 6.908 / (0.1 * (SR / HOP))
MIN_DECAY = 6.908 / (2.0 * (SR / HOP))


def window_padded(device):
    return F.pad(torch.hann_window(WIN, device=device), (0, NFFT - WIN))


def _stft1024(x):
    return torch.stft(x, NFFT, hop_length=HOP, win_length=NFFT, window=window_padded(x.device), center=True,
                      onesided=True, return_complex=True, normalized=False, pad_mode="constant")


def _istft1024(X, length):
    return torch.istft(X, NFFT, hop_length=HOP, win_length=NFFT, window=window_padded(X.device), onesided=True,
                       center=True, normalized=False, return_complex=False, length=length)


def win_norm(device):
    return torch.sqrt(torch.sum(window_padded(device) ** 2))  # sqrt(192)


def loss_stft(x):
    """(B,T) -> (B,513,frames) complex: right-pad 512 zeros, centred constant-pad STFT, / sqrt(sum w^2)."""
    if x.dim() == 1:
        x = x[None]
    return _stft1024(F.pad(x, (0, WIN))) / win_norm(x.device)


def loss_istft(X, length):
    """Inverse of loss_stft up to the window overlap: * sqrt(sum w^2), istft(length + 256), drop first 256."""
    x = _istft1024(X * win_norm(X.device), length + WIN // 2)
    return x[..., WIN // 2:]


def comp_loss(x, x_hat, weight, c=0.667):
    """weight * mean_{frames}( sum_f |Xc - Xc_hat|^2 ) PER UTTERANCE -> (B,)"""
    X, Xh = loss_stft(x), loss_stft(x_hat)
    Xc = (X.abs() + 1e-8) ** c * torch.exp(1j * X.angle())
    Xhc = (Xh.abs() + 1e-8) ** c * torch.exp(1j * Xh.angle())
    return weight * torch.mean(torch.sum((Xc - Xhc).abs() ** 2, dim=-2), dim=-1)


def fast_apply_rir(x, h):
    """(B,N) * (M,) -> (B,N): y = real(ifft(fft(x,L) fft(h,L)))[:N], L = 2^ceil(log2(N+M-1))."""
    N, M = x.shape[-1], h.shape[-1]
    L = int(2 ** math.ceil(math.log2(N + M - 1)))
    return torch.fft.ifft(torch.fft.fft(x, L, dim=-1) * torch.fft.fft(h, L, dim=-1), L, dim=-1)[..., :N].real


def subband_fir(X, H):
    """Y[f,t] = sum_{n<Nf} H[f,n] X[f,t+1-n]  (one pre-impulse frame).  X (B,513,Tf) c64, H (513,Nf) c64."""
    Nf = H.shape[-1]
    pre = int((WIN // HOP) / 2) - 1
    Xp = F.pad(X, (Nf - 1 - pre, pre))
    return F.conv1d(Xp, torch.flip(H, dims=[-1]).unsqueeze(1), groups=H.shape[0])


def blind_degradation(x, H):
    squeeze = x.dim() == 1
    X = loss_stft(x)
    y = loss_istft(subband_fir(X, H), x.shape[-1])
    return y[0] if squeeze else y


def time_rir(H):
    d = torch.zeros(HOP * NF_FRAMES + 1024, device=H.device)
    d[0] = 1
    return blind_degradation(d, H)


def linear_interp_knots(vals, knots, q):
    """Piecewise-linear interpolation of vals[..., k] given at `knots` to query points q (torchcde stand-in)."""
    k = torch.clamp(torch.bucketize(q, knots) - 1, 0, len(knots) - 2)
    frac = (q - knots[k]) / (knots[k + 1] - knots[k])
    return vals[..., k] + frac * (vals[..., k + 1] - vals[..., k])


def hilbert(h):
    n = h.shape[-1]
    win = 2 * torch.heaviside(torch.linspace(-1, 1, steps=n), values=torch.ones(1)).to(h.device)
    win = torch.flip(win, dims=(-1,))
    return torch.fft.ifft(win * torch.fft.fft(h))


def minimum_phase(h):
    T = h.shape[-1]
    Hf = torch.fft.fft(F.pad(h, (0, T)))
    mag = torch.abs(Hf)
    phase = -torch.imag(hilbert(torch.log(mag + 1e-8)))
    hm = torch.real(torch.fft.ifft(mag.type(torch.complex64) * torch.exp(1j * phase)))
    return hm[:-T]


def direct_path_mag():
    h = torch.zeros(HOP * NF_FRAMES)
    h[0] = WIN / (HOP * 2)
    return _stft1024(h)[:, 1:].abs()


def design_magnitude(decays, weights):
    """A (513,Nf): exp-decay bands -> log -> linear interp over frequency -> exp -> OLA correction -> + direct path."""
    dev = decays.device
    n = torch.arange(NF_FRAMES, device=dev).float()
    D = torch.zeros(len(EQ_FREQS), NF_FRAMES, device=dev)
    D[1:-1] = (weights.unsqueeze(-1) * torch.exp(decays).unsqueeze(-1) ** (-n[None, None, :])).sum(0)
    L = torch.log(D.transpose(0, 1) + 1e-6)  # (Nf, 27)
    freqs = torch.fft.rfftfreq(NFFT, d=1 / SR).to(dev)
    A = torch.exp(linear_interp_knots(L, torch.tensor(EQ_FREQS, dtype=torch.float32, device=dev), freqs))
    A = A.transpose(0, 1) + 1e-6  # (513, Nf)
    w = torch.hann_window(WIN, device=dev)
    K = int(WIN / HOP - 1)
    cols = [A[:, k] / (w.sum() / w[(K - k) * HOP:].sum()) for k in range(K)]
    A = torch.cat([torch.stack(cols, dim=1), A[:, K:]], dim=1)
    return A + direct_path_mag().to(dev)


def consistency(H):
    """`cons` (subband_filtering.py:333-351): istft -> min-phase -> h[0]=2 -> stft, (513,Nf) -> (513,Nf)."""
    L = H.shape[-1]
    h = _istft1024(F.pad(H, (1, 1)), HOP * NF_FRAMES)
    h = F.pad(h, (0, HOP))
    h = minimum_phase(h)
    h = torch.cat([torch.full((1,), WIN / (HOP * 2), device=h.device, dtype=h.dtype), h[1:]])
    return _stft1024(h)[:, 1:-1][..., :L]


def design_H(decays, weights, phases):
    return consistency(design_magnitude(decays, weights) * torch.exp(1j * phases))


def project_params(decays, weights):
    """subband_filtering.py:298-331 with num_exponentials == 1."""
    return decays.clamp(MIN_DECAY, MAX_DECAY), weights.clamp(10 ** (0 / 20), 10 ** (40 / 20))
