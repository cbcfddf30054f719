"""Host-side setup of the spectral kernels: DFT matrices (built in fp64, stored fp32), window envelopes,
FFT twiddles — and the four linear maps the network and the likelihood need, each with its adjoint.

  NetSTFT   — NCSNppTime.stft / .istft (networks/ncsnpp.py:464,473-496): n_fft 510, hop 128, periodic Hann,
              centre/reflect, 256 bins, frames zero-padded to a multiple of 16, inverse over ALL padded frames.
  LossSTFT  — operator.apply_stft (testing/operators/subband_filtering.py:41-52,79-80 == reverb.py:54-65,83-84):
              right-pad 512, n_fft 1024 with Hann(512)||0(512), hop 128, centre/constant, / sqrt(sum w^2).
  OperatorSTFT — the operators' own stft / istft / apply_stft / apply_istft (API-compatible operator classes).
  RirConv   — fast_apply_RIR (utils/reverb_utils.py:25-60).
"""
import math

import torch

from . import ops


def _use_fft():
    """1024-point transforms (likelihood STFT, blind operator) as shared-memory FFTs (default) or as DFT-matrix
    products (BUDDY_STFT=dft: the round-1 first implementation, kept for A/B tests).  Same linear maps."""
    import os
    return os.environ.get("BUDDY_STFT", "fft") != "dft"


def irfft_weights(bins, n_fft):
    """onesided -> real synthesis weights: 1 for DC (and Nyquist), 2 elsewhere, / n_fft."""
    a = torch.full((bins,), 2.0, dtype=torch.float64)
    a[0] = 1.0
    if n_fft % 2 == 0 and bins == n_fft // 2 + 1:
        a[-1] = 1.0
    return a / n_fft


def _dft_mats(n_fft, bins, window, k_len, device):
    """analysis [2*bins, k_len]: (w cos, -w sin); synthesis [2*bins, k_len]: irfft weights * window."""
    n = torch.arange(k_len, dtype=torch.float64)
    f = torch.arange(bins, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(f, n) / n_fft
    w = window.double()[:k_len]
    ana = torch.empty(2 * bins, k_len, dtype=torch.float64)
    ana[0::2] = torch.cos(ang) * w
    ana[1::2] = -torch.sin(ang) * w
    a = torch.full((bins,), 2.0, dtype=torch.float64)
    a[0] = 1.0
    if n_fft % 2 == 0 and bins == n_fft // 2 + 1:
        a[-1] = 1.0
    syn = torch.empty(2 * bins, k_len, dtype=torch.float64)
    syn[0::2] = (a[:, None] * torch.cos(ang)) * w / n_fft
    syn[1::2] = (-a[:, None] * torch.sin(ang)) * w / n_fft
    return ana.float().to(device).contiguous(), syn.float().to(device).contiguous()


class NetSTFT:
    N_FFT, HOP, BINS, PAD = 510, 128, 256, 255

    def __init__(self, device):
        self.device = device
        self.window = torch.hann_window(self.N_FFT, periodic=True, dtype=torch.float64)
        self.ana, self.syn = _dft_mats(self.N_FFT, self.BINS, self.window, self.N_FFT, device)
        self._env = {}

    def frames(self, n):
        return 1 + n // self.HOP

    def padded_frames(self, n):
        f = self.frames(n)
        return (f + 15) // 16 * 16

    def _inv_env(self, n):
        """1 / sum_t w^2[j - t*hop] over all padded frames, in padded-signal coordinates j."""
        if n not in self._env:
            Tp = self.padded_frames(n)
            total = (Tp - 1) * self.HOP + self.N_FFT
            env = torch.zeros(total, dtype=torch.float64)
            w2 = self.window ** 2
            for t in range(Tp):
                env[t * self.HOP:t * self.HOP + self.N_FFT] += w2
            inv = torch.where(env > 1e-11, 1.0 / env, torch.zeros_like(env))
            self._env[n] = inv.float().to(self.device)
        return self._env[n]

    def forward(self, x, scale_b=None):
        """x fp32 [B, N] -> spectrogram fp32 [B, 256, Tp, 2];  optional per-utterance input scale (EDM c_in)."""
        B, N = x.shape
        Tp = self.padded_frames(N)
        xp = torch.empty(B, N + 2 * self.PAD, device=x.device)
        ops.pad_signal(x, self.PAD, N + 2 * self.PAD, 1, xp, scale_b=scale_b)
        out = torch.empty(B, self.BINS, Tp, 2, device=x.device)
        return ops.dft_analysis(xp, self.ana, self.HOP, self.frames(N), Tp, out)

    def inverse(self, spec, n, scale_b=None):
        """spectrogram [B, 256, Tp, 2] -> fp32 [B, n]   (envelope of all Tp frames, reference quirk App. A3)."""
        B, _, Tp, _ = spec.shape
        fr = torch.empty(B, Tp, self.N_FFT, device=spec.device)
        ops.dft_synthesis(spec, self.syn, Tp, fr)
        out = torch.empty(B, n, device=spec.device)
        return ops.ola_gather(fr, self.HOP, self.PAD, n, out, tab=self._inv_env(n), scale_b=scale_b)

    def inverse_adjoint(self, g, scale_b=None):
        """adjoint of `inverse`: g fp32 [B, n] -> [B, 256, Tp, 2]."""
        B, n = g.shape
        Tp = self.padded_frames(n)
        total = (Tp - 1) * self.HOP + self.N_FFT
        gp = torch.empty(B, total, device=g.device)
        ops.pad_signal(g, self.PAD, total, 0, gp, tab=self._inv_env(n), scale_b=scale_b)
        out = torch.empty(B, self.BINS, Tp, 2, device=g.device)
        return ops.dft_analysis(gp, self.syn, self.HOP, Tp, Tp, out)

    def forward_adjoint(self, dspec, n, scale_b=None):
        """adjoint of `forward`: [B, 256, Tp, 2] -> fp32 [B, n]."""
        B = dspec.shape[0]
        F = self.frames(n)
        fr = torch.empty(B, F, self.N_FFT, device=dspec.device)
        ops.dft_synthesis(dspec, self.ana, F, fr)
        dxp = torch.empty(B, n + 2 * self.PAD, device=dspec.device)
        ops.ola_gather(fr, self.HOP, 0, n + 2 * self.PAD, dxp)
        out = torch.empty(B, n, device=dspec.device)
        return ops.reflect_fold(dxp, n, self.PAD, out, scale_b=scale_b)


class LossSTFT:
    N_FFT, WIN, HOP, BINS = 1024, 512, 128, 513

    def __init__(self, device):
        self.device = device
        w = torch.hann_window(self.WIN, dtype=torch.float64)
        norm = math.sqrt(float((w ** 2).sum()))
        if _use_fft():
            self.ana = ops.FftMat(torch.full((self.BINS,), 1.0 / norm, dtype=torch.float64), w, device)
        else:
            ana, _ = _dft_mats(self.N_FFT, self.BINS, w, self.WIN, "cpu")
            self.ana = (ana.double() / norm).float().to(device).contiguous()

    def frames(self, n):
        return 1 + (n + self.WIN) // self.HOP

    def forward(self, x):
        """x fp32 [B, N] -> [B, 513, frames, 2]."""
        B, N = x.shape
        F = self.frames(N)
        total = (F - 1) * self.HOP + self.WIN
        xp = torch.empty(B, total, device=x.device)
        ops.pad_signal(x, self.N_FFT // 2, total, 0, xp)
        out = torch.empty(B, self.BINS, F, 2, device=x.device)
        return ops.stft_analysis(xp, self.ana, self.HOP, F, F, out)

    def adjoint(self, G, n, scale_b=None):
        """adjoint of `forward`: [B, 513, frames, 2] -> fp32 [B, n]."""
        B, _, F, _ = G.shape
        fr = torch.empty(B, F, self.WIN, device=G.device)
        ops.stft_synthesis(G, self.ana, F, fr)
        out = torch.empty(B, n, device=G.device)
        return ops.ola_gather(fr, self.HOP, self.N_FFT // 2, n, out, scale_b=scale_b)


class OperatorSTFT:
    """The STFT pair both degradation operators define for themselves (testing/operators/reverb.py:54-84 ==
    subband_filtering.py:41-80): n_fft 1024 with Hann(512)||0(512), hop 128, centre / constant padding.

      stft / istft              — torch.stft / torch.istft with that window, not normalised (:79-84);
      apply_stft / apply_istft  — right-pad 512 and divide by sqrt(sum w^2) / multiply back and drop the 256-sample
                                  delay (:41-65).
    Spectra are fp32 [B, 513, frames, 2] (= view_as_real of the reference's complex tensors)."""
    N_FFT, WIN, HOP, BINS = 1024, 512, 128, 513

    def __init__(self, device):
        self.device = device
        w = torch.hann_window(self.WIN, dtype=torch.float64)
        self.norm = math.sqrt(float((w ** 2).sum()))
        aw = irfft_weights(self.BINS, self.N_FFT)
        one = torch.ones(self.BINS, dtype=torch.float64)
        self.ana = ops.FftMat(one, w, device)                    # stft
        self.ana_n = ops.FftMat(one / self.norm, w, device)      # apply_stft
        self.syn = ops.FftMat(aw, w, device)                     # istft (irfft * window)
        self.syn_n = ops.FftMat(aw * self.norm, w, device)       # apply_istft: X * sqrt(sum w^2) first
        self._env = {}

    def _inv_env(self, frames):
        """1 / OLA(window^2) over `frames` frames, padded-signal coordinates (torch.istft's envelope)."""
        if frames not in self._env:
            total = (frames - 1) * self.HOP + self.WIN
            w2 = torch.hann_window(self.WIN, dtype=torch.float64) ** 2
            env = torch.zeros(total, dtype=torch.float64)
            for t in range(frames):
                env[t * self.HOP:t * self.HOP + self.WIN] += w2
            self._env[frames] = torch.where(env > 1e-11, 1 / env, torch.zeros_like(env)).float().to(self.device)
        return self._env[frames]

    def _analysis(self, x, mat, right_pad):
        B, n = x.shape
        frames = 1 + (n + right_pad) // self.HOP
        total = (frames - 1) * self.HOP + self.WIN
        xp = torch.empty(B, total, device=x.device)
        ops.pad_signal(x, self.N_FFT // 2, total, 0, xp)
        out = torch.empty(B, self.BINS, frames, 2, device=x.device)
        return ops.stft_analysis(xp, mat, self.HOP, frames, frames, out)

    def _synthesis(self, X, mat, skip, n):
        B, _, frames, _ = X.shape
        if skip + n > (frames - 1) * self.HOP + self.WIN:
            # beyond the last frame the zero-padded window leaves no overlap-add envelope: torch.istft refuses too
            raise RuntimeError(f"istft: {frames} frames cannot produce {n} samples (window overlap add min: 1)")
        fr = torch.empty(B, frames, self.WIN, device=X.device)
        ops.stft_synthesis(X, mat, frames, fr)
        out = torch.empty(B, n, device=X.device)
        return ops.ola_gather(fr, self.HOP, skip, n, out, tab=self._inv_env(frames))

    def stft(self, x):
        return self._analysis(x, self.ana, 0)

    def apply_stft(self, x):
        return self._analysis(x, self.ana_n, self.WIN)

    def istft(self, X, length=None):
        """length None: hop * (frames - 1) samples, as torch.istft."""
        n = self.HOP * (X.shape[2] - 1) if length is None else int(length)
        return self._synthesis(X, self.syn, self.N_FFT // 2, n)

    def apply_istft(self, X, length):
        return self._synthesis(X, self.syn_n, self.N_FFT // 2 + self.WIN // 2, int(length))


class RirConv:
    """y = (x * h)[:N] by FFT, and its adjoint (correlation).  h: (M,) shared or (B, M) per utterance.

    One 2^15..2^17-point FFT when N + M - 1 fits (the 4 s case of the reference, reverb_utils.py:25-60); longer signals
    (BASELINE configs[4], 30 s) run the same kernel block-wise: overlap-add with blocks of 2^17 - M + 1 samples."""

    MAX_LOG2 = 17

    def __init__(self, h, n, device):
        h = torch.as_tensor(h, dtype=torch.float32, device=device)
        self.per_utt = h.dim() == 2
        hb = h if self.per_utt else h[None]
        m = hb.shape[-1]
        log2 = max(15, math.ceil(math.log2(n + m - 1)))
        self.block = None
        if log2 > self.MAX_LOG2:
            log2 = self.MAX_LOG2
            self.block = (1 << log2) - m + 1
            if self.block < m:
                raise ValueError(f"RirConv: RIR of {m} taps too long for block convolution with 2^{log2}-point FFTs")
        self.n, self.m, self.log2_n2, self.L = n, m, log2 - 8, 1 << log2
        k = torch.arange(256, dtype=torch.float64)
        self.tw = torch.stack([torch.cos(2 * math.pi * k / 512), -torch.sin(2 * math.pi * k / 512)], -1).float().to(device)
        self.H = torch.empty(hb.shape[0], self.L, 2, device=device)
        ops.fftconv(hb.contiguous(), m, self.log2_n2, self.tw, self.H, None, 0, 0, None, 0)

    def _apply(self, x, mode, first):
        B = x.shape[0]
        work = torch.empty(B, self.L, 2, device=x.device)     # per call: stream-ordered, safe across streams
        stride = self.L * 2 if self.per_utt else 0
        if self.per_utt:
            if first + B > self.H.shape[0]:
                raise ValueError(f"RirConv: utterances [{first}, {first + B}) but only {self.H.shape[0]} RIRs")
            H = self.H[first:first + B]
        else:
            H = self.H
        n, m = self.n, self.m
        if self.block is None:
            y = torch.empty(B, n, device=x.device)
            return ops.fftconv(x, x.shape[1], self.log2_n2, self.tw, work, H, stride, mode, y, n)
        Lb = self.block
        if mode == 1:
            # overlap-add: block k contributes (x_k * h) at offset k*Lb; consecutive blocks overlap by m - 1 samples
            y = torch.zeros(B, n, device=x.device)
            tmp = torch.empty(B, Lb + m - 1, device=x.device)
            for a in range(0, n, Lb):
                nin = min(Lb, n - a)
                nout = min(nin + m - 1, n - a)
                ops.fftconv(x[:, a:a + nin], nin, self.log2_n2, self.tw, work, H, stride, 1, tmp, nout)
                y[:, a:a + nout] += tmp[:, :nout]
            return y
        # adjoint: x_bar block k = correlation of g[k*Lb : k*Lb + Lb + m - 1] with h (blocks do not overlap in x_bar)
        out = torch.empty(B, n, device=x.device)
        for a in range(0, n, Lb):
            nout = min(Lb, n - a)
            nin = min(nout + m - 1, n - a)
            ops.fftconv(x[:, a:a + nin], nin, self.log2_n2, self.tw, work, H, stride, 2, out[:, a:a + nout], nout)
        return out

    def forward(self, x, first=0):
        """x: [B, n] = utterances first .. first+B-1 of the batch the RIRs were given for."""
        return self._apply(x, 1, first)

    def adjoint(self, g, first=0):
        return self._apply(g, 2, first)
