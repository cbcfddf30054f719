"""WPE warm start on the GPU (`warm_initialization.mode: "wpe_scaled"`, testing/EulerHeunSamplerDPS.py:32-54).

The reference leaves the device for this step: numpy STFT (nara_wpe.utils.stft: size 512, shift 128, periodic Blackman
window, `fading` zero padding of size - shift on both sides, tail padded to whole frames) -> nara_wpe.wpe.wpe
(50 taps, delay 2, 5 iterations) -> nara_wpe.utils.istft (biorthogonal synthesis window).  Here: the same three
stages as buddy_b200 kernels (DFT-matrix STFT / iSTFT shared with the network transform, `buddy_wpe`), batched over
utterances — every utterance is its own one-channel problem (the reference's batch axis would be WPE *channels*:
another B = 1 artefact).  nara_wpe is not installed anywhere in this project: the algorithm is
restated from its publication in oracle/wpe.py (parity unpinned at this boundary, stated in DESIGN.md).
"""
import torch

from . import ops
from .spectral import _dft_mats


class WpeDereverb:
    SIZE, SHIFT, BINS = 512, 128, 257

    def __init__(self, device, taps=50, delay=2, iterations=5):
        self.device = torch.device(device)
        self.taps, self.delay, self.iterations = int(taps), int(delay), int(iterations)
        n = torch.arange(self.SIZE, dtype=torch.float64)
        w = 0.42 - 0.5 * torch.cos(2 * torch.pi * n / self.SIZE) + 0.08 * torch.cos(4 * torch.pi * n / self.SIZE)
        # biorthogonal synthesis window: w / sum over the size/shift overlapping shifts of w^2
        den = (w ** 2).view(self.SIZE // self.SHIFT, self.SHIFT).sum(0).repeat(self.SIZE // self.SHIFT)
        self.ana, _ = _dft_mats(self.SIZE, self.BINS, w, self.SIZE, self.device)
        _, self.syn = _dft_mats(self.SIZE, self.BINS, w / den, self.SIZE, self.device)

    def frames(self, n):
        pad = self.SIZE - self.SHIFT
        return -(-(n + 2 * pad - self.SIZE) // self.SHIFT) + 1

    def stft(self, x):
        B, n = x.shape
        T = self.frames(n)
        total = (T - 1) * self.SHIFT + self.SIZE
        xp = torch.empty(B, total, device=x.device)
        ops.pad_signal(x, self.SIZE - self.SHIFT, total, 0, xp)
        return ops.dft_analysis(xp, self.ana, self.SHIFT, T, T, torch.empty(B, self.BINS, T, 2, device=x.device))

    def istft(self, Z, n):
        B, _, T, _ = Z.shape
        fr = torch.empty(B, T, self.SIZE, device=Z.device)
        ops.dft_synthesis(Z, self.syn, T, fr)
        return ops.ola_gather(fr, self.SHIFT, self.SIZE - self.SHIFT, n, torch.empty(B, n, device=Z.device))

    def __call__(self, y):
        """y fp32 [B, n] -> WPE estimate of the dry signal [B, n]."""
        y = y.float().contiguous()
        Z = ops.wpe(self.stft(y), self.taps, self.delay, self.iterations)
        return self.istft(Z, y.shape[1])
