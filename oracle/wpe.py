"""ORACLE (test infrastructure): numpy restatement of the WPE warm start of the blind sampler.

The reference calls the third-party package `nara_wpe` (testing/EulerHeunSamplerDPS.py:6-7,32-54), which is in neither
requirements.txt nor this image: PARITY UNPINNED at this boundary.  What is restated here is the published algorithm
(T. Nakatani et al., "Speech dereverberation based on variance-normalized delayed linear prediction", IEEE TASLP 18(7),
2010; L. Drude et al., "NARA-WPE", ITG 2018) with the conventions of the package's `utils.stft` / `utils.istft` /
`wpe.wpe` defaults as the reference calls them:
  stft : size 512, shift 128, periodic Blackman window, `fading` (size - shift zeros on both sides), tail zero-padded
         to whole frames, rfft per frame                                          -> (..., frames, 257) complex128
  wpe  : per bin, one channel, taps 50, delay 2, 5 iterations, statistics_mode 'full', power floor 1e-10 * max
  istft: irfft per frame * biorthogonal synthesis window (w / sum of shifted w^2), overlap-add, fading removed
"""
import numpy as np


def _window(size):
    n = np.arange(size)
    return 0.42 - 0.5 * np.cos(2 * np.pi * n / size) + 0.08 * np.cos(4 * np.pi * n / size)   # blackman(size + 1)[:-1]


def stft(x, size=512, shift=128):
    x = np.asarray(x, dtype=np.float64)
    pad = size - shift
    x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)])
    frames = -(-(x.shape[-1] - size) // shift) + 1
    x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, (frames - 1) * shift + size - x.shape[-1])])
    idx = np.arange(size)[None, :] + shift * np.arange(frames)[:, None]
    return np.fft.rfft(x[..., idx] * _window(size), n=size, axis=-1)


def istft(X, size=512, shift=128):
    w = _window(size)
    syn = w / np.tile((w ** 2).reshape(size // shift, shift).sum(0), size // shift)
    frames = X.shape[-2]
    out = np.zeros(X.shape[:-2] + (frames * shift + size - shift,))
    seg = np.fft.irfft(X, n=size, axis=-1) * syn
    for t in range(frames):
        out[..., t * shift:t * shift + size] += seg[..., t, :]
    return out[..., size - shift:out.shape[-1] - (size - shift)]


def wpe_bin(y, taps, delay, iterations):
    """y (T,) complex128: one frequency bin of one channel."""
    T = y.shape[0]
    yt = np.zeros((taps, T), dtype=np.complex128)
    for i in range(taps):
        yt[i, delay + i:] = y[:T - delay - i]
    x = y.copy()
    for _ in range(iterations):
        power = np.abs(x) ** 2
        inv = 1.0 / np.maximum(power, 1e-10 * power.max())
        R = (yt * inv) @ yt.conj().T
        P = (yt * inv) @ y.conj()
        try:
            g = np.linalg.solve(R, P)
        except np.linalg.LinAlgError:
            g = np.linalg.lstsq(R, P, rcond=None)[0]
        x = y - g.conj() @ yt
    return x


def wpe(Y, taps=50, delay=2, iterations=5):
    """Y (bins, T) complex128 -> (bins, T)."""
    return np.stack([wpe_bin(Y[f], taps, delay, iterations) for f in range(Y.shape[0])])


def wpe_dereverb(y, taps=50, delay=2, iterations=5):
    """y (N,) float -> WPE estimate (N,) — EulerHeunSamplerDPS.py:36-51 for one utterance."""
    Y = stft(y)                        # (frames, 257)
    Z = wpe(Y.T, taps, delay, iterations).T
    x = istft(Z)
    return x[:y.shape[-1]]
