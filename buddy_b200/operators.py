"""API-compatible degradation operators (reference testing/operators/reverb.py:8-87,
testing/operators/subband_filtering.py:8-351, testing/operators/shared.py:5-28).

The samplers accept the reference's own operator objects (they only read `.params`, `.params_phases`, `.H`,
`.op_hp`); these classes exist so the path can be used and tested without the reference on sys.path — same
constructor arguments, method names, argument meaning, return shapes / dtypes and error behaviour, every method a
sequence of buddy_b200 kernels (CUDA tensors only, no CPU fallback).  Differences by design:
  * methods are forward evaluations — autograd does not record them (the samplers use the hand-written adjoints in
    blind.py / spectral.py);
  * `apply_istft` does not scale its argument in place (reference quirk, subband_filtering.py:61).
"""
import math

import torch

from . import ops
from .spectral import OperatorSTFT, RirConv

EQ_FREQS = [0, 125, 250, 375, 500, 625, 750, 875, 1000, 1250, 1500, 1750, 2000, 2250, 2500, 2750, 3000, 3500, 4000,
            4500, 5000, 5500, 6000, 6500, 7000, 7500, 8000]


def _hp(op_hp, key, default=None):
    if op_hp is None:
        return default
    if isinstance(op_hp, dict):
        return op_hp.get(key, default)
    return getattr(op_hp, key, default)


class Operator(torch.nn.Module):
    """testing/operators/shared.py:5-28."""

    def degradation(self, *args, **kwargs):
        raise NotImplementedError

    def update_params(self, *args, **kwargs):
        raise NotImplementedError

    def prepare_optimization(self, x_den, y):
        return x_den, y

    def constrain_params(self):
        pass


class RIROperator(Operator):
    """Informed operator: convolution with a known room impulse response (reverb.py:8-87)."""

    def __init__(self, op_hp=None, time_kernel_size=10, sample_rate=16000):
        super().__init__()
        self.op_hp = op_hp
        self.time_kernel_size = time_kernel_size
        self.sample_rate = sample_rate
        self.params = None
        self.n_fft = int(_hp(op_hp, "NFFT", 1024))
        self.win_length = int(_hp(op_hp, "win_length", 512))
        self.hop_length = int(_hp(op_hp, "hop", 128))
        window = _hp(op_hp, "window", "hann")
        if window != "hann":
            raise NotImplementedError("window type {} not implemented".format(window))
        if (self.n_fft, self.win_length, self.hop_length) != (1024, 512, 128):
            raise NotImplementedError("the CUDA kernels implement NFFT 1024 / win_length 512 / hop 128 "
                                      "(the shipped op_hp) only")
        self._conv, self._conv_key = None, None
        self._tf = None

    def update_params(self, k, **ignored):
        self.params = torch.as_tensor(k)
        self._conv = None

    def degradation(self, x, rm_delay=False, **ignored):
        assert self.params is not None, "filter is None"
        squeeze = x.dim() == 1
        x2 = (x[None] if squeeze else x).float().contiguous()
        key = (x2.shape[1], bool(rm_delay))
        if self._conv is None or self._conv_key != key:
            h = self.params.to(x2.device)
            if rm_delay:                                   # reverb_utils.py:27-28: start at the strongest tap
                h = h[int(torch.argmax(h)):]
            self._conv, self._conv_key = RirConv(h, x2.shape[1], x2.device), key
        y = self._conv.forward(x2)
        return y[0] if squeeze else y

    def optim_fwd(self, Xden, Y):
        """sum (A(Xden) - Y)^2 (:43-50)."""
        d = self.degradation(Xden) if Xden.dim() == 2 else self.degradation(Xden)[None]
        Y2 = (Y if Y.dim() == 2 else Y[None]).float().contiguous()
        B = d.shape[0]
        one = torch.ones(B, device=d.device)
        e = ops.lincomb3(torch.empty_like(d), d.contiguous(), one, Y2, -one)
        return ops.row_stats(e)[:, 1].sum().float()

    # ---- transforms "just for computing STFT-based losses" (:52-84)
    def _transforms(self, device):
        if self._tf is None:
            self._tf = OperatorSTFT(device)
        return self._tf

    @staticmethod
    def _as_batch(x):
        if x.dim() == 1:
            return x[None]
        if x.dim() == 2:
            return x
        raise ValueError("x must have shape (batch, samples) or (samples)")

    def apply_stft(self, x):
        x2 = self._as_batch(x).float().contiguous()
        return torch.view_as_complex(self._transforms(x2.device).apply_stft(x2))

    def apply_istft(self, X, length=None):
        if length is None:
            raise ValueError("apply_istft needs `length` (the reference warns that istft may crash without it)")
        X3 = X if X.dim() == 3 else X[None]
        Xr = torch.view_as_real(X3.to(torch.complex64).contiguous()).contiguous()
        x = self._transforms(Xr.device).apply_istft(Xr, int(length))
        return x if X.dim() == 3 else x[0]

    def stft(self, x):
        x2 = (x[None] if x.dim() == 1 else x).float().contiguous()
        X = torch.view_as_complex(self._transforms(x2.device).stft(x2))
        return X[0] if x.dim() == 1 else X

    def istft(self, X, length=None):
        X3 = X if X.dim() == 3 else X[None]
        Xr = torch.view_as_real(X3.to(torch.complex64).contiguous()).contiguous()
        x = self._transforms(Xr.device).istft(Xr, length)
        return x if X.dim() == 3 else x[0]

    def get_time_RIR(self):
        return self.params


class SubbandFiltering(Operator):
    """Informed sub-band filter operator (subband_filtering.py:8-113): y = iSTFT(H (*) STFT(x)) per frequency bin."""

    def __init__(self, op_hp=None, sample_rate=16000, device="cuda"):
        super().__init__()
        from .blind import BlindEngine
        self.H = None
        self.sample_rate = sample_rate
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("buddy_b200 operators run on CUDA only (no CPU fallback)")
        self.op_hp = op_hp
        self.n_fft = int(_hp(op_hp, "NFFT", 1024))
        self.win_length = int(_hp(op_hp, "win_length", 512))
        self.hop_length = int(_hp(op_hp, "hop", 128))
        self.Nf = int(_hp(op_hp, "Nf", 100))
        window = _hp(op_hp, "window", "hann")
        if window != "hann":
            raise NotImplementedError("window type {} not implemented".format(window))
        if (self.n_fft, self.win_length, self.hop_length, self.Nf) != (1024, 512, 128, 100):
            raise NotImplementedError("the CUDA kernels implement NFFT 1024 / win_length 512 / hop 128 / Nf 100 "
                                      "(the shipped op_hp) only")
        self.window = torch.hann_window(self.win_length, device=self.device)
        self.window_padded = torch.nn.functional.pad(self.window, (0, self.n_fft - self.win_length))
        self.freqs = torch.fft.rfftfreq(self.n_fft, d=1 / sample_rate).to(self.device)
        self.length_rir = self.hop_length * self.Nf
        self.time = torch.arange(self.Nf, dtype=torch.float32) / (self.sample_rate / self.hop_length)
        self._eng = BlindEngine(self.length_rir, self.device, op_hp=op_hp, sample_rate=sample_rate)
        z = torch.zeros(1, 25, device=self.device)
        self._eng.init_state(1, z, z, torch.zeros(513, 100, device=self.device),
                             torch.zeros(513, 100, dtype=torch.complex64, device=self.device))
        self._eng.select(slice(0, 1))

    # ---- transforms -----------------------------------------------------------------------------------
    @staticmethod
    def _as_batch(x):
        if x.dim() == 1:
            return x[None], True
        if x.dim() == 2:
            return x, False
        raise ValueError("x must have shape (batch, samples) or (samples)")

    def apply_stft(self, x):
        x2, _ = self._as_batch(x)
        return torch.view_as_complex(self._eng.loss_stft.forward(x2.float().contiguous()))

    def apply_istft(self, X, length=None):
        if length is None:
            raise ValueError("apply_istft needs `length` (the reference warns that istft may crash without it)")
        Xr = torch.view_as_real(X.to(torch.complex64).contiguous()).contiguous()
        return self._eng.apply_istft(Xr, int(length))

    def stft(self, x):
        """torch.stft(n_fft 1024, hop 128, window hann(512)||0, center, constant pad), NOT normalised (:79-80)."""
        x2, squeeze = self._as_batch(x)
        x2 = x2.float().contiguous()
        B, n = x2.shape
        frames = 1 + n // self.hop_length
        total = (frames - 1) * self.hop_length + self.win_length
        xp = torch.empty(B, total, device=x2.device)
        ops.pad_signal(x2, self.n_fft // 2, total, 0, xp)
        out = torch.empty(B, self.n_fft // 2 + 1, frames, 2, device=x2.device)
        ops.stft_analysis(xp, self._eng.cons_ana, self.hop_length, frames, frames, out)
        X = torch.view_as_complex(out)
        return X[0] if squeeze else X

    def istft(self, X, length=None):
        """torch.istft with the same window, NOT normalised (:76-77); length None: hop * (frames - 1) samples."""
        X3 = X if X.dim() == 3 else X[None]
        Xr = torch.view_as_real(X3.to(torch.complex64).contiguous()).contiguous()
        B, _, frames, _ = Xr.shape
        n = self.hop_length * (frames - 1) if length is None else int(length)
        if n > self.hop_length * (frames - 1):
            # beyond the last frame the zero-padded window leaves no overlap-add envelope: torch.istft refuses too
            raise RuntimeError(f"istft: {frames} frames cannot produce {n} samples (window overlap add min: 1)")
        fr = torch.empty(B, frames, self.win_length, device=Xr.device)
        ops.stft_synthesis(Xr, self._eng.cons_syn, frames, fr)
        x = ops.ola_gather(fr, self.hop_length, self.n_fft // 2, n, torch.empty(B, n, device=Xr.device),
                           tab=self._eng._inv_env(frames))
        return x if X.dim() == 3 else x[0]

    def subband_filtering(self, X, H):
        Xr = torch.view_as_real(X.to(torch.complex64).contiguous()).contiguous()
        Hr = torch.view_as_real(H.to(torch.complex64).contiguous()).contiguous()
        shared = Hr.dim() == 3          # one filter for every utterance of the batch (the reference's only case)
        Y = ops.subband_fir(Xr, Hr, torch.empty_like(Xr), Nf=self.Nf, pre=1, mode=0, shared_h=shared)
        return torch.view_as_complex(Y)

    def degradation(self, x, mode="waveform", H=None, detach_operator=False):
        init_shape = x.shape
        X = self.apply_stft(x)
        if H is None:
            assert self.H is not None, "filter is not initialized"
            H = self.H
        Y = self.subband_filtering(X, H.detach())
        if mode == "waveform":
            y = self.apply_istft(Y, length=init_shape[-1])
            return y.squeeze(0) if len(init_shape) == 1 else y
        if mode == "STFT":
            return Y
        raise ValueError(mode)

    def get_time_RIR(self, excitation=None, H=None):
        if excitation is None:
            x = torch.zeros(int(self.length_rir + 1024), device=self.device)
            x[0] = 1
        else:
            x = torch.as_tensor(excitation, dtype=torch.float32, device=self.device)
        return self.degradation(x, H=H)

    def update_H(self, rir=None, H=None):
        if rir is not None:
            Hn = self.stft(torch.as_tensor(rir, dtype=torch.float32, device=self.device))
            Hn = Hn * ((8) / (self.win_length / self.hop_length))
            Hn = Hn[:, 1:]
            if self.Nf > Hn.shape[-1]:
                Hn = torch.cat((Hn, torch.zeros(Hn.shape[0], self.Nf - Hn.shape[-1], dtype=Hn.dtype,
                                                device=Hn.device)), -1)
            else:
                Hn = Hn[..., 0:self.Nf]
            self.H = Hn
        elif H is not None:
            self.H = H
        else:
            raise ValueError("Either rir or H must be specified. This is the informed scenario, so we need to know "
                             "the filter")
        assert self.H.shape[0] == self.n_fft // 2 + 1 and self.H.shape[1] == self.Nf, "H.shape: {}".format(self.H.shape)


class BlindSubbandFiltering(SubbandFiltering):
    """Blind operator: sub-band filters parameterised by per-band exponential decays, weights and free phases
    (subband_filtering.py:116-351).  State lives in `params = [decays (1,25), weights (1,25)]`,
    `params_phases = [phases (513,100)]` and `H` (513,100) complex, exactly the attributes the samplers read and
    write back."""

    def __init__(self, op_hp=None, sample_rate=16000, magnitude_distance=True, H_cplx=False, device="cuda"):
        super().__init__(op_hp, sample_rate, device=device)
        self.Amin, self.Amax = _hp(op_hp, "Amin", 0), _hp(op_hp, "Amax", 40)
        self.EQ_freqs = torch.tensor(_hp(op_hp, "EQ_freqs", EQ_FREQS), dtype=torch.float32, device=self.device)
        self.fix_EQ_extremes = bool(_hp(op_hp, "fix_EQ_extremes", True))
        self.fix_direct_path = bool(_hp(op_hp, "fix_direct_path", True))
        if len(self.EQ_freqs) != 27 or not self.fix_EQ_extremes or not self.fix_direct_path or \
                not bool(_hp(op_hp, "minimum_phase", True)) or not bool(_hp(op_hp, "clamp_decay", True)) or \
                bool(_hp(op_hp, "strictly_decreasing_decay", False)):
            raise NotImplementedError("the CUDA kernels implement the shipped op_hp (27 EQ knots with fixed extremes, "
                                      "minimum phase, fixed direct path, clamped non-monotone decays) only")
        self.num_bands = len(self.EQ_freqs) - 2
        ip = _hp(op_hp, "init_params", None)
        t60 = list(_hp(ip, "T60_breakpoints", [0.1]))
        wts = list(_hp(ip, "multiexp_weighting", [2]))
        if _hp(op_hp, "init_single_value", True):
            t60 = [self.num_bands * [v] for v in t60]
            wts = [self.num_bands * [v] for v in wts]
        t60 = torch.tensor(t60, dtype=torch.float32, device=self.device)
        wts = torch.tensor(wts, dtype=torch.float32, device=self.device)
        assert len(wts) == len(t60), "multiexp_weighting must have the same length as T60_breakpoints"
        if t60.shape[0] != 1:
            raise NotImplementedError("num_exponentials > 1 is not on the hot path (shipped config: one exponential)")
        assert t60.shape[-1] == self.num_bands and wts.shape[-1] == self.num_bands, \
            "T60_breakpoints must have the same length as EQ_freqs-2"
        self.num_exponentials = 1
        self.params_decay = 6.908 / (t60 * (self.sample_rate / self.hop_length))
        self.params_decay_weighting = wts
        self.max_decay, self.min_decay = self._eng.max_decay, self._eng.min_decay
        self.phases = torch.rand((self.n_fft // 2 + 1, self.Nf), dtype=torch.float32).to(self.device) * 2 * math.pi \
            - math.pi
        self.params = [self.params_decay, self.params_decay_weighting]
        self.params_phases = [self.phases]
        self.direct_path_mag_correction = self._eng.tabs["dpmag"]
        init = _hp(op_hp, "init_phases", "random_coherent")
        if init == "random_coherent":
            self.update_H(use_noise=True)
        elif init == "random":
            self.update_H()
        else:
            raise NotImplementedError("This is not implemented yet")

    def _load_state(self, phases):
        st = self._eng.state
        st["decays"].copy_(self.params[0].detach().reshape(1, 25))
        st["weights"].copy_(self.params[1].detach().reshape(1, 25))
        st["phases"].copy_(phases.detach().reshape(1, 513, 100))

    def compute_direct_path_mag_correction(self):
        """|STFT(h)|[:, 1:] of h = (win_length / (2 hop)) * delta (:206-210)."""
        h = torch.zeros((self.length_rir,), device=self.device)
        h[0] = 1 * (self.win_length / (self.hop_length * 2))
        self.direct_path_mag_correction = self.stft(h)[:, 1:].abs()

    def correct_OLA(self, A, inverse=False):
        """Divide (inverse: multiply) the first win/hop - 1 frames by sum(w) / sum(w[(K - k) hop:]), in place
        (:212-222)."""
        corr = self._eng.tabs["corr"]
        K = corr.numel()
        if inverse:
            A[:, :K] *= corr
        else:
            A[:, :K] /= corr
        return A

    def design_filter(self, correct_OLA=True):
        """Magnitudes A (513,100): exponential decays per band -> log-linear interpolation over frequency -> OLA
        correction of the first frames + direct-path magnitude (:224-251)."""
        self._load_state(self.params_phases[0])
        bf = self._eng.buf
        ops.blind_design_fwd(self._eng.state["decays"], self._eng.state["weights"], self._eng.state["phases"],
                             self._eng.tabs, bf["A"], bf["H0"])
        A = bf["A"][0].clone()
        if not correct_OLA:      # the kernel fuses the correction: take it back out of the first frames
            A = self.correct_OLA(A - self.direct_path_mag_correction, inverse=True) + self.direct_path_mag_correction
        return A

    def design_subband_filter(self):
        """Band magnitudes interpolated over frequency, before the +1e-6 floor, the OLA correction and the direct-path
        magnitude (:224-239)."""
        H2 = self.design_filter(correct_OLA=False) - self.direct_path_mag_correction - 1e-6
        assert not torch.isnan(H2).any(), "decay is Nan"
        return H2

    def get_noise(self, noise=None):
        if noise is None:
            noise = torch.randn((self.length_rir,)).to(self.device)
        N = self.stft(noise) / torch.sqrt(torch.sum(self.window_padded ** 2))
        return N[:, 1:]

    def cons(self, X, length=None):
        Xr = torch.view_as_real(X.to(torch.complex64).contiguous())[None].contiguous()
        return torch.view_as_complex(self._eng.cons(Xr))[0].clone()

    def update_H(self, rir=None, H=None, use_noise=False, noise=None, phases=None):
        if rir is not None:
            super().update_H(rir=rir)
        elif H is not None:
            super().update_H(H=H)
        else:
            if use_noise:
                ph = self.get_noise(noise).angle()[:, :self.Nf].contiguous()
            elif phases is not None:
                self.params_phases[0] = phases
                ph = phases
            else:
                ph = self.params_phases[0]
            self._load_state(ph)
            self.H = torch.view_as_complex(self._eng.update_H().contiguous())[0].clone()
            if use_noise:
                self.params_phases[0] = torch.angle(self.H).detach()
        assert self.H.shape[0] == self.n_fft // 2 + 1 and self.H.shape[1] == self.Nf, "H.shape: {}".format(self.H.shape)

    def update_params(self, params_dict):
        t60 = torch.tensor(_hp(params_dict, "T60_breakpoints"), dtype=torch.float32, device=self.device)
        wts = torch.tensor(_hp(params_dict, "multiexp_weighting"), dtype=torch.float32, device=self.device)
        assert len(wts) == len(t60), "multiexp_weighting must have the same length as T60_breakpoints"
        if t60.shape[0] != 1:
            raise NotImplementedError("num_exponentials > 1 is not on the hot path")
        self.params[0] = 6.908 / (t60 * (self.sample_rate / self.hop_length))
        self.params[1] = wts

    def project_params(self):
        """Clamp decays to [6.908/(T60max*125), 6.908/(T60min*125)] and weights to [10^(Amin/20), 10^(Amax/20)]
        (:298-331); NaN parameters raise like the reference's asserts."""
        self.params[0] = torch.clamp(self.params[0].detach(), min=self.min_decay, max=self.max_decay)
        self.params[1] = torch.clamp(self.params[1].detach(), min=10 ** (self.Amin / 20), max=10 ** (self.Amax / 20))
        assert not torch.isnan(self.params[0]).any(), "decay is Nan"
        assert not torch.isnan(self.params[1]).any(), "weights is Nan"
