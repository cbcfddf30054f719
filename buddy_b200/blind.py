"""Blind reverb operator on the GPU, batched over utterances — the operator side of blind BUDDy.

Replaces, for the sampler, `BlindSubbandFiltering` (reference testing/operators/subband_filtering.py:116-351):
`update_H()` (design_filter -> cons -> minimum_phase_version), `degradation()` (apply_stft -> subband_filtering ->
apply_istft), `get_time_RIR()`, `project_params()`, and the body of `EulerHeunSamplerDPS.optimize_op`
(testing/EulerHeunSamplerDPS.py:71-113): rec-loss + RIR-noise regulariser, hand-derived gradients w.r.t. the
25 + 25 + 513x100 parameters of every utterance, Adam, projection.  Every utterance owns its H / parameters / Adam
state (the reference's single shared H is a B = 1 artefact).

Only kernel orchestration lives here; the arithmetic is in csrc/blind.cu and csrc/spectral.cu.
"""
import math

import torch

from . import ops
from .spectral import LossSTFT, _dft_mats, _use_fft, irfft_weights

LOSS_NORMS = {"l2_comp_stft_summean": "summean", "l2_comp_stft_sum": "sum", "l2_comp_stft_mean": "mean"}


def loss_norm(norm, bins, frames):
    """Factor on `weight` for the compressed-STFT loss variants of utils/losses.py:48-67, relative to the kernel's
    built-in mean over frames of the sum over bins (`summean`): plain sum / mean over all bins and frames."""
    return {"summean": 1.0, "sum": float(frames), "mean": 1.0 / bins}[norm]

EQ_FREQS = [0, 125, 250, 375, 500, 625, 750, 875, 1000, 1250, 1500, 1750, 2000, 2250, 2500, 2750, 3000, 3500, 4000,
            4500, 5000, 5500, 6000, 6500, 7000, 7500, 8000]


class BlindEngine:
    NFFT, WIN, HOP, F, NF = 1024, 512, 128, 513, 100
    LEN_RIR = 128 * 100            # 12 800
    T_MP = LEN_RIR + 128           # 12 928: length fed to minimum_phase_version
    N_BIG = 2 * T_MP               # 25 856 = 101 * 256
    N1 = 101
    RIR_LEN = LEN_RIR + 1024       # 13 824: excitation length of get_time_RIR

    def __init__(self, n, device, op_hp=None, sample_rate=16000):
        self.n, self.device = n, torch.device(device)
        dev = self.device
        g = (lambda k, d: (op_hp[k] if isinstance(op_hp, dict) else getattr(op_hp, k)) if op_hp is not None else d)
        t60min, t60max = g("T60min", 0.1), g("T60max", 2)
        self.max_decay = 6.908 / (t60min * (sample_rate / self.HOP))
        self.min_decay = 6.908 / (t60max * (sample_rate / self.HOP))
        self.wmin, self.wmax = 10 ** (g("Amin", 0) / 20), 10 ** (g("Amax", 40) / 20)
        self.loss_stft = LossSTFT(dev)
        w = torch.hann_window(self.WIN, dtype=torch.float64)
        norm = math.sqrt(float((w ** 2).sum()))
        if _use_fft():
            aw = irfft_weights(self.F, self.NFFT)
            self.cons_ana = ops.FftMat(torch.ones(self.F, dtype=torch.float64), w, dev)   # stft, not normalised
            self.cons_syn = ops.FftMat(aw, w, dev)                                         # istft (irfft * window)
            self.istft_syn = ops.FftMat(aw * norm, w, dev)                                 # apply_istft: X * sqrt(sum w^2)
        else:
            ana, syn = _dft_mats(self.NFFT, self.F, w, self.WIN, "cpu")
            self.cons_ana = ana.to(dev).contiguous()                               # stft, not normalised
            self.cons_syn = syn.to(dev).contiguous()                               # istft (irfft * window)
            self.istft_syn = (syn.double() * norm).float().to(dev).contiguous()    # apply_istft: X * sqrt(sum w^2) first
        # interpolation tables 27 knots -> 513 bins (torchcde stand-in: piecewise linear)
        freqs = torch.fft.rfftfreq(self.NFFT, d=1 / sample_rate)
        knots = torch.tensor(g("EQ_freqs", EQ_FREQS), dtype=torch.float32)
        k = torch.clamp(torch.bucketize(freqs, knots) - 1, 0, len(knots) - 2)
        frac = (freqs - knots[k]) / (knots[k + 1] - knots[k])
        wf = w.float()
        K = int(self.WIN / self.HOP - 1)
        corr = torch.stack([wf.sum() / wf[(K - i) * self.HOP:].sum() for i in range(K)])
        # direct-path magnitude correction |STFT(2 delta)|[:, 1:]  (compute_direct_path_mag_correction, :206-210)
        h = torch.zeros(self.LEN_RIR)
        h[0] = self.WIN / (self.HOP * 2)
        wp = torch.nn.functional.pad(wf, (0, self.NFFT - self.WIN))
        dp = torch.stft(h, self.NFFT, hop_length=self.HOP, win_length=self.NFFT, window=wp, center=True, onesided=True,
                        return_complex=True, normalized=False, pad_mode="constant")[:, 1:].abs()
        self.tabs = {"kidx": k.to(torch.int32).to(dev).contiguous(), "frac": frac.float().to(dev).contiguous(),
                     "corr": corr.float().to(dev).contiguous(), "dpmag": dp.float().to(dev).contiguous()}
        kk = torch.arange(256, dtype=torch.float64)
        self.tw512 = torch.stack([torch.cos(2 * math.pi * kk / 512), -torch.sin(2 * math.pi * kk / 512)],
                                 -1).float().to(dev)
        self.direct = torch.tensor([self.WIN / (self.HOP * 2)], device=dev)      # h[0] := 2.0
        self._env = {}
        # excitation spectrum of get_time_RIR (delta of length 13 824)
        d = torch.zeros(1, self.RIR_LEN, device=dev)
        d[0, 0] = 1
        self.Xdelta = self.loss_stft.forward(d)                                  # [1, 513, 113, 2]
        self.state = None

    # ---------------------------------------------------------------- tables
    def _inv_env(self, frames):
        """1 / OLA(window^2) of `frames` frames of the zero-padded Hann(512) window, padded-signal coordinates."""
        if frames not in self._env:
            total = (frames - 1) * self.HOP + self.WIN
            w2 = torch.hann_window(self.WIN, dtype=torch.float64) ** 2
            env = torch.zeros(total, dtype=torch.float64)
            for t in range(frames):
                env[t * self.HOP:t * self.HOP + self.WIN] += w2
            self._env[frames] = torch.where(env > 1e-11, 1 / env, torch.zeros_like(env)).float().to(self.device)
        return self._env[frames]

    # ---------------------------------------------------------------- state
    def init_state(self, B, decays, weights, phases, H):
        """decays/weights (1|B, 25), phases (513,100)|(B,513,100), H complex (513,100)|(B,513,100)."""
        dev = self.device

        def bexp(t, shape):
            t = torch.as_tensor(t).detach().to(dev, torch.float32).reshape(-1, *shape)
            assert t.shape[0] in (1, B), (t.shape, B)
            return t.expand(B, *shape).contiguous().clone()

        st = {"decays": bexp(decays, (25,)), "weights": bexp(weights, (25,)), "phases": bexp(phases, (self.F, self.NF))}
        Hc = torch.as_tensor(H).detach().to(dev).to(torch.complex64).reshape(-1, self.F, self.NF).contiguous()
        st["H"] = torch.view_as_real(Hc).expand(B, self.F, self.NF, 2).contiguous().clone()
        for k in ("decays", "weights", "phases"):
            st["m_" + k] = torch.zeros_like(st[k])
            st["v_" + k] = torch.zeros_like(st[k])
        st["B"] = B
        self.full = st
        self.steps = {}
        self.state = None
        self.buf = None
        self._bufs = {}
        return st

    def select(self, sl):
        """Work on the utterances full[sl] (views: in-place updates land in the full state)."""
        nb = len(range(*sl.indices(self.full["B"])))
        self.state = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in self.full.items()}
        self.state["B"] = nb
        self.state["key"] = sl.start or 0
        self.steps.setdefault(self.state["key"], 0)
        # scratch is per (micro-batch size, CUDA stream): micro-batches may run concurrently on different streams
        key = (nb, torch.cuda.current_stream().cuda_stream)
        if key not in self._bufs:
            self._alloc(nb)
            self._bufs[key] = self.buf
        self.buf = self._bufs[key]

    def _alloc(self, B):
        dev, N = self.device, self.N_BIG
        c = lambda: torch.empty(B, N, 2, device=dev)
        r = lambda: torch.empty(B, N, device=dev)
        self.buf = {"u": torch.zeros(B, N, device=dev), "Hf": c(), "m": r(), "phi": r(), "c1": c(), "c2": c(),
                    "work": c(), "r1": r(), "A": torch.empty(B, self.F, self.NF, device=dev),
                    "H0": torch.empty(B, self.F, self.NF + 2, 2, device=dev),
                    "sig": torch.empty(B, 384 + self.T_MP, device=dev),
                    "Xd": self.Xdelta.expand(B, -1, -1, -1).contiguous()}

    # ---------------------------------------------------------------- H = cons(A e^{j phi})
    def update_H(self):
        """design_filter + cons (istft -> minimum phase -> h[0]=2 -> stft); keeps what the backward needs."""
        st, bf = self.state, self.buf
        B, N, T = st["B"], self.N_BIG, self.T_MP
        ops.blind_design_fwd(st["decays"], st["weights"], st["phases"], self.tabs, bf["A"], bf["H0"])
        H = self._cons_tail()
        st["H"].copy_(H)
        return H

    def cons(self, Hin):
        """`BlindSubbandFiltering.cons` (subband_filtering.py:333-351) of an arbitrary filter [B,513,100,2]:
        pad one frame each side -> istft(12 800) -> pad 128 -> minimum phase -> h[0] = 2 -> stft, drop first/last frame."""
        bf = self.buf
        bf["H0"].zero_()
        bf["H0"][:, :, 1:-1].copy_(Hin)
        return self._cons_tail()

    def _cons_tail(self):
        st, bf = self.state, self.buf
        B, N, T = st["B"], self.N_BIG, self.T_MP
        fr = torch.empty(B, self.NF + 2, self.WIN, device=self.device)
        ops.stft_synthesis(bf["H0"], self.cons_syn, self.NF + 2, fr)
        ops.ola_gather(fr, self.HOP, self.NFFT // 2, self.LEN_RIR, bf["u"], tab=self._inv_env(self.NF + 2))
        ops.fft_mixed(bf["u"], True, bf["work"], bf["Hf"], self.N1, -1, self.tw512)
        ops.minphase_pw(0, B, N, T, c0=bf["Hf"], or0=bf["m"], oc=bf["c1"])
        ops.fft_mixed(bf["c1"], False, bf["work"], bf["c2"], self.N1, -1, self.tw512)
        ops.minphase_pw(1, B, N, T, c0=bf["c2"], oc=bf["c1"])
        ops.fft_mixed(bf["c1"], False, bf["work"], bf["c2"], self.N1, +1, self.tw512)
        ops.minphase_pw(2, B, N, T, c0=bf["c2"], r0=bf["m"], or0=bf["phi"], oc=bf["c1"])
        ops.fft_mixed(bf["c1"], False, bf["work"], bf["c2"], self.N1, +1, self.tw512)
        h2 = torch.empty(B, T, device=self.device)
        ops.minphase_pw(3, B, N, T, c0=bf["c2"], r0=self.direct, or0=h2)
        ops.pad_signal(h2, 384, 384 + T, 0, bf["sig"])
        H = torch.empty(B, self.F, self.NF, 2, device=self.device)
        ops.stft_analysis(bf["sig"], self.cons_ana, self.HOP, self.NF, self.NF, H)
        return H

    def update_H_backward(self, dH):
        """dL/dH [B,513,100,2] -> gradients w.r.t. (decays, weights, phases)."""
        st, bf = self.state, self.buf
        B, N, T = st["B"], self.N_BIG, self.T_MP
        dev = self.device
        fr = torch.empty(B, self.NF, self.WIN, device=dev)
        ops.stft_synthesis(dH, self.cons_ana, self.NF, fr)
        dh2 = torch.empty(B, T, device=dev)
        ops.ola_gather(fr, self.HOP, 384, T, dh2)
        gz = bf["r1"]
        ops.minphase_pw(4, B, N, T, r0=dh2, or0=gz)
        ops.fft_mixed(gz, True, bf["work"], bf["c1"], self.N1, -1, self.tw512)
        gm1 = torch.empty(B, N, device=dev)
        ops.minphase_pw(5, B, N, T, c0=bf["c1"], r0=bf["m"], r1=bf["phi"], or0=gm1, oc=bf["c2"])
        ops.fft_mixed(bf["c2"], False, bf["work"], bf["c1"], self.N1, -1, self.tw512)
        ops.minphase_pw(1, B, N, T, c0=bf["c1"], oc=bf["c2"], scale_inv_n=True)
        ops.fft_mixed(bf["c2"], False, bf["work"], bf["c1"], self.N1, +1, self.tw512)
        ops.minphase_pw(6, B, N, T, c0=bf["c1"], c1=bf["Hf"], r0=bf["m"], r1=gm1, oc=bf["c2"])
        ops.fft_mixed(bf["c2"], False, bf["work"], bf["c1"], self.N1, +1, self.tw512)
        dh = torch.empty(B, T, device=dev)
        ops.minphase_pw(7, B, N, T, c0=bf["c1"], or0=dh)
        # adjoint of istft(length 12 800): d frames[t][n] = dh[128 t + n - 512] / env
        frames = self.NF + 2
        total = (frames - 1) * self.HOP + self.WIN
        dpad = torch.empty(B, total, device=dev)
        ops.pad_signal(dh[:, :self.LEN_RIR], self.NFFT // 2, total, 0, dpad, tab=self._inv_env(frames))
        dH0 = torch.empty(B, self.F, frames, 2, device=dev)
        ops.stft_analysis(dpad, self.cons_syn, self.HOP, frames, frames, dH0)
        dph = torch.empty_like(st["phases"])
        dd, dw = torch.empty_like(st["decays"]), torch.empty_like(st["weights"])
        ops.blind_design_bwd(st["decays"], st["weights"], st["phases"], bf["A"], self.tabs, dH0, dph, dd, dw)
        return dd, dw, dph

    # ---------------------------------------------------------------- A_H and adjoints
    def apply_istft(self, Ys, n):
        B, _, frames, _ = Ys.shape
        fr = torch.empty(B, frames, self.WIN, device=self.device)
        ops.stft_synthesis(Ys, self.istft_syn, frames, fr)
        out = torch.empty(B, n, device=self.device)
        return ops.ola_gather(fr, self.HOP, self.NFFT // 2 + self.WIN // 2, n, out, tab=self._inv_env(frames))

    def apply_istft_adjoint(self, g):
        B, n = g.shape
        frames = self.loss_stft.frames(n)
        total = (frames - 1) * self.HOP + self.WIN
        gp = torch.empty(B, total, device=self.device)
        ops.pad_signal(g, self.NFFT // 2 + self.WIN // 2, total, 0, gp, tab=self._inv_env(frames))
        out = torch.empty(B, self.F, frames, 2, device=self.device)
        return ops.stft_analysis(gp, self.istft_syn, self.HOP, frames, frames, out)

    def degradation_from_stft(self, X, H, n):
        Ys = ops.subband_fir(X, H, torch.empty_like(X), Nf=self.NF, pre=1, mode=0)
        return self.apply_istft(Ys, n)

    def degradation(self, x, H=None):
        H = self.state["H"] if H is None else H
        return self.degradation_from_stft(self.loss_stft.forward(x), H, x.shape[1])

    def get_time_RIR(self, H=None):
        H = self.state["H"] if H is None else H
        return self.degradation_from_stft(self.buf["Xd"][:H.shape[0]], H, self.RIR_LEN)

    def likelihood_grad(self, x_den, Y, weight, comp, norm="summean"):
        """rec = loss(y, A_H(x_den)) per utterance and d rec / d x_den, with the current (detached) H."""
        n = x_den.shape[1]
        B = x_den.shape[0]
        H = self.state["H"]
        X = self.loss_stft.forward(x_den)
        y_hat = self.degradation_from_stft(X, H, n)
        Yh = self.loss_stft.forward(y_hat)
        loss = torch.empty(B, device=self.device, dtype=torch.float64)
        G = torch.empty_like(Yh)
        ops.comp_loss(Y, Yh, Yh.shape[2], comp, weight * loss_norm(norm, Yh.shape[1], Yh.shape[2]), loss, G)
        gYs = self.apply_istft_adjoint(self.loss_stft.adjoint(G, n))
        gX = ops.subband_fir(gYs, H, torch.empty_like(gYs), Nf=self.NF, pre=1, mode=1)
        return self.loss_stft.adjoint(gX, n), loss

    # ---------------------------------------------------------------- optimize_op
    def optimize(self, x_den, Y, t_hat, noise_fn, hp):
        """`op_updates_per_step` Adam iterations on every utterance's operator (EulerHeunSamplerDPS.py:71-113)."""
        st = self.state
        B, n = x_den.shape
        dev = self.device
        X = self.loss_stft.forward(x_den)
        t_op = max(min(float(t_hat), hp["crop_max"]), hp["crop_min"])
        for _ in range(hp["iters"]):
            H = self.update_H()
            dH = torch.empty_like(H) if hp.get("use_rec", True) else torch.zeros_like(H)
            loss = lreg = None
            if hp.get("use_rec", True):
                # reconstruction loss
                y_hat = self.degradation_from_stft(X, H, n)
                Yh = self.loss_stft.forward(y_hat)
                loss = torch.empty(B, device=dev, dtype=torch.float64)
                G = torch.empty_like(Yh)
                ops.comp_loss(Y, Yh, Yh.shape[2], hp["comp"],
                              hp["w_rec"] * loss_norm(hp.get("norm_rec", "summean"), Yh.shape[1], Yh.shape[2]), loss, G)
                gYs = self.apply_istft_adjoint(self.loss_stft.adjoint(G, n))
                ops.subband_fir(X, gYs, dH, Nf=self.NF, pre=1, mode=2)
            if hp.get("use_reg", True):
                # RIR-noise regulariser: loss(rir, (rir + t_op * noise).detach())
                Xd = self.buf["Xd"]
                rir = self.degradation_from_stft(Xd, H, self.RIR_LEN)
                noisy = ops.lincomb3(torch.empty_like(rir), rir, torch.ones(B, device=dev),
                                     noise_fn((B, self.RIR_LEN)), torch.full((B,), t_op, device=dev))
                R, Rt = self.loss_stft.forward(rir), self.loss_stft.forward(noisy)
                lreg = torch.empty(B, device=dev, dtype=torch.float64)
                Gr = torch.empty_like(R)
                ops.comp_loss(Rt, R, R.shape[2], hp.get("comp_reg", hp["comp"]),
                              hp["w_reg"] * loss_norm(hp.get("norm_reg", "summean"), R.shape[1], R.shape[2]), lreg, Gr)
                gYd = self.apply_istft_adjoint(self.loss_stft.adjoint(Gr, self.RIR_LEN))
                ops.subband_fir(Xd, gYd, dH, Nf=self.NF, pre=1, mode=2, accumulate=True)
            dd, dw, dph = self.update_H_backward(dH)
            self.steps[st["key"]] += 1
            step = self.steps[st["key"]]
            inf = float("inf")
            ops.adam_project(st["decays"], dd, st["m_decays"], st["v_decays"], step, hp["lr"], hp["beta1"],
                             hp["beta2"], 1e-8, self.min_decay, self.max_decay, -inf, inf)
            ops.adam_project(st["weights"], dw, st["m_weights"], st["v_weights"], step, hp["lr"], hp["beta1"],
                             hp["beta2"], 1e-8, self.wmin, self.wmax, -inf, inf)
            ops.adam_project(st["phases"].view(B, -1), dph.view(B, -1), st["m_phases"].view(B, -1),
                             st["v_phases"].view(B, -1), step, hp["lr"], hp["beta1"], hp["beta2"], 1e-8, -inf,
                             inf, -inf, inf)
            self.last_losses = (loss, lreg)
