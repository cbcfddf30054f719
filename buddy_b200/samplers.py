"""Drop-in samplers: `Sampler`, `EulerHeunSampler`, `EulerHeunSamplerDPS`
(reference testing/Sampler.py:6-72, testing/EulerHeunSampler.py:6-107, testing/EulerHeunSamplerDPS.py:14-204).

Same constructor `(model, diff_params, args)`, same methods and return values, so Hydra can swap them in by
`tester.sampler._target_=buddy_b200.samplers.EulerHeunSamplerDPS` with `test.py` unchanged.  Differences by design
(SURVEY.md App. C2/C7):
  * a batch is B independent utterances: every `.std()`, `torch.norm`, loss mean is per utterance
    (the reference only ever runs B = 1, where the two coincide);
  * the whole loop stays on the device: the schedule, gamma, t_hat and every coefficient are host floats computed
    up-front, so there is no device->host sync inside the loop; noise comes from a per-utterance Philox stream
    (or from `self.noise_source`, an iterator of pre-drawn tensors, for parity tests);
  * the likelihood gradient is not taken by autograd but by the hand-written VJP kernels.
"""
import math

import torch

from . import ops
from .ncsnpp import NCSNppTime
from .spectral import LossSTFT, RirConv


def _get(cfg, name, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default) if default is not None or hasattr(cfg, name) else default


class Sampler:
    def __init__(self, model, diff_params, args):
        self.model = model.eval()
        self.diff_params = diff_params
        self.args = args
        sp = args.tester.sampling_params
        self.sde_hp = diff_params.sde_hp if sp.same_as_training else sp.sde_hp
        self.T = sp.T
        self.step_counter = 0
        # extensions (all optional)
        self.noise_source = None          # iterator of pre-drawn N(0,1) tensors (parity tests)
        # Philox stream of utterance b = seed + utterance_offset + b.  seed_base = None (default): every predict*()
        # call draws a fresh 62-bit seed from torch's global CPU generator, so `torch.manual_seed` controls the run and
        # successive calls / successive utterances of the reference's one-at-a-time Tester loop get different noise —
        # the behaviour of the reference's `torch.randn` (EulerHeunSampler.py:21,43).  seed_base = int: fully explicit
        # streams (results independent of batch split and GPU count; repeated calls return identical samples).
        self.seed_base = None
        self._run_seed = 0
        self.utterance_offset = 0         # global index of this rank's first utterance (multi-GPU shards)
        self.nan_guard = True             # one device->host finiteness check per predict*() call (DPS.py:90,103)
        self.utterance_ids = None         # or: explicit global index of every utterance of the batch
        self.micro_batch = 32             # utterances per network evaluation (~1.9 GB of activations each)
        self.n_streams = 1                # micro-batches in flight on separate CUDA streams (results identical).  >1
        #                                   overlaps one micro-batch's HBM-bound kernels with another's convolutions;
        #                                   measured gain on B200 is <1 % because the convolutions already run at the
        #                                   board power cap, so the default stays 1 (half the activation memory)
        self._streams = None
        self.use_graphs = True            # replay captured CUDA graphs of the network forward / VJP for small batches
        self._draw = 0

    ACT_BYTES_PER_SAMPLE = 1.9e9 / 65536    # saved activations of one network evaluation, per audio sample

    def _for_micro_batches(self, B, body):
        """Run body(slice) for every micro-batch, round-robin over `n_streams` CUDA streams forked from / joined
        to the current stream.  Micro-batches touch disjoint utterances, so the order of execution is free."""
        slices = [slice(s, min(B, s + self.micro_batch)) for s in range(0, B, self.micro_batch)]
        ns = min(self.n_streams, len(slices))
        if ns > 1:
            # every micro-batch in flight keeps its saved activations (1.9 GB per 4.096 s utterance); past the device's
            # memory the caching allocator falls back to synchronous cudaFree / cudaMalloc retries and the run crawls
            # (seen on B200 with 2 x 32 utterances in the blind configuration) — refuse instead
            n = getattr(self, "y", None).shape[1] if getattr(self, "y", None) is not None else 65536
            need = ns * min(self.micro_batch, B) * self.ACT_BYTES_PER_SAMPLE * n
            total = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
            if need > 0.6 * total:
                raise ValueError(f"n_streams={ns} x micro_batch={self.micro_batch} needs ~{need / 2**30:.0f} GiB of saved "
                                 f"activations ({total / 2**30:.0f} GiB on this device): lower micro_batch or n_streams")
        if ns <= 1:
            for sl in slices:
                body(sl)
            return
        main = torch.cuda.current_stream()
        if self._streams is None or len(self._streams) < ns or self._streams[0].device != main.device:
            self._streams = [torch.cuda.Stream(device=main.device) for _ in range(ns)]
        fork = torch.cuda.Event()
        fork.record(main)
        for k, sl in enumerate(slices):
            st = self._streams[k % ns]
            if k < ns:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                body(sl)
        for st in self._streams[:ns]:
            main.wait_stream(st)

    # ---- interface (Sampler.py:23-37: "abstract" placeholders that return None in the base class) -----
    def predict(self, *args, **kwargs):
        return None

    def predict_unconditional(self, *args, **kwargs):
        return None

    def predict_conditional(self, *args, **kwargs):
        return None

    def step(self, *args, **kwargs):
        return None

    # ---- schedule (Sampler.py:39-56) -----------------------------------------------------------------
    def create_schedule(self, sigma_min=None, sigma_max=None, rho=None, T=None):
        sigma_min = self.sde_hp.sigma_min if sigma_min is None else sigma_min
        sigma_max = self.sde_hp.sigma_max if sigma_max is None else sigma_max
        rho = self.sde_hp.rho if rho is None else rho
        T = self.T if T is None else T
        if self.args.tester.sampling_params.schedule == "edm":
            a = torch.arange(0, T + 1)
            t = (sigma_max ** (1 / rho) + a / (T - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
            t[-1] = 0
            return t
        raise NotImplementedError(f"schedule {self.args.tester.sampling_params.schedule} not implemented")

    def Tweedie2score(self, tweedie, xt, t):
        return self.diff_params.Tweedie2score(tweedie, xt, t)

    def get_Tweedie_estimate(self, x, t_i):
        return self.diff_params.denoiser(x.unsqueeze(1), self.model, t_i).squeeze(1)

    # ---- noise ---------------------------------------------------------------------------------------
    def _randn(self, shape, device, draw=None, first=0):
        """N(0,1) [B, n].  Philox stream of utterance b = run seed + utterance_offset + first + b; `draw` is the draw
        index inside the stream (explicit, so results do not depend on micro-batching or on the GPU count)."""
        if self.noise_source is not None:
            z = next(self.noise_source).to(device=device, dtype=torch.float32)
            if z.dim() == 1:
                z = z[None]
            assert tuple(z.shape) == tuple(shape), (z.shape, shape)
            return z.contiguous()
        B, n = shape
        if self.utterance_ids is not None:      # explicit global utterance indices (batches that are not contiguous)
            ids = torch.as_tensor(self.utterance_ids, dtype=torch.int64)[first:first + B]
            assert ids.numel() == B, "utterance_ids must list one index per utterance of the batch"
            seeds = ids.to(device) + self._run_seed
        else:
            seeds = torch.arange(B, dtype=torch.int64, device=device) + (self._run_seed + self.utterance_offset + first)
        out = torch.empty(B, n, device=device)
        if draw is None:
            draw = self._draw
            self._draw += 1
        ops.philox_normal(seeds, draw, out)
        return out

    def _start_run(self):
        """Per-call noise state: draw counter back to 0 and the run seed (see `seed_base`)."""
        self._draw = 0
        if self.seed_base is None:
            self._run_seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        else:
            self._run_seed = int(self.seed_base)

    def _check_finite(self, x, what):
        """The reference asserts on NaN inside the loop (EulerHeunSamplerDPS.py:90,103; subband_filtering.py:238,
        330-331) — a host sync per check.  Here NaN/Inf propagate through every kernel (no clamp swallows them) and
        are checked ONCE per call on the per-utterance sums of the result."""
        if not self.nan_guard:
            return
        st = ops.row_stats(x.reshape(x.shape[0], -1).contiguous())
        bad = (~torch.isfinite(st).all(dim=1)).nonzero().flatten().tolist()
        if bad:
            raise FloatingPointError(f"{what} is NaN/Inf for utterance(s) {bad} of the batch")


class NoSampler(Sampler):
    """Sampler.py:74-86: the do-nothing sampler some tester configurations name (every method returns None)."""


def _vec(v, B, device):
    return torch.full((B,), float(v), device=device, dtype=torch.float32)


class EulerHeunSampler(Sampler):
    def __init__(self, model, diff_params, args):
        super().__init__(model, diff_params, args)
        sp = self.args.tester.sampling_params
        self.Schurn, self.Snoise, self.Stmin, self.Stmax, self.order = sp.Schurn, sp.Snoise, sp.Stmin, sp.Stmax, sp.order

    def initialize_x(self, shape, device, schedule):
        return float(schedule[0]) * self._randn(shape, device)

    def get_gamma(self, t):
        N = t.shape[0]
        gamma = torch.zeros(t.shape)
        idx = torch.logical_and(t > self.Stmin, t < self.Stmax)
        gamma[idx] = gamma[idx] + torch.min(torch.Tensor([self.Schurn / N, 2 ** (1 / 2) - 1]))
        return gamma

    def stochastic_timestep(self, x, t, gamma, Snoise=1):
        t, gamma = float(t), float(gamma)
        t_hat = t + gamma * t
        eps = self._randn(x.shape, x.device)
        B = x.shape[0]
        x_hat = ops.lincomb3(torch.empty_like(x), x, _vec(1.0, B, x.device), eps,
                             _vec(math.sqrt(max(t_hat ** 2 - t ** 2, 0.0)) * Snoise, B, x.device))
        return x_hat, t_hat

    # ---- denoiser on raw kernels (no autograd graph) --------------------------------------------------
    def _edm_scalars(self, sigma):
        sd = float(self.diff_params.sigma_data)
        s2 = sigma * sigma + sd * sd
        return sd * sd / s2, sigma * sd / math.sqrt(s2), 1.0 / math.sqrt(s2), 0.25 * math.log(sigma)

    def _denoise(self, x_hat, sigma, save):
        """x_den = cskip*x + cout*F(cin*x, cnoise) for a micro-batch; returns (x_den, ctx)."""
        net = self.model
        if not isinstance(net, NCSNppTime):
            raise TypeError("buddy_b200 samplers drive buddy_b200.NCSNppTime (load the reference checkpoint into it)")
        B, n = x_hat.shape
        dev = x_hat.device
        cskip, cout, cin, cnoise = self._edm_scalars(sigma)
        eng, st = net.engine(), net.stft_engine()
        spec = st.forward(x_hat, scale_b=_vec(cin, B, dev))
        fspec, ctx = eng.forward(spec, _vec(cnoise, B, dev), save=save, graph=self.use_graphs and self.n_streams == 1)
        F = st.inverse(fspec, n)
        x_den = ops.lincomb3(torch.empty_like(x_hat), x_hat, _vec(cskip, B, dev), F, _vec(cout, B, dev))
        return x_den, ctx

    def step(self, x_i, t_i, t_iplus1, gamma_i, blind=False):
        x_hat, t_hat = self.stochastic_timestep(x_i, t_i, gamma_i)
        t_next = float(t_iplus1)
        B, dev = x_hat.shape[0], x_hat.device
        d, x_den = self._ode_term(x_hat, t_hat)
        dt = t_next - t_hat
        if t_next != 0 and self.order == 2:
            x_prime = ops.lincomb3(torch.empty_like(x_hat), x_hat, _vec(1.0, B, dev), d, _vec(dt, B, dev))
            d2, x_den = self._ode_term(x_prime, t_next, second=True)
            x_next = ops.lincomb3(torch.empty_like(x_hat), x_hat, _vec(1.0, B, dev), d, _vec(0.5 * dt, B, dev), d2,
                                  _vec(0.5 * dt, B, dev))
        else:
            x_next = ops.lincomb3(torch.empty_like(x_hat), x_hat, _vec(1.0, B, dev), d, _vec(dt, B, dev))
        return x_next, x_den

    def _ode_term(self, x, sigma, second=False):
        """d = (x - D(x, sigma)) / sigma   (Tweedie2score + _ode_integrand, edm.py:83-96), micro-batched."""
        B, dev = x.shape[0], x.device
        d = torch.empty_like(x)
        x_den = torch.empty_like(x)

        def body(sl):
            xd, _ = self._denoise(x[sl].contiguous(), sigma, save=False)
            x_den[sl] = xd
            nb = xd.shape[0]
            d[sl] = ops.lincomb3(torch.empty_like(xd), x[sl].contiguous(), _vec(1.0 / sigma, nb, dev), xd,
                                 _vec(-1.0 / sigma, nb, dev))

        self._for_micro_batches(B, body)
        return d, x_den

    def predict(self, shape, device, blind=False):
        t = self.create_schedule()
        self._start_run()
        x = self.initialize_x(tuple(shape), device, t)
        gamma = self.get_gamma(t)
        x_den = None
        for i in range(self.T):
            self.step_counter = i
            x, x_den = self.step(x, t[i], t[i + 1], gamma[i], blind)
        self._last_x_den = x_den
        self._check_finite(x, "sample")
        return x.detach()

    def predict_unconditional(self, shape, device):
        self.y = None
        self.degradation = None
        return self.predict(shape, device)

    def predict_conditional(self, *args, **kwargs):
        raise NotImplementedError


class EulerHeunSamplerDPS(EulerHeunSampler):
    """Diffusion posterior sampling with the compressed-STFT likelihood (EulerHeunSamplerDPS.py:14-204)."""

    def __init__(self, model, diff_params, args):
        super().__init__(model, diff_params, args)
        self.zeta = self.args.tester.posterior_sampling.zeta

    def initialize_x(self, shape, device, schedule):
        ps = self.args.tester.posterior_sampling
        mode = ps.warm_initialization.mode
        z = self._randn(shape, device)
        B = shape[0]
        if mode == "none":
            return float(schedule[0]) * z
        if mode == "reverb_scaled":
            st = ops.row_stats(self.y)
            n = self.y.shape[1]
            std = torch.sqrt((st[:, 1] - st[:, 0] ** 2 / n) / (n - 1)).float()     # unbiased, per utterance
            coef = (float(ps.warm_initialization.scaling_factor) / std).contiguous()
            return ops.lincomb3(torch.empty_like(z), self.y, coef, z, _vec(float(schedule[0]), B, device))
        if mode == "wpe_scaled":
            # EulerHeunSamplerDPS.py:32-54 — the reference goes to the CPU (nara_wpe, numpy); here STFT -> WPE -> iSTFT
            # stay on the device, one one-channel problem per utterance (buddy_b200/wpe.py, csrc/wpe.cu)
            from .wpe import WpeDereverb
            w = ps.warm_initialization.wpe
            x_pred = WpeDereverb(device, taps=w.taps, delay=w.delay, iterations=w.iterations)(self.y)
            st = ops.row_stats(x_pred)
            n = x_pred.shape[1]
            std = torch.sqrt((st[:, 1] - st[:, 0] ** 2 / n) / (n - 1)).float()     # unbiased, per utterance
            coef = (float(ps.warm_initialization.scaling_factor) / std).contiguous()
            return ops.lincomb3(torch.empty_like(z), x_pred, coef, z, _vec(float(schedule[0]), B, device))
        raise NotImplementedError(mode)

    # ---- operator binding -----------------------------------------------------------------------------
    _OP_HP = dict(NFFT=1024, win_length=512, hop=128, window="hann")
    _OP_HP_BLIND = dict(Nf=100, minimum_phase=True, fix_direct_path=True, fix_EQ_extremes=True, clamp_decay=True,
                        strictly_decreasing_decay=False)

    def _validate_operator(self, operator, B, blind):
        """The kernels implement the shipped operator configuration (conf/tester/*.yaml op_hp); anything else must fail
        loudly instead of silently computing a different loss / filter."""
        hp = getattr(operator, "op_hp", None)
        want = dict(self._OP_HP, **(self._OP_HP_BLIND if blind else {}))
        for k, v in want.items():
            got = _get(hp, k)
            if got is not None and got != v:
                raise NotImplementedError(f"operator op_hp.{k} = {got!r}: the CUDA path implements {v!r} only")
        if not blind:
            return
        eq = _get(hp, "EQ_freqs")
        if eq is not None and len(eq) != 27:
            raise NotImplementedError(f"op_hp.EQ_freqs has {len(eq)} knots: the CUDA path implements 27 (25 bands)")
        if int(getattr(operator, "num_exponentials", 1)) != 1:
            raise NotImplementedError("num_exponentials > 1 is not on the hot path (shipped config: one exponential)")
        p0 = operator.params[0]
        if p0.shape[-1] != 25 or (p0.dim() == 2 and p0.shape[0] not in (1, B)):
            raise NotImplementedError(f"operator.params[0] has shape {tuple(p0.shape)}: expected (1, 25) or (B, 25)")
        if float(_get(self.args.tester.posterior_sampling.blind_hp, "weight_decay", 0) or 0) != 0:
            raise NotImplementedError("blind_hp.weight_decay != 0 is not implemented (shipped config: 0)")
        if str(_get(self.args.tester.posterior_sampling.blind_hp, "optimizer", "adam")).lower() != "adam":
            raise NotImplementedError("blind_hp.optimizer: only adam is implemented")

    def _bind_operator(self, operator, y, blind):
        dev = y.device
        n = y.shape[1]
        self._validate_operator(operator, y.shape[0], blind)
        self._loss_stft = LossSTFT(dev)
        ps = self.args.tester.posterior_sampling
        self._loss_w = float(ps.rec_loss.weight)
        self._loss_c = float(ps.rec_loss.compression_factor)
        from .blind import LOSS_NORMS
        if ps.rec_loss.name not in LOSS_NORMS:
            raise NotImplementedError(f"rec_loss {ps.rec_loss.name}: the compressed-STFT losses {sorted(LOSS_NORMS)} "
                                      "are on the hot path (utils/losses.py:48-67)")
        self._loss_norm = LOSS_NORMS[ps.rec_loss.name]
        self._Y = self._loss_stft.forward(y)
        self._is_blind = bool(blind)          # operator parameters are optimised along the trajectory
        self._subband = bool(blind)           # the likelihood goes through the sub-band (STFT-domain) operator
        self._eval_index = 0
        if blind:
            from .blind import BlindEngine
            hp = ps.blind_hp
            rp, reg = ps.rec_loss_params, ps.RIR_noise_regularization
            # utils/losses.py:19-20: name "none" switches a term off (EulerHeunSamplerDPS.py:87-104)
            names = (rp.name, reg.loss.name)
            if any(nm != "none" and nm not in LOSS_NORMS for nm in names) or names == ("none", "none"):
                raise NotImplementedError(f"blind path: losses {names}: the compressed-STFT losses {sorted(LOSS_NORMS)} "
                                          "(or 'none' for one of the two terms) are on the hot path")
            self._blind = BlindEngine(n, dev, op_hp=getattr(operator, "op_hp", None),
                                      sample_rate=self.args.exp.sample_rate)
            self._blind.init_state(y.shape[0], operator.params[0], operator.params[1], operator.params_phases[0],
                                   operator.H)
            self._blind_hp = dict(iters=int(hp.op_updates_per_step), lr=float(hp.lr_op), beta1=float(hp.beta1),
                                  beta2=float(hp.beta2), comp=float(_get(rp, "compression_factor", 1.0)),
                                  w_rec=float(_get(rp, "weight", 1.0)), w_reg=float(_get(reg.loss, "weight", 1.0)),
                                  crop_max=float(reg.crop_sigma_max),
                                  crop_min=float(reg.crop_sigma_min), use_rec=rp.name != "none",
                                  use_reg=reg.loss.name != "none", norm_rec=LOSS_NORMS.get(rp.name, "summean"),
                                  norm_reg=LOSS_NORMS.get(reg.loss.name, "summean"),
                                  comp_reg=float(_get(reg.loss, "compression_factor", _get(rp, "compression_factor", 1.0))))
            return
        rir = getattr(operator, "params", None)
        H = getattr(operator, "H", None)
        if torch.is_tensor(H) and H.is_complex() and not torch.is_tensor(rir):
            # informed SubbandFiltering operator (subband_filtering.py:8-113, `update_H(rir=...)` / `update_H(H=...)`): a
            # KNOWN filter H (513, 100) or one per utterance (B, 513, 100); same likelihood chain as the blind path,
            # nothing is optimised
            from .blind import BlindEngine
            B = y.shape[0]
            self._blind = BlindEngine(n, dev, op_hp=getattr(operator, "op_hp", None),
                                      sample_rate=self.args.exp.sample_rate)
            z = torch.zeros(1, 25, device=dev)
            self._blind.init_state(B, z, z, torch.zeros(self._blind.F, self._blind.NF, device=dev), H)
            self._subband = True
            return
        if rir is None:
            raise ValueError("operator has no RIR (`.params`): call operator.update_params(rir) first")
        self._rir = RirConv(rir.detach().to(dev), n, dev)

    def get_likelihood_score(self, x_den, ctx, x_hat, sigma, Y=None, first=0):
        """zeta/(||g||/sqrt(audio_len)+1e-8) * g with g = d rec_loss / d x_hat through the denoiser; per utterance.
        `Y` = observation spectra of these utterances (default: the whole bound batch), `first` = their offset."""
        if Y is None:
            Y = self._Y[first:first + x_den.shape[0]]
        B, n = x_den.shape
        dev = x_den.device
        cskip, cout, cin, _ = self._edm_scalars(sigma)
        net = self.model
        eng, st = net.engine(), net.stft_engine()
        if self._subband:
            if not self._is_blind:
                self._blind.select(slice(first, first + B))
            gd, loss = self._blind.likelihood_grad(x_den, Y, self._loss_w, self._loss_c, self._loss_norm)
        else:
            y_hat = self._rir.forward(x_den, first)
            Yh = self._loss_stft.forward(y_hat)
            loss = torch.empty(B, device=dev, dtype=torch.float64)
            G = torch.empty_like(Yh)
            from .blind import loss_norm
            ops.comp_loss(Y, Yh, Yh.shape[2], self._loss_c,
                          self._loss_w * loss_norm(self._loss_norm, Yh.shape[1], Yh.shape[2]), loss, G)
            gd = self._rir.adjoint(self._loss_stft.adjoint(G, n), first)    # d loss / d x_den
        dspec = eng.vjp(ctx, st.inverse_adjoint(gd))
        v = st.forward_adjoint(dspec, n, scale_b=_vec(cin * cout, B, dev))
        g = ops.lincomb3(torch.empty_like(gd), gd, _vec(cskip, B, dev), v, _vec(1.0, B, dev))
        gs = ops.row_stats(g)
        normguide = torch.sqrt(gs[:, 1]).float() / (self.args.exp.audio_len ** 0.5)
        coef = (self.zeta / (normguide + 1e-8)).contiguous()
        return g, coef, loss

    def _ode_term(self, x, sigma, second=False):
        """d = (x - x_den)/sigma + lh_score   (EulerHeunSamplerDPS.py:118-134), micro-batched.
        second = the Heun correction evaluation (:136-150): the reference rescales x_den (magnitude constraint,
        :128-129) only after the FIRST evaluation of a step, the second one enters Tweedie2score as is."""
        B, dev = x.shape[0], x.device
        d = torch.empty_like(x)
        x_den = torch.empty_like(x)
        ps = self.args.tester.posterior_sampling
        rescale = bool(ps.constraint_speech_magnitude.use) and not second
        def body(sl):
            xs = x[sl].contiguous()
            nb = xs.shape[0]
            xd, ctx = self._denoise(xs, sigma, save=True)
            if self._is_blind:
                self.optimize_op(xd, sigma, sl)
            g, coef, loss = self.get_likelihood_score(xd, ctx, xs, sigma, self._Y[sl], sl.start or 0)
            del ctx
            if rescale:
                st = ops.row_stats(xd)
                n = xd.shape[1]
                std = torch.sqrt((st[:, 1] - st[:, 0] ** 2 / n) / (n - 1)).float()
                sc = (float(ps.constraint_speech_magnitude.speech_scaling) / std).contiguous()
                xd = ops.lincomb3(torch.empty_like(xd), xd, sc)
            x_den[sl] = xd
            d[sl] = ops.lincomb3(torch.empty_like(xd), xs, _vec(1.0 / sigma, nb, dev), xd, _vec(-1.0 / sigma, nb, dev),
                                 g, coef)
            self.rec_loss_value = loss

        self._for_micro_batches(B, body)
        self._eval_index += 1
        return d, x_den

    def optimize_op(self, x_den, t, sl=None):
        """Operator-parameter updates for the utterances in `sl` (EulerHeunSamplerDPS.optimize_op, :71-113)."""
        sl = slice(0, x_den.shape[0]) if sl is None else sl
        first = sl.start or 0
        k = [0]
        base = 1_000_000 + self._eval_index * 256

        def noise_fn(shape):
            z = self._randn(shape, x_den.device, draw=base + k[0], first=first)
            k[0] += 1
            return z

        self._blind.select(sl)
        self._blind.optimize(x_den, self._Y[sl], t, noise_fn, self._blind_hp)

    def predict(self, shape, device, blind=False):
        t = self.create_schedule()
        self._start_run()
        self._eval_index = 0
        x = self.initialize_x(tuple(shape), device, t)
        gamma = self.get_gamma(t)
        x_den = None
        for i in range(self.T):
            self.step_counter = i
            x, x_den = self.step(x, t[i], t[i + 1], gamma[i], blind)
        self._check_finite(x_den, "x_den")
        if blind:
            st = self._blind.full
            self._check_finite(torch.cat([st["decays"], st["weights"], st["H"].reshape(st["B"], -1)], dim=1),
                               "blind operator (decays / weights / H)")
            self._write_back_operator()
        return x_den.detach()

    def _write_back_operator(self):
        """After a blind run the estimated filter lives in the operator object again (tester.py:161 reads it)."""
        st, op = self._blind.full, self.operator
        H = torch.view_as_complex(st["H"].contiguous())
        if st["B"] == 1:
            op.params[0] = st["decays"].clone()
            op.params[1] = st["weights"].clone()
            op.params_phases[0] = st["phases"][0].clone()
            op.H = H[0].clone()
        # batches: the reference object has ONE filter; the per-utterance results travel as extra attributes
        op.H_batch, op.params_batch = H, (st["decays"], st["weights"], st["phases"])

    def predict_unconditional(self, *args, **kwargs):
        raise ValueError("DPS not made for unconditional sampling")

    def predict_conditional(self, y, operator, shape=None, blind=False, **kwargs):
        if not y.is_cuda:
            raise RuntimeError("buddy_b200 samplers run on CUDA tensors only (no CPU fallback)")
        self.operator = operator
        self.y = y.detach().float().contiguous()
        self._bind_operator(operator, self.y, blind)
        if shape is None:
            shape = y.shape
        return self.predict(shape, y.device, blind)
