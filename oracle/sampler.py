"""ORACLE (test infrastructure): restatement of the reference samplers, per utterance.

  * create_schedule        — testing/Sampler.py:39-56
  * get_gamma              — testing/EulerHeunSampler.py:24-39
  * denoise (EDM)          — diff_params/shared.py:98-120 + diff_params/edm.py:44-75
  * euler_heun             — testing/EulerHeunSampler.py:41-94 (unconditional; returns x)
  * dps_informed/dps_blind — testing/EulerHeunSamplerDPS.py:25-204 (returns the last x_den)

Noise is explicit: `noise` is the list of N(0,1) draws in the reference's call order (initialize_x first, then one
per stochastic_timestep; for the blind sampler one `randn_like(rir)` per operator-Adam iteration in between).
"""
import math

import torch

from . import net as onet
from . import operators as oop

SIGMA_DATA = 0.05


def create_schedule(T, sigma_min=1e-4, sigma_max=0.5, rho=10):
    a = torch.arange(0, T + 1)
    t = (sigma_max ** (1 / rho) + a / (T - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    t[-1] = 0
    return t


def get_gamma(t, Schurn, Stmin=0, Stmax=10):
    N = t.shape[0]
    g = torch.zeros_like(t)
    idx = torch.logical_and(t > Stmin, t < Stmax)
    g[idx] = min(Schurn / N, 2 ** 0.5 - 1)
    return g


def denoise(sd, x, sigma):
    """x (B,T) fp32, sigma python float or 0-d tensor -> D(x, sigma) (B,T)."""
    sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device)
    s2 = sigma ** 2 + SIGMA_DATA ** 2
    cskip, cout, cin = SIGMA_DATA ** 2 / s2, sigma * SIGMA_DATA * s2 ** -0.5, s2 ** -0.5
    cnoise = (0.25 * torch.log(sigma)).repeat(x.shape[0])
    return cskip * x + cout * onet.ncsnpp_time_forward(sd, (cin * x).unsqueeze(1), cnoise).squeeze(1)


def _perturb(x, t, gamma, eps):
    t_hat = t + gamma * t
    return x + ((t_hat ** 2 - t ** 2) ** 0.5) * eps, t_hat


def euler_heun(sd, shape, T, noise, Schurn=10, order=2):
    t = create_schedule(T)
    gamma = get_gamma(t, Schurn)
    it = iter(noise)
    x = t[0] * next(it)
    with torch.no_grad():
        for i in range(T):
            x_hat, t_hat = _perturb(x, t[i], gamma[i], next(it))
            d = (x_hat - denoise(sd, x_hat, t_hat)) / t_hat
            dt = t[i + 1] - t_hat
            if t[i + 1] != 0 and order == 2:
                x_p = x_hat + dt * d
                d2 = (x_p - denoise(sd, x_p, t[i + 1])) / t[i + 1]
                x = x_hat + dt * 0.5 * (d + d2)
            else:
                x = x_hat + dt * d
    return x


def _likelihood(sd, x_in, sigma, y, degrade, zeta, audio_len, rescale):
    """One DPS evaluation for ONE utterance (x_in (1,T)): returns ode integrand d, and x_den (detached)."""
    x_in = x_in.detach().requires_grad_(True)
    x_den = denoise(sd, x_in, sigma)
    rec = oop.comp_loss(y, degrade(x_den), 512.0).sum()
    g = torch.autograd.grad(rec, x_in)[0]
    lh = zeta / (torch.norm(g) / (audio_len ** 0.5) + 1e-8) * g
    x_den = x_den.detach()
    if rescale:
        x_den = SIGMA_DATA / x_den.std() * x_den
    score = (x_den - x_in.detach()) / sigma ** 2
    return -sigma * score + lh, x_den, x_in.detach()


def dps_informed(sd, y, rir, T, noise, zeta=2.75, Schurn=10, order=2, audio_len=65536, warm="reverb_scaled",
                 rescale=False, degrade_fn=None):
    """y (1,N) observation, rir (M,).  Informed DPS (conf/tester/informed_dereverberation_DPS.yaml).
    rescale = constraint_speech_magnitude.use: applied after the FIRST evaluation of a step only — the reference's
    Heun correction (EulerHeunSamplerDPS.py:136-150) has no rescale line."""
    t = create_schedule(T)
    gamma = get_gamma(t, Schurn)
    it = iter(noise)
    x = t[0] * next(it)
    if warm == "reverb_scaled":
        x = SIGMA_DATA * y.clone() / y.std() + x
    # degrade_fn: any other known operator, e.g. an informed SubbandFiltering (subband_filtering.py:82-101) with fixed H
    degrade = degrade_fn or (lambda v: oop.fast_apply_rir(v, rir))
    x_den = None
    for i in range(T):
        x_hat, t_hat = _perturb(x, t[i], gamma[i], next(it))
        d, x_den, x_hat = _likelihood(sd, x_hat, t_hat, y, degrade, zeta, audio_len, rescale)
        dt = t[i + 1] - t_hat
        if t[i + 1] != 0 and order == 2:
            x_p = x_hat + dt * d
            d2, x_den, _ = _likelihood(sd, x_p, t[i + 1], y, degrade, zeta, audio_len, False)
            x = x_hat + dt * 0.5 * (d + d2)
        else:
            x = x_hat + dt * d
    return x_den


class BlindState:
    """Per-utterance blind-operator state: parameters + Adam moments (EulerHeunSamplerDPS.py:198; torch.optim.Adam)."""

    def __init__(self, decays, weights, phases, H):
        self.decays = decays.clone().requires_grad_(True)      # (1,25)
        self.weights = weights.clone().requires_grad_(True)    # (1,25)
        self.phases = phases.clone().requires_grad_(True)      # (513,100)
        self.H = H.clone()
        self.opt = torch.optim.Adam([self.decays, self.weights, self.phases], lr=0.1, betas=(0.9, 0.99),
                                    weight_decay=0)


def optimize_op(state, x_den, y, t, rir_noise_iter, n_iter=10, crop_max=0.01, crop_min=5e-4):
    """EulerHeunSamplerDPS.optimize_op (:71-113) for one utterance."""
    for _ in range(n_iter):
        state.H = oop.design_H(state.decays, state.weights, state.phases)
        rec = oop.comp_loss(y, oop.blind_degradation(x_den, state.H), 512.0).sum()
        rir = oop.time_rir(state.H)
        t_op = max(min(float(t), crop_max), crop_min)
        reg = oop.comp_loss(rir[None], (rir + t_op * next(rir_noise_iter)).detach()[None], 2560.0).sum()
        state.opt.zero_grad()
        (rec + reg).backward()
        state.opt.step()
        with torch.no_grad():
            d, w = oop.project_params(state.decays, state.weights)
            state.decays.copy_(d)
            state.weights.copy_(w)


def dps_blind(sd, y, state, T, noise, rir_noise, zeta=0.5, Schurn=50, audio_len=65536, warm="reverb_scaled",
              n_iter=10, denoise_fn=None):
    """Blind DPS (conf/tester/blind_dereverberation_BUDDy.yaml; order 1) for ONE utterance y (1,N).
    noise: [init, step0, step1, ...]; rir_noise: iterable of (13824,) draws, one per operator-Adam iteration.
    denoise_fn (sensitivity experiments only): replaces `denoise`, e.g. to perturb the network output at the
    tolerance the network itself is tested to."""
    denoise = denoise_fn or globals()["denoise"]
    t = create_schedule(T)
    gamma = get_gamma(t, Schurn)
    it, rit = iter(noise), iter(rir_noise)
    x = t[0] * next(it)
    if warm == "reverb_scaled":
        x = SIGMA_DATA * y.clone() / y.std() + x
    x_den = None
    for i in range(T):
        x_hat, t_hat = _perturb(x, t[i], gamma[i], next(it))
        x_hat = x_hat.detach().requires_grad_(True)
        x_den = denoise(sd, x_hat, t_hat)
        optimize_op(state, x_den.clone().detach(), y, t_hat, rit, n_iter)
        rec = oop.comp_loss(y, oop.blind_degradation(x_den, state.H.detach()), 512.0).sum()
        g = torch.autograd.grad(rec, x_hat)[0]
        lh = zeta / (torch.norm(g) / (audio_len ** 0.5) + 1e-8) * g
        x_den = x_den.detach()
        x_den = SIGMA_DATA / x_den.std() * x_den
        d = (x_hat.detach() - x_den) / t_hat + lh
        x = x_hat.detach() + (t[i + 1] - t_hat) * d
    return x_den
