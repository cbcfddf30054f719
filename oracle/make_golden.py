"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden
Every fixture stores the seeds/inputs needed to regenerate its input plus the reference's output, so the
restatement in oracle/ and the CUDA path can be checked where the reference itself cannot travel.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NS = 8192  # small utterance: 65 frames -> 80 padded


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def synth_utterance(seed, n):
    """Speech surrogate per SURVEY.md §8d: 1-pole low-passed white noise scaled to sigma_data."""
    w = randn(seed, n).numpy().astype(np.float64)
    s = np.zeros(n)
    acc = 0.0
    for i in range(n):
        acc = 0.95 * acc + w[i]
        s[i] = acc
    s = torch.from_numpy(s).float()
    return 0.05 * s / s.std()


def synth_rir(seed, m, t60):
    h = randn(seed, m) * torch.exp(-6.908 * torch.arange(m) / (t60 * 16000))
    h[0] = 1.0
    return h / h.abs().max()


def save(name, obj):
    path = os.path.join(GOLD, name)
    torch.save(obj, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    rh.install()
    os.makedirs(GOLD, exist_ok=True)
    sd = make_state_dict(0)
    net = rh.build_network(sd)
    edm = rh.build_edm()

    # ---- state_dict contract
    ref_sd = rh.build_network().state_dict()
    save("state_dict_spec.pt", {"keys": [(k, tuple(v.shape)) for k, v in ref_sd.items()]})

    # ---- network STFT / iSTFT (ncsnpp.py:473-496)
    x = randn(21, 1, 1, 3000)
    spec = net.stft(x)
    save("net_stft.pt", {"seed": 21, "shape": (1, 1, 3000), "spec": spec, "istft": net.istft(spec.clone(), 3000),
                         "spec_full_frames": net.stft(randn(22, 1, 1, 65536)).shape[-1]})

    # ---- network forward + VJP at the small shape
    xs = randn(11, 2, 1, NS) * 0.5
    sig = torch.tensor([0.3, 0.01])
    tc = 0.25 * torch.log(sig)
    cot = randn(12, 2, 1, NS)
    xs_r = xs.clone().requires_grad_(True)
    out = net(xs_r, tc)
    (g,) = torch.autograd.grad((out * cot).sum(), xs_r)
    save("net_small.pt", {"x_seed": 11, "x_scale": 0.5, "cot_seed": 12, "sigma": sig, "out": out.detach(), "vjp": g})

    # ---- EDM denoiser (shared.py:98-120)
    xd = randn(13, 1, NS) * 0.3
    with torch.no_grad():
        den = edm.denoiser(xd.unsqueeze(1), net, torch.tensor(0.2)).squeeze(1)
    save("edm_denoiser.pt", {"x_seed": 13, "x_scale": 0.3, "sigma": 0.2, "out": den})

    # ---- schedule / gamma (Sampler.py:39-56, EulerHeunSampler.py:24-39)
    from testing.EulerHeunSampler import EulerHeunSampler
    sch = {}
    for mode, T in (("informed", 35), ("blind", 60), ("informed", 3), ("blind", 2)):
        s = EulerHeunSampler(net, edm, rh.make_args(mode, T))
        t = s.create_schedule()
        sch[f"{mode}_{T}"] = {"t": t, "gamma": s.get_gamma(t)}
    save("schedule.pt", sch)

    # ---- operators + loss (reverb.py, subband_filtering.py, losses.py, reverb_utils.py)
    from testing.operators.reverb import RIROperator
    from testing.operators.subband_filtering import BlindSubbandFiltering
    from utils.losses import get_loss
    from utils import reverb_utils
    args = rh.make_args("blind", 2)
    hp = args.tester.informed_dereverberation.op_hp
    n_op = 4096
    s = synth_utterance(31, n_op)
    h = synth_rir(32, 1500, 0.4)
    op = RIROperator(hp, time_kernel_size=h.shape[-1], sample_rate=16000)
    op.update_params(h)
    y = op.degradation(s[None])
    loss = get_loss(args.tester.posterior_sampling.rec_loss, operator=op)
    xh = (s + 0.01 * randn(33, n_op))[None].requires_grad_(True)
    lval = loss(y, op.degradation(xh))
    (lgrad,) = torch.autograd.grad(lval, xh)
    X = op.apply_stft(s[None])
    ops_gold = {"n": n_op, "s_seed": 31, "h_seed": 32, "h_len": 1500, "h_t60": 0.4, "pert_seed": 33, "s": s, "h": h,
                "y": y, "loss_stft": X, "loss": lval.detach(), "loss_grad": lgrad}
    torch.manual_seed(5)
    bop = BlindSubbandFiltering(hp, sample_rate=16000)
    with torch.no_grad():
        bop.update_H(use_noise=True)
    H_init = bop.H.detach().clone()
    with torch.no_grad():
        bop.update_H()
    ops_gold.update({"blind_H_init": H_init, "blind_decays": bop.params[0].detach().clone(), "blind_weights": bop.params[1].detach().clone(),
                     "blind_phases": bop.params_phases[0].detach().clone(), "blind_H": bop.H.detach().clone(),
                     "blind_y": bop.degradation(s[None]).detach(), "blind_rir": bop.get_time_RIR().detach(),
                     "blind_A": bop.design_filter().detach(),
                     "minphase_in_seed": 34,
                     "minphase": reverb_utils.minimum_phase_version(randn(34, 640) * torch.exp(-torch.arange(640) / 80.0))})
    save("operators.pt", ops_gold)

    # ---- known-answer data shipped with the reference (audio_examples/: reverberant = g * clean (*) rir)
    import wave

    def read_wav(path):
        import soundfile  # may be the stub
        raise RuntimeError

    def read_f32_wav(path):
        with open(path, "rb") as f:
            b = f.read()
        i = b.find(b"data")
        n = int.from_bytes(b[i + 4:i + 8], "little")
        fmt = b.find(b"fmt ")
        code = int.from_bytes(b[fmt + 8:fmt + 10], "little")
        bits = int.from_bytes(b[fmt + 22:fmt + 24], "little")
        raw = b[i + 8:i + 8 + n]
        if code == 3 and bits == 32:
            return torch.from_numpy(np.frombuffer(raw, dtype="<f4").copy())
        if code == 1 and bits == 16:
            return torch.from_numpy(np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0)
        raise RuntimeError(f"wav format {code}/{bits}")

    ex = os.path.join(rh.REF_ROOT, "audio_examples")
    clean = read_f32_wav(os.path.join(ex, "clean/p226/p226_003.wav"))
    rir = read_f32_wav(os.path.join(ex, "rir/p226/p226_003.wav"))
    rev = read_f32_wav(os.path.join(ex, "reverberant/p226/p226_003.wav"))
    ncrop = 16384
    save("audio_kat_p226.pt", {"clean": clean[:ncrop].clone(), "rir": rir.clone(), "reverberant": rev[:ncrop].clone(),
                               "note": "first 16384 samples of audio_examples/*/p226/p226_003.wav; "
                                       "reverberant == g * fast_apply_RIR(clean, rir) up to a scalar gain g"})

    # ---- samplers with injected noise
    from testing.EulerHeunSamplerDPS import EulerHeunSamplerDPS
    T = 3
    s = synth_utterance(41, NS)
    h = synth_rir(42, 2000, 0.5)
    op = RIROperator(hp, time_kernel_size=h.shape[-1], sample_rate=16000)
    op.update_params(h)
    y = op.degradation(s[None])
    noise = [randn(100 + i, 1, NS) for i in range(T + 1)]
    smp = EulerHeunSamplerDPS(net, edm, rh.make_args("informed", T, audio_len=65536))
    with rh.injected_noise(noise):
        pred = smp.predict_conditional(y, op, shape=(1, NS), blind=False)
    save("sampler_informed_T3.pt", {"T": T, "n": NS, "s_seed": 41, "h_seed": 42, "h_len": 2000, "h_t60": 0.5,
                                    "noise_seed0": 100, "s": s, "h": h, "y": y, "pred": pred})
    smu = EulerHeunSampler(net, edm, rh.make_args("unconditional", T))
    with rh.injected_noise(noise):
        xu = smu.predict_unconditional((1, NS), "cpu")
    save("sampler_uncond_T3.pt", {"T": T, "n": NS, "noise_seed0": 100, "x": xu})

    # ---- blind DPS (order 1, 10 operator-Adam iterations per step) with injected noise
    T = 2
    s = synth_utterance(61, NS)
    h = synth_rir(62, 2000, 0.5)
    op = RIROperator(hp, time_kernel_size=h.shape[-1], sample_rate=16000)
    op.update_params(h)
    y = op.degradation(s[None])
    torch.manual_seed(7)
    bop = BlindSubbandFiltering(hp, sample_rate=16000)
    with torch.no_grad():
        bop.update_H(use_noise=True)
    init = {"decays": bop.params[0].detach().clone(), "weights": bop.params[1].detach().clone(),
            "phases": bop.params_phases[0].detach().clone(), "H": bop.H.detach().clone()}
    # single operator-iteration gradients (strong pin: the multi-iteration Adam trajectory amplifies rounding)
    for k in range(2):
        bop.params[k].requires_grad = True
    bop.params_phases[0].requires_grad = True
    bop.update_H()
    x_probe = s[None] + 0.01 * randn(63, 1, NS)
    n_probe = randn(64, 13824)
    bargs = rh.make_args("blind", T)
    lrec = get_loss(bargs.tester.posterior_sampling.rec_loss_params, operator=bop)
    lreg = get_loss(bargs.tester.posterior_sampling.RIR_noise_regularization.loss, operator=bop)
    rec_v = lrec(y, bop.degradation(x_probe, mode="waveform"))
    rir_v = bop.get_time_RIR()
    reg_v = lreg(rir_v, (rir_v + 0.01 * n_probe).detach())
    gr = torch.autograd.grad(rec_v + reg_v, [bop.params[0], bop.params[1], bop.params_phases[0]])
    iter_gold = {"x_probe_seed": 63, "noise_seed": 64, "t_op": 0.01, "rec": rec_v.detach(), "reg": reg_v.detach(),
                 "g_decays": gr[0], "g_weights": gr[1], "g_phases": gr[2], "rir": rir_v.detach()}
    for k in range(2):
        bop.params[k].requires_grad = False
    bop.params_phases[0] = bop.params_phases[0].detach()
    with torch.no_grad():
        bop.update_H()
    step_noise = [randn(300 + i, 1, NS) for i in range(T + 1)]
    rir_noise = [randn(400 + i, 13824) for i in range(10 * T)]
    # reference call order: initialize_x, then per step: stochastic_timestep, 10 x randn_like(rir)
    order = [step_noise[0]]
    for i in range(T):
        order.append(step_noise[1 + i])
        order += rir_noise[10 * i:10 * (i + 1)]
    smp = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", T, audio_len=65536))
    with rh.injected_noise(order):
        pred = smp.predict_conditional(y, bop, shape=(1, NS), blind=True)
    save("sampler_blind_T2.pt", {"T": T, "n": NS, "y": y, "init": init, "step_noise_seed0": 300, "rir_noise_seed0": 400,
                                 "pred": pred, "s": s, "iter": iter_gold, "final_decays": bop.params[0].detach().clone(),
                                 "final_weights": bop.params[1].detach().clone(),
                                 "final_phases": bop.params_phases[0].detach().clone(),
                                 "final_H": bop.H.detach().clone()})


if __name__ == "__main__":
    main()
