"""Pins the oracle (oracle/, a PyTorch restatement) to fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import os

import pytest
import torch

from oracle import net as onet
from oracle import operators as oop
from oracle import sampler as osm
from oracle.weights import make_state_dict, param_spec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def sd():
    torch.set_num_threads(os.cpu_count())
    return make_state_dict(0)


def test_state_dict_contract():
    g = gold("state_dict_spec.pt")
    assert [(k, tuple(s)) for k, s in g["keys"]] == param_spec()


def test_net_stft_istft():
    g = gold("net_stft.pt")
    x = randn(g["seed"], *g["shape"])
    spec = onet.net_stft(x)
    assert spec.shape == g["spec"].shape == (1, 1, 256, 32)
    assert rel(torch.view_as_real(spec), torch.view_as_real(g["spec"])) < 1e-6
    assert rel(onet.net_istft(spec, 3000), g["istft"]) < 1e-6
    assert g["spec_full_frames"] == 528


def test_network_forward_and_vjp(sd):
    g = gold("net_small.pt")
    x = (randn(g["x_seed"], 2, 1, 8192) * g["x_scale"]).requires_grad_(True)
    cot = randn(g["cot_seed"], 2, 1, 8192)
    out = onet.ncsnpp_time_forward(sd, x, 0.25 * torch.log(g["sigma"]))
    (vjp,) = torch.autograd.grad((out * cot).sum(), x)
    assert rel(out.detach(), g["out"]) < 1e-5
    assert rel(vjp, g["vjp"]) < 1e-5


def test_edm_denoiser(sd):
    g = gold("edm_denoiser.pt")
    x = randn(g["x_seed"], 1, 8192) * g["x_scale"]
    with torch.no_grad():
        out = osm.denoise(sd, x, g["sigma"])
    assert rel(out, g["out"]) < 1e-5


def test_schedule_gamma():
    g = gold("schedule.pt")
    for key, schurn in (("informed_35", 10), ("blind_60", 50), ("informed_3", 10), ("blind_2", 50)):
        T = int(key.split("_")[1])
        t = osm.create_schedule(T)
        assert torch.equal(t, g[key]["t"]), key
        assert torch.equal(osm.get_gamma(t, schurn), g[key]["gamma"]), key


def test_operators_and_loss():
    g = gold("operators.pt")
    s, h, y = g["s"], g["h"], g["y"]
    assert rel(oop.fast_apply_rir(s[None], h), y) < 1e-6
    assert rel(torch.view_as_real(oop.loss_stft(s[None])), torch.view_as_real(g["loss_stft"])) < 1e-6
    xh = (s + 0.01 * randn(g["pert_seed"], g["n"]))[None].requires_grad_(True)
    l = oop.comp_loss(y, oop.fast_apply_rir(xh, h), 512.0).sum()
    (gr,) = torch.autograd.grad(l, xh)
    assert abs(l.item() - g["loss"].item()) / abs(g["loss"].item()) < 1e-5
    assert rel(gr, g["loss_grad"]) < 1e-4


def test_blind_operator_chain():
    g = gold("operators.pt")
    A = oop.design_magnitude(g["blind_decays"], g["blind_weights"])
    assert rel(A, g["blind_A"]) < 1e-5
    H = oop.design_H(g["blind_decays"], g["blind_weights"], g["blind_phases"])
    assert rel(torch.view_as_real(H), torch.view_as_real(g["blind_H"])) < 1e-4
    assert rel(torch.angle(g["blind_H_init"]), g["blind_phases"]) < 1e-6
    assert rel(oop.blind_degradation(g["s"][None], g["blind_H"]), g["blind_y"]) < 1e-5
    assert rel(oop.time_rir(g["blind_H"]), g["blind_rir"]) < 1e-5
    hin = randn(g["minphase_in_seed"], 640) * torch.exp(-torch.arange(640) / 80.0)
    assert rel(oop.minimum_phase(hin), g["minphase"]) < 1e-5


def test_audio_known_answer():
    """The reference's only known-answer data: reverberant == gain * (clean (*) rir)."""
    g = gold("audio_kat_p226.pt")
    y = oop.fast_apply_rir(g["clean"][None], g["rir"])[0]
    gain = (y @ g["reverberant"]) / (y @ y)
    assert rel(gain * y, g["reverberant"]) < 1e-5


def test_unconditional_sampler(sd):
    g = gold("sampler_uncond_T3.pt")
    noise = [randn(g["noise_seed0"] + i, 1, g["n"]) for i in range(g["T"] + 1)]
    x = osm.euler_heun(sd, (1, g["n"]), g["T"], noise)
    assert rel(x, g["x"]) < 1e-4


def test_informed_dps_sampler(sd):
    g = gold("sampler_informed_T3.pt")
    noise = [randn(g["noise_seed0"] + i, 1, g["n"]) for i in range(g["T"] + 1)]
    pred = osm.dps_informed(sd, g["y"], g["h"], g["T"], noise)
    assert rel(pred, g["pred"]) < 1e-3


def test_informed_dps_order2_with_magnitude_constraint(sd):
    """Order 2 + constraint_speech_magnitude: the rescale follows the first evaluation of a step only."""
    g = gold("sampler_informed_T2_rescale.pt")
    noise = [randn(g["noise_seed0"] + i, 1, g["n"]) for i in range(g["T"] + 1)]
    pred = osm.dps_informed(sd, g["y"], g["h"], g["T"], noise, rescale=True)
    assert rel(pred, g["pred"]) < 1e-3


def test_blind_dps_sampler(sd):
    """Blind path incl. the filter-design chain, Adam, projection and the RIR-noise regulariser.  The 27->513 band
    interpolation is the torchcde stand-in on BOTH sides (parity unpinned at that one call, see oracle/__init__.py)."""
    g = gold("sampler_blind_T2.pt")
    T, n = g["T"], g["n"]
    st = osm.BlindState(g["init"]["decays"], g["init"]["weights"], g["init"]["phases"], g["init"]["H"])
    noise = [randn(g["step_noise_seed0"] + i, 1, n) for i in range(T + 1)]
    rir_noise = [randn(g["rir_noise_seed0"] + i, 13824) for i in range(10 * T)]
    # one operator iteration: loss values and gradients w.r.t. all parameters (tight)
    it = g["iter"]
    x_probe = g["s"][None] + 0.01 * randn(it["x_probe_seed"], 1, n)
    H = oop.design_H(st.decays, st.weights, st.phases)
    rec = oop.comp_loss(g["y"], oop.blind_degradation(x_probe, H), 512.0).sum()
    rir = oop.time_rir(H)
    reg = oop.comp_loss(rir[None], (rir + it["t_op"] * randn(it["noise_seed"], 13824)).detach()[None], 2560.0).sum()
    gd, gw, gp = torch.autograd.grad(rec + reg, [st.decays, st.weights, st.phases])
    assert abs(rec.item() - it["rec"].item()) / it["rec"].item() < 1e-5
    assert abs(reg.item() - it["reg"].item()) / it["reg"].item() < 1e-5
    assert rel(rir.detach(), it["rir"]) < 1e-5
    assert rel(gd, it["g_decays"]) < 1e-4 and rel(gw, it["g_weights"]) < 1e-4 and rel(gp, it["g_phases"]) < 1e-4
    # 2 sampler steps = 20 Adam iterations.  Adam divides every element by its own gradient scale, so elements whose
    # gradient is at rounding-noise level move by +-lr in an implementation-dependent direction: after 20 iterations
    # two fp32 executions of the SAME algorithm differ by 1e-3..3e-3 in the output and ~1e-2 in the filter as soon as
    # the summation order changes (another thread count, another CPU's SIMD kernels: measured 9.9e-4 / 3.1e-3 / 2.5e-3
    # with 1 / 2 / 4 threads in one container, < 1e-3 with 8 threads in another).  The bound is therefore measured
    # here, from the restatement's own spread on the running machine: (i) 1 thread vs all threads, (ii) its response
    # to an observation perturbed at fp32 rounding level (1e-7 relative).  Floor: the 1e-3 tolerance; cap: 1e-2.
    def run(y, threads):
        torch.set_num_threads(threads)
        s0 = osm.BlindState(g["init"]["decays"], g["init"]["weights"], g["init"]["phases"], g["init"]["H"])
        p = osm.dps_blind(sd, y, s0, T, noise, rir_noise)
        torch.set_num_threads(os.cpu_count())
        return p, s0

    pred, st = run(g["y"], os.cpu_count())
    pred_1, st_1 = run(g["y"], 1)
    pred_p, st_p = run(g["y"] * (1 + 1e-7 * randn(999, *g["y"].shape)), os.cpu_count())
    H = lambda s0: torch.view_as_real(s0.H.detach())
    spread_pred = max(rel(pred_1, pred), rel(pred_p, pred))
    spread_H = max(rel(H(st_1), H(st)), rel(H(st_p), H(st)))
    print(f"blind T=2 oracle vs reference fixture: pred {rel(pred, g['pred']):.2e} (1 thread {rel(pred_1, g['pred']):.2e}), "
          f"own spread pred {spread_pred:.2e} / H {spread_H:.2e}")
    assert rel(st.decays.detach(), g["final_decays"]) < 5e-3
    assert rel(st.weights.detach(), g["final_weights"]) < 5e-3
    assert rel(H(st), torch.view_as_real(g["final_H"])) < min(max(2e-2, 2 * spread_H), 4e-2)
    assert min(rel(pred, g["pred"]), rel(pred_1, g["pred"])) < min(max(1e-3, 2 * spread_pred), 1e-2)


def test_upfirdn2d_oracle_vs_reference_native():
    """oracle/upfirdn.py against outputs of the reference's own `upfirdn2d_native` (op/upfirdn2d.py:157-200)."""
    from oracle import upfirdn as ou
    for c in gold("upfirdn2d.pt"):
        x, k = randn(c["x_seed"], *c["shape"]), randn(c["k_seed"], c["taps"], c["taps"])
        out = ou.upfirdn2d(x, k, c["up"], c["down"], c["pad"])
        # same formula, same operand order: identical up to the summation order of the CPU's convolution kernel
        # (bit-equal on the container the fixture was made in, 3e-9 of the output range on another CPU model)
        assert out.shape == c["out"].shape
        assert ((out - c["out"]).abs().max() / c["out"].abs().max()).item() < 1e-6
