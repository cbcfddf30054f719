"""Blind operator parity: CUDA kernels vs the oracle restatement (autograd) and the reference fixtures."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def randn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_fft_mixed_25856():
    from buddy_b200 import ops
    from buddy_b200.blind import BlindEngine
    eng = BlindEngine(8192, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    B, N = 2, 25856
    x = torch.randn(B, N, 2, device="cuda", generator=g)
    work, out = torch.empty_like(x), torch.empty_like(x)
    ops.fft_mixed(x, False, work, out, 101, -1, eng.tw512)
    ref = torch.fft.fft(torch.view_as_complex(x))
    assert rel(out, torch.view_as_real(ref)) < 2e-6
    ops.fft_mixed(x, False, work, out, 101, +1, eng.tw512)
    ref = torch.fft.ifft(torch.view_as_complex(x)) * N
    assert rel(out, torch.view_as_real(ref)) < 2e-6
    xr = torch.randn(B, N, device="cuda", generator=g)
    ops.fft_mixed(xr, True, work, out, 101, -1, eng.tw512)
    assert rel(out, torch.view_as_real(torch.fft.fft(xr))) < 2e-6


def test_subband_fir_fwd_bwd():
    from buddy_b200 import ops
    from oracle import operators as oop
    g = torch.Generator(device="cuda").manual_seed(2)
    B, F, T, Nf = 2, 513, 517, 100
    X = torch.randn(B, F, T, 2, device="cuda", generator=g)
    H = torch.randn(B, F, Nf, 2, device="cuda", generator=g)
    Y = ops.subband_fir(X, H, torch.empty_like(X), Nf=Nf, pre=1, mode=0)
    Xc = torch.view_as_complex(X).clone().requires_grad_(True)
    Hc = torch.view_as_complex(H).clone().requires_grad_(True)
    ref = torch.stack([oop.subband_fir(Xc[b:b + 1], Hc[b])[0] for b in range(B)])
    assert rel(Y, torch.view_as_real(ref.detach())) < 1e-5
    dY = torch.randn(B, F, T, 2, device="cuda", generator=g)
    gX, gH = torch.autograd.grad(ref, [Xc, Hc], torch.view_as_complex(dY))
    dX = ops.subband_fir(dY, H, torch.empty_like(X), Nf=Nf, pre=1, mode=1)
    dH = ops.subband_fir(X, dY, torch.empty_like(H), Nf=Nf, pre=1, mode=2)
    assert rel(dX, torch.view_as_real(gX)) < 1e-5
    assert rel(dH, torch.view_as_real(gH)) < 1e-5


def _engine_from_gold(g, B=1):
    from buddy_b200.blind import BlindEngine
    eng = BlindEngine(g["n"], "cuda")
    i = g["init"]
    eng.init_state(B, i["decays"], i["weights"], i["phases"], i["H"])
    eng.select(slice(0, B))
    return eng


def test_update_H_and_time_rir_vs_reference_fixture():
    g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
    from oracle import operators as oop
    eng = _engine_from_gold(g)
    H = eng.update_H()
    i = g["init"]
    ref = oop.design_H(i["decays"], i["weights"], i["phases"])
    assert rel(H[0].cpu(), torch.view_as_real(ref)) < 1e-4
    rir = eng.get_time_RIR()
    assert rel(rir[0].cpu(), g["iter"]["rir"]) < 1e-4


def test_operator_iteration_gradients_vs_reference_fixture():
    """rec + reg losses and their gradients w.r.t. (decays, weights, phases) for one operator iteration."""
    g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
    from buddy_b200 import ops
    it, n = g["iter"], g["n"]
    eng = _engine_from_gold(g)
    x_probe = (g["s"][None] + 0.01 * randn(it["x_probe_seed"], 1, n)).cuda()
    y = g["y"].cuda()
    Y = eng.loss_stft.forward(y)
    noise = randn(it["noise_seed"], 13824).cuda()[None]
    st = eng.state
    p0 = {k: st[k].clone() for k in ("decays", "weights", "phases")}
    grads = []
    orig = ops.adam_project

    def capture(p, gr, m, v, *a):     # intercept the optimiser step: record the gradient, leave parameters alone
        grads.append(gr.clone())
    import buddy_b200.blind as bl
    bl.ops.adam_project = capture
    try:
        hp = dict(iters=1, lr=0.1, beta1=0.9, beta2=0.99, comp=0.667, w_rec=512.0, w_reg=2560.0, crop_max=0.01,
                  crop_min=5e-4)
        eng.optimize(x_probe, Y, 0.5, lambda shape: noise, hp)
    finally:
        bl.ops.adam_project = orig
    rec, reg = eng.last_losses
    assert abs(rec.item() - it["rec"].item()) / it["rec"].item() < 1e-4
    assert abs(reg.item() - it["reg"].item()) / it["reg"].item() < 1e-4
    gd, gw, gp = grads[0], grads[1], grads[2].view(513, 100)
    e = (rel(gd.cpu(), it["g_decays"]), rel(gw.cpu(), it["g_weights"]), rel(gp.cpu(), it["g_phases"]))
    print(f"\n[blind iteration grads] decays {e[0]:.2e} weights {e[1]:.2e} phases {e[2]:.2e}")
    assert max(e) < 1e-3


def test_adam_project_matches_torch():
    from buddy_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, n = 2, 51300
    p = torch.randn(B, n, device="cuda", generator=g)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=0.1, betas=(0.9, 0.99))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    inf = float("inf")
    for step in range(1, 4):
        gr = torch.randn(B, n, device="cuda", generator=g)
        pr.grad = gr.clone()
        opt.step()
        ops.adam_project(p, gr, m, v, step, 0.1, 0.9, 0.99, 1e-8, -inf, inf, -inf, inf)
    assert rel(p, pr.detach()) < 1e-6
    d = torch.full((1, 25), 0.5, device="cuda")
    ops.adam_project(d, torch.full_like(d, -1.0), torch.zeros_like(d), torch.zeros_like(d), 1, 0.1, 0.9, 0.99, 1e-8,
                     0.027632, 0.55264, -inf, inf)
    assert abs(d.max().item() - 0.55264) < 1e-6


def test_blind_dps_trajectory_vs_reference_fixture():
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from oracle import ref_harness as rh
    from oracle.weights import make_state_dict
    g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
    T, n = g["T"], g["n"]
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(make_state_dict(0))
    net = net.cuda().eval()
    smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)),
                              rh.make_args("blind", T))
    step_noise = [randn(g["step_noise_seed0"] + i, 1, n) for i in range(T + 1)]
    rir_noise = [randn(g["rir_noise_seed0"] + i, 13824) for i in range(10 * T)]
    order = [step_noise[0]]
    for i in range(T):
        order.append(step_noise[1 + i])
        order += rir_noise[10 * i:10 * (i + 1)]
    smp.noise_source = iter(order)

    class Op:      # duck-typed stand-in for the reference BlindSubbandFiltering object (only its state is read)
        pass
    op = Op()
    i = g["init"]
    op.params = [i["decays"].clone(), i["weights"].clone()]
    op.params_phases = [i["phases"].clone()]
    op.H = i["H"].clone()
    pred = smp.predict_conditional(g["y"].cuda(), op, shape=(1, n), blind=True)
    e_pred = rel(pred.cpu(), g["pred"])
    e_H = rel(torch.view_as_real(op.H.cpu()), torch.view_as_real(g["final_H"]))
    e_d = rel(op.params[0].cpu(), g["final_decays"])
    e_w = rel(op.params[1].cpu(), g["final_weights"])
    print(f"\n[blind DPS T2] pred {e_pred:.2e}  H {e_H:.2e}  decays {e_d:.2e}  weights {e_w:.2e}")
    # 20 Adam iterations: the optimiser divides each element by its own gradient scale, so elements with noise-level
    # gradients take +-lr steps in an implementation-dependent direction and the trajectory amplifies rounding-level
    # differences.  The bound is therefore MEASURED here, from the reference algorithm itself (the fp32 oracle, run on
    # this GPU with cuDNN/cuFFT, TF32 off): (i) its distance to the CPU reference fixture = backend spread of the
    # unmodified algorithm; (ii) its own response to a perturbation of the denoiser output at the tolerance the
    # network is tested to (5e-4 relative, tests/test_gpu_network.py), several noise seeds.  Our trajectory must lie
    # within twice the largest of those.  (CPU, scripts/blind_backend_spread.py: the oracle with 1 vs 2/4/8 threads
    # differs by 1.5e-3..1.7e-3 (pred) / 3.9e-3..4.6e-3 (H); perturbations of 1e-4..5e-4 move it by 7.7e-4..5.5e-3 /
    # 3.7e-3..1.7e-2.)  Every deterministic stage is pinned tightly elsewhere in this file (one-iteration
    # losses/gradients, update_H 1e-4) and the single-iteration trajectory below holds 1e-3.
    from oracle import sampler as osm
    sdc = {k: v.cuda() for k, v in make_state_dict(0).items()}

    def oracle_run(eps, seed):
        st = osm.BlindState(i["decays"].cuda(), i["weights"].cuda(), i["phases"].cuda(), i["H"].cuda())
        k = [0]

        def dn(sdd, x, sigma):
            d = osm.denoise(sdd, x, sigma)
            if eps:
                z = randn(seed + k[0], *d.shape).cuda()
                k[0] += 1
                d = d + eps * d.detach().norm() / z.norm() * z
            return d
        p = osm.dps_blind(sdc, g["y"].cuda(), st, T, [z.cuda() for z in step_noise], [z.cuda() for z in rir_noise],
                          denoise_fn=dn)
        return p.detach().cpu(), torch.view_as_real(st.H.detach()).cpu()

    base = oracle_run(0.0, 0)
    spread = [(rel(base[0], g["pred"]), rel(base[1], torch.view_as_real(g["final_H"])))]
    for seed in (7, 70, 700, 7000):
        r = oracle_run(5e-4, seed)
        spread.append((rel(r[0], base[0]), rel(r[1], base[1])))
    b_pred, b_H = 2 * max(v[0] for v in spread), 2 * max(v[1] for v in spread)
    print("[blind DPS T2] reference-algorithm spread on this GPU (pred, H): " +
          ", ".join(f"({a:.1e}, {b:.1e})" for a, b in spread) + f" -> bounds {b_pred:.1e} / {b_H:.1e}")
    assert e_pred < max(b_pred, 1e-3) and e_H < max(b_H, 1e-3) and e_d < 5e-3 and e_w < 5e-3
    assert e_pred < 1e-2        # whatever the measured spread: never looser than this


def _blind_case(B, n):
    """B synthetic blind problems with DIFFERENT operator initialisations (decays / weights / phases per utterance)."""
    from buddy_b200.blind import BlindEngine
    y = (randn(900, B, n) * 0.05).cuda()
    decays = torch.stack([torch.full((25,), 0.5 - 0.05 * b) for b in range(B)]).cuda()
    weights = torch.stack([torch.full((25,), 2.0 + 0.2 * b) for b in range(B)]).cuda()
    phases = ((torch.rand(B, 513, 100, generator=torch.Generator().manual_seed(901)) * 2 - 1) * 3.14159).cuda()
    eng = BlindEngine(n, "cuda")
    eng.init_state(B, decays, weights, phases, torch.zeros(B, 513, 100, dtype=torch.complex64))
    eng.select(slice(0, B))
    H = torch.view_as_complex(eng.update_H().contiguous()).clone()
    return y, decays, weights, phases, H


class _Op:      # duck-typed stand-in for the reference BlindSubbandFiltering object (only its state is read)
    def __init__(self, decays, weights, phases, H):
        self.params, self.params_phases, self.H = [decays.clone(), weights.clone()], [phases.clone()], H.clone()


def test_blind_batched_matches_single_utterance_runs():
    """B = 4 blind DPS (T = 2, 10 operator iterations per step) in micro-batches of 2 == every utterance run alone:
    per-utterance H / parameter / Adam-state slicing (`BlindEngine.select`), micro-batch offsets, and the per-utterance
    Philox streams of the step noise and of the RIR-noise regulariser (samplers.optimize_op)."""
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from oracle import ref_harness as rh
    from oracle.weights import make_state_dict
    B, n, T = 4, 8192, 2
    y, decays, weights, phases, H = _blind_case(B, n)
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(make_state_dict(0))
    net = net.cuda().eval()
    edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))

    def run(rows, mb, offset):
        smp = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", T))
        smp.seed_base, smp.utterance_offset, smp.micro_batch = 3000, offset, mb
        op = _Op(decays[rows], weights[rows], phases[rows] if rows.stop - rows.start > 1 else phases[rows][0],
                 H[rows] if rows.stop - rows.start > 1 else H[rows][0])
        pred = smp.predict_conditional(y[rows], op, shape=(rows.stop - rows.start, n), blind=True)
        return pred, op

    pred, op = run(slice(0, B), 2, 0)
    assert op.H_batch.shape == (B, 513, 100)
    for b in range(B):
        p1, o1 = run(slice(b, b + 1), 1, b)
        e, eH = rel(p1, pred[b:b + 1]), rel(torch.view_as_real(o1.H), torch.view_as_real(op.H_batch[b]))
        ed = rel(o1.params[0], op.params_batch[0][b:b + 1])
        print(f"\n[blind batched vs alone, utt {b}] pred {e:.1e} H {eH:.1e} decays {ed:.1e}")
        assert e < 1e-5 and eH < 1e-5 and ed < 1e-5


def test_blind_full_size_trajectory_vs_oracle():
    """The benchmarked blind shape (65 536 samples, BASELINE configs[2]/[3]), B = 2 with different operator
    initialisations, T = 2 with one operator update per step (no chaotic amplification) vs per-utterance oracle runs
    on the GPU."""
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from oracle import ref_harness as rh
    from oracle import sampler as osm
    from oracle.weights import make_state_dict
    B, n, T = 2, 65536, 2
    y, decays, weights, phases, H = _blind_case(B, n)
    sd = make_state_dict(0)
    sdc = {k: v.cuda() for k, v in sd.items()}
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    step_noise = [randn(910 + k, B, n).cuda() for k in range(T + 1)]
    rir_noise = [randn(920 + k, B, 13824).cuda() for k in range(T)]
    args = rh.make_args("blind", T)
    args.tester.posterior_sampling.blind_hp["op_updates_per_step"] = 1
    smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)), args)
    order = [step_noise[0]]
    for k in range(T):
        order += [step_noise[1 + k], rir_noise[k]]
    smp.noise_source = iter(order)
    op = _Op(decays, weights, phases, H)
    pred = smp.predict_conditional(y, op, shape=(B, n), blind=True)
    for b in range(B):
        st = osm.BlindState(decays[b:b + 1], weights[b:b + 1], phases[b], H[b])
        pref = osm.dps_blind(sdc, y[b:b + 1], st, T, [z[b:b + 1] for z in step_noise], [z[b] for z in rir_noise],
                             n_iter=1)
        e = rel(pred[b:b + 1], pref)
        eH = rel(torch.view_as_real(op.H_batch[b]), torch.view_as_real(st.H.detach()))
        print(f"\n[blind DPS full size, utt {b}] pred {e:.2e} H {eH:.2e}")
        assert e < 1e-3 and eH < 5e-3


def test_blind_single_iteration_trajectory_vs_oracle():
    """T=2 blind DPS with ONE operator update per step (no chaotic amplification): 1e-3 on the sampler output."""
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from oracle import ref_harness as rh
    from oracle import sampler as osm
    from oracle.weights import make_state_dict
    g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
    T, n, i = 2, g["n"], g["init"]
    sd = make_state_dict(0)
    sdc = {k: v.cuda() for k, v in sd.items()}
    step_noise = [randn(300 + k, 1, n) for k in range(T + 1)]
    rir_noise = [randn(400 + k, 13824) for k in range(T)]
    st = osm.BlindState(i["decays"].cuda(), i["weights"].cuda(), i["phases"].cuda(), i["H"].cuda())
    pref = osm.dps_blind(sdc, g["y"].cuda(), st, T, [z.cuda() for z in step_noise], [z.cuda() for z in rir_noise],
                         n_iter=1)
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(sd)
    net = net.cuda().eval()
    args = rh.make_args("blind", T)
    args.tester.posterior_sampling.blind_hp["op_updates_per_step"] = 1
    smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)), args)
    order = [step_noise[0]]
    for k in range(T):
        order += [step_noise[1 + k], rir_noise[k]]
    smp.noise_source = iter(order)

    class Op:
        pass
    op = Op()
    op.params, op.params_phases, op.H = [i["decays"].clone(), i["weights"].clone()], [i["phases"].clone()], i["H"].clone()
    pred = smp.predict_conditional(g["y"].cuda(), op, shape=(1, n), blind=True)
    e = rel(pred, pref)
    eH = rel(torch.view_as_real(op.H), torch.view_as_real(st.H.detach()))
    print(f"\n[blind DPS T2, 1 op-iteration/step] pred {e:.2e} H {eH:.2e}")
    # Adam's first steps move every phase element by +-lr whatever the size of its gradient: the ~0.5 % of elements
    # whose gradient is at fp32 rounding level (taps where the filter magnitude is ~1e-6) go in an implementation-
    # dependent direction — 229 of 51300 differ between our own FFT and DFT-matrix STFT forms (scripts/debug_blind_fft.py),
    # which moves H by 2e-3 and the sampler output by 6e-5.  The output is what the 1e-3 tolerance is about.
    assert e < 1e-3 and eH < 5e-3


def test_blind_front_end_matches_single_utterance_runs():
    """tester front-end, blind half (tester.py:147-161): utterances of two lengths, bucketed and batched, with the
    operator initialised per utterance == each utterance run alone from the same initial operator state; the
    estimated time-domain RIRs come back per utterance."""
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSamplerDPS
    from buddy_b200.tester import BatchedDereverb
    from oracle import ref_harness as rh
    from oracle.weights import make_state_dict
    T = 2
    lens = [8192, 6000, 8192]
    ys = [(randn(950 + i, n) * 0.05).cuda() for i, n in enumerate(lens)]
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net.load_state_dict(make_state_dict(0))
    net = net.cuda().eval()
    edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))
    smp = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", T))
    smp.seed_base = 3000
    fe = BatchedDereverb(smp, max_batch=8)
    inits = []
    orig = fe.init_blind_operator

    def recording(B, device, generator=None):
        op = orig(B, device, generator)
        inits.append((op.params[0].clone(), op.params[1].clone(), op.params_phases[0].clone(), op.H.clone()))
        return op
    fe.init_blind_operator = recording
    preds, rirs = fe.blind(ys, generator=torch.Generator().manual_seed(77))
    assert [p.shape[0] for p in preds] == lens and all(r.shape == (13824,) for r in rirs)
    buckets = [[0, 2], [1]]                                     # first-seen order of the two lengths
    for bk, (d0, w0, ph0, H0) in zip(buckets, inits):
        for r, i in enumerate(bk):
            s1 = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", T))
            s1.seed_base, s1.utterance_offset = 3000, i
            op = _Op(d0[r:r + 1], w0[r:r + 1], ph0[r], H0[r])
            alone = s1.predict_conditional(ys[i][None], op, shape=(1, lens[i]), blind=True)[0]
            assert rel(preds[i], alone) < 1e-5, (i, rel(preds[i], alone))
