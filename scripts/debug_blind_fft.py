"""Dev tool (GPU): where does H differ between the FFT and the DFT-matrix STFT forms after 2 single-iteration steps?"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_blind import GOLD, randn, rel
from buddy_b200.edm import EDM
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.samplers import EulerHeunSamplerDPS
from oracle import ref_harness as rh
from oracle.weights import make_state_dict
g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
T, n, i = 2, g["n"], g["init"]
sd = make_state_dict(0)
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(sd); net = net.cuda().eval()
step_noise = [randn(300 + k, 1, n) for k in range(T + 1)]
rir_noise = [randn(400 + k, 13824) for k in range(T)]
res = {}
for mode in ("dft", "fft"):
    os.environ["BUDDY_STFT"] = mode
    args = rh.make_args("blind", T)
    args.tester.posterior_sampling.blind_hp["op_updates_per_step"] = 1
    smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)), args)
    order = [step_noise[0]]
    for k in range(T):
        order += [step_noise[1 + k], rir_noise[k]]
    smp.noise_source = iter(order)
    class Op: pass
    op = Op()
    op.params, op.params_phases, op.H = [i["decays"].clone(), i["weights"].clone()], [i["phases"].clone()], i["H"].clone()
    pred = smp.predict_conditional(g["y"].cuda(), op, shape=(1, n), blind=True)
    res[mode] = (pred, torch.view_as_real(op.H.clone()), smp._blind.full["phases"].clone(), smp._blind.full["decays"].clone(), smp._blind.full["weights"].clone())
a, b = res["dft"], res["fft"]
print("pred", rel(b[0], a[0]), "H", rel(b[1], a[1]), "phases", rel(b[2], a[2]), "decays", rel(b[3], a[3]), "weights", rel(b[4], a[4]))
dph = (b[2] - a[2]).abs()[0]            # [513, 100]
print("phase diff > 0.05 count", int((dph > 0.05).sum()), "of", dph.numel())
idx = torch.nonzero(dph > 0.05)
print("bins with flips:", sorted(set(idx[:, 0].tolist()))[:40])
print("taps with flips:", sorted(set(idx[:, 1].tolist()))[:40])
Hd = (b[1] - a[1]).pow(2).sum(-1).sum(-1)[0] if b[1].dim() == 4 else None
