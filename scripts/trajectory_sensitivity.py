"""Experiment (GPU; oracle + CUDA path): how do complete 35-step informed trajectories compare, and how sensitive is
the REFERENCE algorithm itself (fp32 oracle) over 69 network evaluations?

Prints (a) CUDA path vs oracle after T = 3, 5, 10, 20, 35 steps, (b) the oracle's own distance to a run whose
NETWORK output is perturbed by relative white noise eps at every evaluation (eps = 3.6e-4, the measured forward error
of the CUDA network), (c) the oracle with TF32 convolutions — PyTorch's default on a GPU, i.e. what the unmodified
reference does there — vs strict fp32.  Random-init (non-degenerate) weights, 0.5 s utterance.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from buddy_b200.edm import EDM
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.operators import RIROperator
from buddy_b200.samplers import EulerHeunSamplerDPS
from oracle import operators as oop, ref_harness as rh, sampler as osm
from oracle.weights import make_state_dict

rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
randn = lambda seed, *s: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
sd = make_state_dict(0)
sdc = {k: v.cuda() for k, v in sd.items()}
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(sd)
net = net.cuda().eval()
n = 8192
h = (randn(1, 2000) * torch.exp(-6.908 * torch.arange(2000) / (0.4 * 16000))).cuda()
h[0] = 1
y = oop.fast_apply_rir((randn(2, 1, n) * 0.05).cuda(), h)
edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))
from oracle import net as onet
base_net = onet.ncsnpp_time_forward


def oracle(T, noise, eps=0.0, seed=0):
    k = [0]

    def fwd(sdd, x, tc):
        f = base_net(sdd, x, tc)
        if eps:
            z = randn(seed + k[0], *f.shape).to(f.device)
            k[0] += 1
            f = f + eps * f.detach().norm() / z.norm() * z
        return f
    onet.ncsnpp_time_forward = fwd
    try:
        return osm.dps_informed(sdc, y, h, T, noise).detach()
    finally:
        onet.ncsnpp_time_forward = base_net


for T in (3, 5, 10, 20, 35):
    noise = [randn(100 + i, 1, n).cuda() for i in range(T + 1)]
    smp = EulerHeunSamplerDPS(net, edm, rh.make_args("informed", T))
    smp.noise_source = iter(noise)
    op = RIROperator()
    op.update_params(h)
    ours = smp.predict_conditional(y, op, shape=(1, n))
    ref = oracle(T, noise)
    pert = [rel(oracle(T, noise, 3.6e-4, s), ref) for s in (7, 70)]
    torch.backends.cudnn.allow_tf32 = True
    tf32 = rel(oracle(T, noise), ref)
    torch.backends.cudnn.allow_tf32 = False
    print(f"T={T:2d}: CUDA path vs oracle {rel(ours, ref):.2e} | oracle vs oracle(+3.6e-4 on the network output) "
          f"{pert[0]:.2e} {pert[1]:.2e} | oracle with TF32 convolutions vs fp32 {tf32:.2e}", flush=True)
