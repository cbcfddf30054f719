"""Dev tool (GPU): one blind DPS step, CUDA path vs oracle, with a configurable number of operator iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import sampler as osm, operators as oop, ref_harness as rh
from oracle.weights import make_state_dict
from buddy_b200.edm import EDM
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.samplers import EulerHeunSamplerDPS

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.load("tests/golden/sampler_blind_T2.pt", weights_only=False)
n = g["n"]
rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()
def randn(seed, *shape): return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
sd = make_state_dict(0)
sdc = {k: v.cuda() for k, v in sd.items()}
i = g["init"]
step_noise = [randn(300 + k, 1, n) for k in range(T + 1)]
rir_noise = [randn(400 + k, 13824) for k in range(n_iter * T)]
# oracle on GPU
st = osm.BlindState(i["decays"].cuda(), i["weights"].cuda(), i["phases"].cuda(), i["H"].cuda())
pref = osm.dps_blind(sdc, g["y"].cuda(), st, T, [z.cuda() for z in step_noise], [z.cuda() for z in rir_noise], n_iter=n_iter)
# ours
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(sd); net = net.cuda().eval()
args = rh.make_args("blind", T)
args.tester.posterior_sampling.blind_hp["op_updates_per_step"] = n_iter
smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)), args)
order = [step_noise[0]]
for k in range(T):
    order.append(step_noise[1 + k]); order += rir_noise[n_iter * k:n_iter * (k + 1)]
smp.noise_source = iter(order)
class Op: pass
op = Op(); op.params = [i["decays"].clone(), i["weights"].clone()]; op.params_phases = [i["phases"].clone()]; op.H = i["H"].clone()
pred = smp.predict_conditional(g["y"].cuda(), op, shape=(1, n), blind=True)
print(f"n_iter={n_iter} T={T}: pred {rel(pred, pref):.2e} H {rel(torch.view_as_real(op.H), torch.view_as_real(st.H.detach())):.2e} "
      f"decays {rel(op.params[0], st.decays.detach()):.2e} weights {rel(op.params[1], st.weights.detach()):.2e} "
      f"phases {rel(op.params_phases[0], st.phases.detach()):.2e}")
