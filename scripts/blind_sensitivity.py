"""CPU experiment (oracle only, test infrastructure): how sensitive is the REFERENCE algorithm's blind trajectory to
perturbations at fp32 rounding level?  Runs oracle.sampler.dps_blind on the golden T=2 case twice: as is, and with the
observation y perturbed by `eps` relative white noise, and prints the relative L2 distance of the outputs / filters.
If a 1e-7 perturbation moves the output by ~1e-3..1e-2, no fp32 implementation can pin the 20-Adam-iteration
trajectory tighter than that (each implementation rounds differently)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sampler as osm
from oracle.weights import make_state_dict
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
g = torch.load(os.path.join(GOLD, "sampler_blind_T2.pt"), weights_only=False)
sd = make_state_dict(0)
T, n = g["T"], g["n"]
randn = lambda seed, *s: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
noise = [randn(g["step_noise_seed0"] + i, 1, n) for i in range(T + 1)]
rir_noise = [randn(g["rir_noise_seed0"] + i, 13824) for i in range(10 * T)]

def run(eps):
    st = osm.BlindState(g["init"]["decays"], g["init"]["weights"], g["init"]["phases"], g["init"]["H"])
    y = g["y"] * (1.0 + eps * randn(99, *g["y"].shape)) if eps else g["y"]
    pred = osm.dps_blind(sd, y, st, T, noise, rir_noise)
    return pred.detach(), torch.view_as_real(st.H.detach()), st.decays.detach(), st.weights.detach()

base = run(0.0)
print("oracle vs reference fixture: pred %.2e H %.2e" % (rel(base[0], g["pred"]), rel(base[1], torch.view_as_real(g["final_H"]))))
for eps in (1e-7, 1e-6, 1e-5):
    r = run(eps)
    print("eps %.0e: pred %.2e  H %.2e  decays %.2e  weights %.2e" % ((eps,) + tuple(rel(a, b) for a, b in zip(r, base))), flush=True)
