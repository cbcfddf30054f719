"""Oracle-only experiment (test infrastructure): backend spread of the REFERENCE algorithm's blind trajectory.

Runs oracle.sampler.dps_blind (the fp32 restatement, unchanged) on the golden T=2 / 20-Adam-iteration case with
different CPU thread counts (different reduction orders inside MKL/oneDNN/pocketfft) and, when a GPU is present, on
CUDA (cuDNN/cuFFT fp32, TF32 off), and prints each run's distance to the reference fixture and to the 1-thread run.
tests/test_gpu_blind.py derives its trajectory bound from the same measurement, in-test."""
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
rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()
noise = [randn(g["step_noise_seed0"] + i, 1, n) for i in range(T + 1)]
rir_noise = [randn(g["rir_noise_seed0"] + i, 13824) for i in range(10 * T)]


def run(dev):
    i = g["init"]
    st = osm.BlindState(i["decays"].to(dev), i["weights"].to(dev), i["phases"].to(dev), i["H"].to(dev))
    sdd = {k: v.to(dev) for k, v in sd.items()}
    pred = osm.dps_blind(sdd, g["y"].to(dev), st, T, [z.to(dev) for z in noise], [z.to(dev) for z in rir_noise])
    return pred.detach().cpu(), torch.view_as_real(st.H.detach()).cpu()


runs = {}
for nt in (1, 2, 4, 8):
    torch.set_num_threads(nt)
    runs[f"cpu{nt}"] = run("cpu")
if torch.cuda.is_available():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    runs["cuda"] = run("cuda")
base = runs["cpu1"]
for k, (p, H) in runs.items():
    print(f"{k:6s} vs reference fixture: pred {rel(p, g['pred']):.2e} H {rel(H, torch.view_as_real(g['final_H'])):.2e}"
          f"   vs cpu1: pred {rel(p, base[0]):.2e} H {rel(H, base[1]):.2e}", flush=True)
