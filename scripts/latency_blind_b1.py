"""Dev tool (GPU): B = 1 blind BUDDy run (the reference's own usage through test.py): wall time of a full
predict_conditional(blind=True), T = 60, 10 operator iterations per step, reverb_scaled and wpe_scaled warm starts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200.edm import EDM
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.samplers import EulerHeunSamplerDPS
from buddy_b200.tester import BatchedDereverb
from oracle import ref_harness as rh
from oracle.weights import make_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(make_state_dict(0))
net = net.cuda().eval()
edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))
y = (torch.randn(B, 65536, generator=torch.Generator().manual_seed(0)) * 0.05).cuda()
for warm in ("reverb_scaled", "wpe_scaled"):
    smp = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", T, warm=warm))
    smp.seed_base = 3000
    fe = BatchedDereverb(smp, max_batch=B)
    for rep in range(2):
        op = fe.init_blind_operator(B, "cuda", torch.Generator().manual_seed(1))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pred = smp.predict_conditional(y, op, shape=(B, 65536), blind=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"B={B} T={T} warm={warm}: {dt:.3f} s per run = {1e3 * dt / T:.1f} ms per step, {B / dt:.2f} utterances/s", flush=True)
