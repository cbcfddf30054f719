"""Dev tool (GPU): ONE blind operator update (10 Adam iterations) at B = 1 — target of an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200.edm import EDM
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.samplers import EulerHeunSamplerDPS
from buddy_b200.tester import BatchedDereverb
from oracle import ref_harness as rh
from oracle.weights import make_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(make_state_dict(0))
net = net.cuda().eval()
edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))
y = (torch.randn(B, 65536, generator=torch.Generator().manual_seed(0)) * 0.05).cuda()
smp = EulerHeunSamplerDPS(net, edm, rh.make_args("blind", 60))
smp.seed_base, smp.micro_batch = 3000, B
fe = BatchedDereverb(smp, max_batch=B)
op = fe.init_blind_operator(B, "cuda", torch.Generator().manual_seed(1))
smp.operator, smp.y = op, y
smp._bind_operator(op, y, True)
smp._start_run()
xd = y * 0.9
for k in range(3):
    if k == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    smp.optimize_op(xd, 0.3, slice(0, B))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
