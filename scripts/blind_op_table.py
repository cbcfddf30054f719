"""Dev tool (GPU): per-kernel device time of the blind operator update (10 Adam iterations) + likelihood gradient.
    python scripts/blind_op_table.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200 import ops
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
smp.seed_base = 3000
smp.micro_batch = B
fe = BatchedDereverb(smp, max_batch=B)
op = fe.init_blind_operator(B, "cuda", torch.Generator().manual_seed(1))
smp.operator, smp.y = op, y
smp._bind_operator(op, y, True)
smp._start_run()
xd = y * 0.9
for _ in range(2):
    smp.optimize_op(xd, 0.3, slice(0, B))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    smp.optimize_op(xd, 0.3, slice(0, B))
e1.record()
torch.cuda.synchronize()
print(f"# B={B}: optimize_op (10 iterations) {e0.elapsed_time(e1) / 3:.3f} ms wall (no timers)")
with ops.KernelTimer() as kt:
    smp.optimize_op(xd, 0.3, slice(0, B))
tab = kt.by_tag()
tot = sum(v[1] for v in tab.values())
print(f"# with per-kernel events: {tot:.3f} ms over {sum(v[0] for v in tab.values())} launches")
agg = {}
for (name, tag), (c, ms, w) in tab.items():
    a = agg.setdefault(name, [0, 0.0])
    a[0] += c
    a[1] += ms
for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:22s} {c:5d} launches {ms:9.3f} ms {100 * ms / tot:6.1f} %  {1e3 * ms / c:8.1f} us each")
