import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_sampler import make_args, edm, randn, rel
from oracle.weights import make_state_dict
from buddy_b200.ncsnpp import NCSNppTime
from buddy_b200.operators import RIROperator
from buddy_b200.samplers import EulerHeunSamplerDPS
from oracle import operators as oop
net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
net.load_state_dict(make_state_dict(0)); net = net.cuda().eval()
n, T, B = 8192, 2, 4
h = torch.stack([randn(60 + b, 2000) * torch.exp(-torch.arange(2000) / (200.0 + 100 * b)) for b in range(B)]).cuda()
s_clean = (randn(70, B, n) * 0.05).cuda()
y = torch.cat([oop.fast_apply_rir(s_clean[b:b + 1], h[b]) for b in range(B)])
noise = [randn(300 + i, B, n).cuda() for i in range(T + 1)]
def run(mb, ns, rows=slice(0, B)):
    smp = EulerHeunSamplerDPS(net, edm(), make_args("informed", T))
    smp.micro_batch, smp.n_streams = mb, ns
    smp.noise_source = iter([z[rows] for z in noise])
    op = RIROperator(); op.update_params(h[rows])
    return smp.predict_conditional(y[rows], op, shape=(y[rows].shape[0], n))
ref = run(4, 1); ref2 = run(4, 1)
print("repeat 4/1:", rel(ref2, ref))
for mb, ns in [(2, 1), (2, 2), (1, 1), (1, 2), (1, 4)]:
    o = run(mb, ns)
    print(f"mb {mb} streams {ns}: ", rel(o, ref), [round(rel(o[b:b+1], ref[b:b+1]), 7) for b in range(B)])
for b in range(B):
    o = run(1, 1, slice(b, b + 1))
    print("alone", b, rel(o, ref[b:b+1]))
