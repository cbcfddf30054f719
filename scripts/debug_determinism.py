import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.weights import make_state_dict
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT, LossSTFT, RirConv
from buddy_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
B = 2
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
eng = Engine(make_state_dict(0), "cuda", precision=prec)
st = NetSTFT("cuda")
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, n, generator=g) * 0.2).cuda()
tc = torch.full((B,), -0.5, device="cuda")
cot = torch.randn(B, n, generator=g).cuda()
def ev():
    spec = st.forward(x)
    out, ctx = eng.forward(spec, tc, save=True)
    y = st.inverse(out, n)
    d = eng.vjp(ctx, st.inverse_adjoint(cot))
    return spec, out, y, d, st.forward_adjoint(d, n)
a = ev()
for it in range(3):
    # perturb allocator state between runs
    junk = [torch.randn(1 << 22, device="cuda") for _ in range(it + 1)]
    b = ev()
    print(it, "spec %.2e out %.2e y %.2e dspec %.2e dx %.2e" % tuple(rel(u, v) for u, v in zip(b, a)))
    del junk
