"""Dev tool (GPU): run N network forward+VJP evaluations (target for ncu kernel filters)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from oracle.weights import make_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16c8"
N = 65536
eng = Engine(make_state_dict(0), "cuda", precision=prec)
st = NetSTFT("cuda")
x = torch.randn(B, N, device="cuda") * 0.2
tc = torch.full((B,), -0.5, device="cuda")
g = torch.randn(B, N, device="cuda")
for _ in range(reps):
    out, ctx = eng.forward(st.forward(x), tc, save=True)
    st.forward_adjoint(eng.vjp(ctx, st.inverse_adjoint(g)), N)
torch.cuda.synchronize()
print("done")
