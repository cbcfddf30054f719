"""Dev tool (GPU): ONE network forward + data-gradient at B = 1 (after warm-up) — target of an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from oracle.weights import make_state_dict
eng = Engine(make_state_dict(0), "cuda", precision="mixed")
st = NetSTFT("cuda")
x = torch.randn(1, 65536, device="cuda") * 0.2
tc = torch.full((1,), -0.5, device="cuda")
g = torch.randn(1, 65536, device="cuda")
for k in range(3):
    if k == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    out, ctx = eng.forward(st.forward(x), tc, save=True)
    st.forward_adjoint(eng.vjp(ctx, st.inverse_adjoint(g)), 65536)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
