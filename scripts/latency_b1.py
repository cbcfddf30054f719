"""Dev tool (GPU): single-utterance (B=1, the reference's own usage) DPS evaluation: wall time vs summed kernel time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200 import ops, _capi
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from oracle.weights import make_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = 65536
eng = Engine(make_state_dict(0), "cuda")
st = NetSTFT("cuda")
x = torch.randn(B, N, device="cuda") * 0.2
tc = torch.full((B,), -0.5, device="cuda")
g = torch.randn(B, N, device="cuda")
GRAPH = os.environ.get("NOGRAPH") is None
def ev():
    out, ctx = eng.forward(st.forward(x), tc, save=True, graph=GRAPH)
    y = st.inverse(out, N)
    return st.forward_adjoint(eng.vjp(ctx, st.inverse_adjoint(g)), N)
for _ in range(3): ev()
torch.cuda.synchronize()
_capi.reset_launch_count()
t0 = time.perf_counter(); n = 10
for _ in range(n): ev()
t_issue = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
t_wall = (time.perf_counter() - t0) / n
launches = _capi.launch_count() / n
print(f"B={B} graphs={GRAPH}: wall {1e3 * t_wall:.2f} ms per fwd+VJP, host issue {1e3 * t_issue:.2f} ms, "
      f"{launches:.0f} buddy launches issued from the host per evaluation")
