"""Dev tool (GPU): rms of every fp16 dgrad operand (and forward operand) in one fwd+VJP, to size fp8 scales."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
from oracle.weights import make_state_dict
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from buddy_b200 import ops, engine as E

sig = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
eng = Engine(make_state_dict(0), "cuda", precision="fp16")
st = NetSTFT("cuda")
N = 65536
x = torch.randn(1, N, device="cuda") * math.sqrt(sig * sig + 0.0025) / math.sqrt(sig * sig + 0.0025)
tc = torch.full((1,), 0.25 * math.log(sig), device="cuda")
rec = []
orig_bwd, orig_apply = ops.gn_bwd, ops.gn_apply
def bwd(*a, **k):
    r = orig_bwd(*a, **k)
    for key in ("g16a", "g16b"):
        if k.get(key) is not None:
            t = k[key].float()
            rec.append(("bwd " + key, tuple(t.shape), t.pow(2).mean().sqrt().item(), t.abs().max().item()))
    return r
def app(*a, **k):
    r = orig_apply(*a, **k)
    t = a[4].float()
    rec.append(("fwd act", tuple(t.shape), t.pow(2).mean().sqrt().item(), t.abs().max().item()))
    if k.get("out_raw") is not None:
        t = k["out_raw"].float()
        rec.append(("fwd raw", tuple(t.shape), t.pow(2).mean().sqrt().item(), t.abs().max().item()))
    return r
E.ops.gn_bwd, E.ops.gn_apply = bwd, app
spec = st.forward(x)
out, ctx = eng.forward(spec, tc, save=True)
g = torch.randn(1, N, device="cuda")
g = g / g.pow(2).mean().sqrt()
d = eng.vjp(ctx, st.inverse_adjoint(g))
import collections
for kind in ("fwd act", "fwd raw", "bwd g16a"):
    vals = [(r[2], r[3]) for r in rec if r[0] == kind]
    rms = [v[0] for v in vals]; mx = [v[1] for v in vals]
    print(f"sigma={sig} {kind}: n={len(vals)} rms min {min(rms):.3e} max {max(rms):.3e}; absmax max {max(mx):.3e}")
print("bwd g16a rms per call:", " ".join(f"{r[2]:.2e}" for r in rec if r[0] == "bwd g16a"))
