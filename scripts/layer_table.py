"""Dev tool (GPU): per-shape device time of every kernel in one network forward + VJP.

    python scripts/layer_table.py [B] [precision]
Conv rows: algorithmic TFLOP/s and fp16-pass-equivalent issue rate.  GN rows: GB/s of algorithmic bytes are not
computed here (see bench.py roofline); the table is for finding which shapes run below par.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from buddy_b200 import ops
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from oracle.weights import make_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16c8"
N = 65536
eng = Engine(make_state_dict(0), "cuda", precision=prec)
st = NetSTFT("cuda")
x = torch.randn(B, N, device="cuda") * 0.2
tc = torch.full((B,), -0.5, device="cuda")
g = torch.randn(B, N, device="cuda")


def fwdbwd():
    out, ctx = eng.forward(st.forward(x), tc, save=True)
    return st.forward_adjoint(eng.vjp(ctx, st.inverse_adjoint(g)), N)


for _ in range(2):
    fwdbwd()
torch.cuda.synchronize()
with ops.KernelTimer() as kt:
    for _ in range(2):
        fwdbwd()
tab = kt.by_tag()
tot = sum(v[1] for v in tab.values())
pe = 2.0 if prec == "fp16c8" else float(eng.np)
print(f"# B={B} {prec}: total {tot / 2:.2f} ms per fwd+vjp ({tot / 2 / B:.3f} ms/utt)")
print(f"{'op':12s} {'tag':58s} {'calls':>5s} {'ms':>9s} {'%':>6s} {'alg TF/s':>9s} {'issued':>8s}")
for (name, tag), (c, ms, w) in sorted(tab.items(), key=lambda kv: -kv[1][1])[:60]:
    tf = w / (ms * 1e-3) / 1e12 if w else 0.0
    print(f"{name:12s} {str(tag):58s} {c // 2:5d} {ms / 2:9.3f} {100 * ms / tot:6.2f} {tf:9.1f} {tf * pe:8.1f}")
