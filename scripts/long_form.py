"""Dev tool (GPU): one 30 s utterance (480000 samples -> 256 x 3760 spectrogram, attention over 15040 tokens) through
the network forward + VJP; prints time and peak memory (BASELINE configs[4] feasibility)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from oracle.weights import make_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 480000
eng = Engine(make_state_dict(0), "cuda")
st = NetSTFT("cuda")
x = torch.randn(B, N, device="cuda") * 0.2
tc = torch.full((B,), -0.5, device="cuda")
g = torch.randn(B, N, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    spec = st.forward(x)
    out, ctx = eng.forward(spec, tc, save=True)
    y = st.inverse(out, N)
    d = st.forward_adjoint(eng.vjp(ctx, st.inverse_adjoint(g)), N)
    torch.cuda.synchronize()
    print(f"B={B} N={N} spec {tuple(spec.shape)}: fwd+vjp {1e3 * (time.time() - t0):.1f} ms, peak mem "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, finite {bool(torch.isfinite(d).all())}", flush=True)
