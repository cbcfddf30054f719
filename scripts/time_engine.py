"""Dev tool (GPU): wall/device time of one network forward / VJP and of one DPS evaluation."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.weights import make_state_dict
from buddy_b200.engine import Engine
from buddy_b200.spectral import NetSTFT
from buddy_b200 import _capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
N = 65536
t0 = time.time()
sd = make_state_dict(0)
print("make_state_dict", time.time() - t0)
t0 = time.time()
eng = Engine(sd, "cuda", precision=prec)
torch.cuda.synchronize()
print("Engine pack", time.time() - t0)
st = NetSTFT("cuda")
x = torch.randn(B, N, device="cuda") * 0.2
tc = torch.full((B,), -0.5, device="cuda")
g = torch.randn(B, N, device="cuda")

def ev(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.time() - t0) / iters * 1e3

def fwd():
    spec = st.forward(x)
    out, ctx = eng.forward(spec, tc, save=True)
    return st.inverse(out, N), ctx
def fwdbwd():
    y, ctx = fwd()
    d = eng.vjp(ctx, st.inverse_adjoint(g))
    return st.forward_adjoint(d, N)
_capi.reset_launch_count()
print("fwd     ms (device, wall):", ev(fwd))
print("fwd+vjp ms (device, wall):", ev(fwdbwd))
print("launches per fwd+vjp ~", _capi.launch_count() / 8)
print("per-utterance DPS eval ms:", ev(fwdbwd)[0] / B, " -> TFLOP/s algorithmic:", 2578.7e9 * B / (ev(fwdbwd)[0] * 1e-3) / 1e12)
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
from buddy_b200 import ops
with ops.KernelTimer() as kt:
    fwdbwd()
s = kt.summary()
tot = sum(v[1] for v in s.values())
print(f"--- per-op device time for one fwd+vjp (B={B}, {prec}), total {tot:.2f} ms")
for k, (c, ms, w) in sorted(s.items(), key=lambda kv: -kv[1][1]):
    extra = f"  {w / (ms * 1e-3) / 1e12:8.1f} TFLOP/s (issued)" if w else ""
    print(f"{k:18s} calls {c:4d}  {ms:9.3f} ms  {100 * ms / tot:5.1f}%{extra}")
