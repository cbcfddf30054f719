"""Dev tool (GPU): per-module relative error of the CUDA engine vs the oracle along the network depth."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import net as onet
from oracle.weights import make_state_dict
from buddy_b200.engine import Engine
from buddy_b200 import engine as E

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sd = make_state_dict(0)
sdc = {k: v.cuda() for k, v in sd.items()}
eng = Engine(sd, "cuda")
x = (torch.randn(1, 1, N, generator=torch.Generator().manual_seed(6)) * 0.2).cuda()
tc = torch.tensor([0.25 * torch.log(torch.tensor(0.1))]).cuda()
spec = onet.net_stft(x)
onet.RECORD = {}
with torch.no_grad():
    ref = onet.ncsnpp_forward(sdc, spec, tc)
rec = onet.RECORD
onet.RECORD = None

# capture engine block outputs by wrapping the block drivers
got = {}
orig_rb, orig_attn, orig_head = Engine._rb_fwd, Engine._attn_fwd, Engine._head_fwd
def rb(self, i, *a, **k):
    out = orig_rb(self, i, *a, **k); got[i] = out[0]; return out
def at(self, *a, **k):
    out = orig_attn(self, *a, **k); got[self.attn_idx] = out[0]; return out
def hd(self, i, *a, **k):
    out = orig_head(self, i, *a, **k); got[i + 1] = out; return out
Engine._rb_fwd, Engine._attn_fwd, Engine._head_fwd = rb, at, hd
s = torch.view_as_real(spec[:, 0].contiguous()).contiguous()
out, ctx = eng.forward(s, tc, save=True)
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
for i in sorted(rec):
    if i in got:
        print(f"module {i:2d}: rel {rel(got[i], rec[i].permute(0, 2, 3, 1)):.2e}   shape {tuple(got[i].shape)}")
print("final:", rel(out, torch.view_as_real(ref[:, 0].contiguous())))
