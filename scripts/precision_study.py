"""CPU emulation of tensor-core operand precision schemes on the oracle network (development aid, not product).

Every conv of oracle.net is replaced by an autograd Function whose forward AND data-gradient are computed from
quantised operands the way the CUDA engine would issue them:  main = q16(a) * q16(w)  plus optional first-order
corrections  c_a = qa(a - q16(a)) * qw(q16(w))  and  c_w = qa(q16(a)) * qw(w - q16(w))  with qa/qw in
{e4m3, e2m1 block-scaled (32 along K)}.  Prints relative L2 error of net output and VJP against plain fp32.

    python scripts/precision_study.py [frames]
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import net as onet
from oracle.weights import make_state_dict


def q16(x):
    return x.half().float()


def qe4m3(x, dim=1):
    """per-tensor power-of-two scale to the e4m3 window, then round"""
    m = x.abs().max().clamp_min(1e-30)
    s = 2.0 ** math.floor(math.log2(256.0 / m.item()))
    return (x * s).clamp(-448, 448).to(torch.float8_e4m3fn).float() / s


E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])


def qe2m1_block(x, dim=1, block=32):
    """MXFP4-style: ue8m0 scale per `block` consecutive elements along `dim`, e2m1 values, round to nearest."""
    xm = x.movedim(dim, -1)
    shp = xm.shape
    K = shp[-1]
    pad = (-K) % block
    if pad:
        xm = F.pad(xm, (0, pad))
    xb = xm.reshape(*xm.shape[:-1], -1, block)
    mx = xb.abs().amax(-1, keepdim=True).clamp_min(1e-30)
    sc = torch.exp2(torch.ceil(torch.log2(mx / 6.0)))
    y = (xb / sc).clamp(-6, 6)
    idx = (y.abs()[..., None] - E2M1).abs().argmin(-1)
    q = E2M1[idx] * torch.sign(y) * sc
    q = q.reshape(*xm.shape)[..., :K].reshape(shp)
    return q.movedim(-1, dim)


QS = {"e4m3": qe4m3, "e2m1": qe2m1_block, "f16": lambda x, dim=1: q16(x), "none": None}

SCHEME = {"corr_a": "none", "corr_w": "none", "bwd_same": True, "policy": None}
COST = {"flops": 0.0, "passes": 0.0}
PASS_COST = {"none": 0.0, "e4m3": 0.5, "e2m1": 0.25, "f16": 1.0}


def _emul(a, w, conv, kdim_a, kdim_w, H, direction):
    ca, cw = SCHEME["corr_a"], SCHEME["corr_w"]
    if SCHEME["policy"] is not None:
        SCHEME["w_shape"] = tuple(w.shape[:2])      # (Cout, Cin) of the conv weight, both directions
        ca, cw = SCHEME["policy"](H, direction)
    ah, wh = q16(a), q16(w)
    out = conv(ah, wh)
    fl = float(out.numel()) * w.shape[1 if kdim_w == 1 else 0] * w.shape[2] * w.shape[3]
    COST["flops"] += fl
    COST["passes"] += fl * (1.0 + PASS_COST[ca] + PASS_COST[cw])
    qa = QS[ca]
    if qa is not None:
        out = out + conv(qa(a - ah, kdim_a), qa(wh, kdim_w))
    qw = QS[cw]
    if qw is not None:
        out = out + conv(qw(ah, kdim_a), qw(w - wh, kdim_w))
    return out


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding):
        ctx.save_for_backward(w)
        ctx.padding = padding
        ctx.xshape = x.shape
        if SCHEME.get("exact"):
            return F.conv2d(x, w, b, padding=padding)
        out = _emul(x, w, lambda a, ww: F.conv2d(a, ww, None, padding=padding), 1, 1, x.shape[2], 'fwd')
        return out + b[None, :, None, None] if b is not None else out

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        p = ctx.padding
        if SCHEME.get("exact"):
            return F.conv_transpose2d(g, w, None, padding=p), None, None, None
        # normalise like the engine (per-tensor power of two) so fp16 range is no issue
        s = 2.0 ** math.floor(math.log2(1.0 / g.abs().max().clamp_min(1e-30).item())) * 64.0
        dx = _emul(g * s, w, lambda a, ww: F.conv_transpose2d(a, ww, None, padding=p), 1, 0, g.shape[2], 'bwd') / s
        return dx, None, None, None


def patched_conv2d(x, w, b=None, stride=1, padding=0):
    if x.shape[1] <= 2 and False:
        return F.conv2d(x, w, b, padding=padding)
    return QConv.apply(x, w, b, padding)


class _FShim:
    def __getattr__(self, k):
        return patched_conv2d if k == "conv2d" else getattr(F, k)


def run(frames=80, seed=0, sigma=0.1):
    torch.manual_seed(seed)
    sd = make_state_dict(seed)
    g = torch.Generator().manual_seed(1 + seed)
    spec = torch.complex(torch.randn(1, 1, 256, frames, generator=g), torch.randn(1, 1, 256, frames, generator=g)) * 3.0
    tc = torch.tensor([0.25 * math.log(sigma)])
    cot = torch.randn(1, 2, 256, frames, generator=g)

    def evaluate():
        re = spec.real.clone().requires_grad_(True)
        im = spec.imag.clone()
        out = onet.ncsnpp_forward(sd, torch.complex(re, im), tc)
        o = torch.cat([out.real, out.imag], 1)
        (gr,) = torch.autograd.grad((o * cot).sum(), re)
        return o.detach(), gr

    SCHEME["exact"] = True
    ref_o, ref_g = evaluate()
    SCHEME["exact"] = False
    onet.F = _FShim()
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    c8 = ("e4m3", "e4m3")
    x1 = ("none", "none")
    a8 = ("e4m3", "none")
    w8 = ("none", "e4m3")
    BIG = {(256, 256, 256), (256, 128, 384), (256, 128, 256), (128, 256, 512)}        # (H, Cout, Cin)
    MID = {(128, 256, 384), (128, 256, 256), (256, 128, 128)}
    big = lambda H: (H,) + SCHEME["w_shape"] in BIG
    mid = lambda H: (H,) + SCHEME["w_shape"] in MID
    policies = {
        "A: top5 x1, rest c8": lambda H, d: x1 if big(H) else c8,
        "A2: top5 a8, rest c8": lambda H, d: a8 if big(H) else c8,
        "B: top5 x1, mid a8, rest c8": lambda H, d: x1 if big(H) else (a8 if mid(H) else c8),
        "C: top5 x1 fwd / a8 bwd, rest c8": lambda H, d: (x1 if d == "fwd" else a8) if big(H) else c8,
        "D: top5+mid x1, rest c8": lambda H, d: x1 if (big(H) or mid(H)) else c8,
        "E: fwd top5 x1 / bwd top5+mid x1": lambda H, d: x1 if (big(H) or (d == "bwd" and mid(H))) else c8,
        "F: fwd top5 / bwd top5+(256,128,128)": lambda H, d: x1 if (big(H) or (d == "bwd" and (H,) + SCHEME["w_shape"] == (256, 128, 128))) else c8,
        "G: fwd top5 / bwd top5+(128,256,256)": lambda H, d: x1 if (big(H) or (d == "bwd" and (H,) + SCHEME["w_shape"] == (128, 256, 256))) else c8,
        "H: fwd top5 / bwd ALL x1": lambda H, d: x1 if (big(H) or d == "bwd") else c8,
    }
    import os
    if os.environ.get("ONLY"):
        policies = {k: v for k, v in policies.items() if k[0] in os.environ["ONLY"]}
    if seed != 0:
        policies = {"all x1": lambda H, d: x1, "all c8": lambda H, d: c8, "A: top5 x1, rest c8": policies["A: top5 x1, rest c8"]}
    for name, pol in policies.items():
        SCHEME["policy"] = pol
        COST["flops"] = COST["passes"] = 0.0
        o, gr = evaluate()
        print(f"{name:32s} fwd {rel(o, ref_o):.2e}  vjp {rel(gr, ref_g):.2e}  passes {COST['passes'] / COST['flops']:.3f}",
              flush=True)
    onet.F = F


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 80, int(sys.argv[2]) if len(sys.argv) > 2 else 0,
        float(sys.argv[3]) if len(sys.argv) > 3 else 0.1)
