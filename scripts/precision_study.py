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

SCHEME = {"corr_a": "none", "corr_w": "none", "bwd_same": True}


def _emul(a, w, conv, kdim_a, kdim_w):
    ah, wh = q16(a), q16(w)
    out = conv(ah, wh)
    qa = QS[SCHEME["corr_a"]]
    if qa is not None:
        out = out + conv(qa(a - ah, kdim_a), qa(wh, kdim_w))
    qw = QS[SCHEME["corr_w"]]
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
        out = _emul(x, w, lambda a, ww: F.conv2d(a, ww, None, padding=padding), 1, 1)
        return out + b[None, :, None, None] if b is not None else out

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        p = ctx.padding
        if SCHEME.get("exact"):
            return F.conv_transpose2d(g, w, None, padding=p), None, None, None
        # normalise like the engine (per-tensor power of two) so fp16 range is no issue
        s = 2.0 ** math.floor(math.log2(1.0 / g.abs().max().clamp_min(1e-30).item())) * 64.0
        dx = _emul(g * s, w, lambda a, ww: F.conv_transpose2d(a, ww, None, padding=p), 1, 0) / s
        return dx, None, None, None


def patched_conv2d(x, w, b=None, stride=1, padding=0):
    if x.shape[1] <= 2 and False:
        return F.conv2d(x, w, b, padding=padding)
    return QConv.apply(x, w, b, padding)


class _FShim:
    def __getattr__(self, k):
        return patched_conv2d if k == "conv2d" else getattr(F, k)


def run(frames=80):
    torch.manual_seed(0)
    sd = make_state_dict(0)
    g = torch.Generator().manual_seed(1)
    spec = torch.complex(torch.randn(1, 1, 256, frames, generator=g), torch.randn(1, 1, 256, frames, generator=g)) * 3.0
    tc = torch.tensor([0.25 * math.log(0.1)])
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
    for name, ca, cw in [("fp16 1-pass", "none", "none"), ("act corr e4m3", "e4m3", "none"),
                         ("wgt corr e4m3", "none", "e4m3"), ("both e4m3 (fp16c8)", "e4m3", "e4m3"),
                         ("both e2m1 block32", "e2m1", "e2m1"), ("act e2m1 only", "e2m1", "none"),
                         ("both f16 (fp16x3)", "f16", "f16")]:
        SCHEME["corr_a"], SCHEME["corr_w"] = ca, cw
        o, gr = evaluate()
        print(f"{name:24s} fwd {rel(o, ref_o):.2e}  vjp {rel(gr, ref_g):.2e}", flush=True)
    onet.F = F


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 80)
