"""General module walk for the NCSN++ variants other than the shipped output_skip / input_skip graph
(`progressive` residual | none, `progressive_input` residual | none; reference networks/ncsnpp.py:196-274 builds the
module list, :340-445 runs it).  Same kernels as engine.py — the ResBlock / attention / head drivers are reused as they
are — but the schedule is data-driven: `netspec.plan()` lists the modules, the forward pass records a tape, and the
data-gradient pass walks the tape backwards with fp32 gradients accumulated per tensor (the shipped graph keeps its
hand-scheduled walk in engine.py, which passes fp16 gradient operands from block to block instead).
"""
import torch

from . import netspec, ops, upfirdn2d
from .engine import INV_SQRT2, NF, _RB
from .ops import MODE_DOWN, MODE_NONE, MODE_UP

_RB_KINDS = ("rb_d", "rb_m", "rbcat", "rbdown", "rbup")


class Val:
    """fp32 channels-last tensor + (lazily computed) GroupNorm bundle statistics."""
    __slots__ = ("t", "s")

    def __init__(self, t, s=None):
        self.t, self.s = t, s


def _st(v):
    if v.s is None:
        v.s = ops.gn_stats(v.t)
    return v.s


# ---------------------------------------------------------------------------------------------------- packing
def build(eng, resblock_type, progressive, progressive_input):
    sd = eng.sd
    eng.plan, n_modules = netspec.plan(resblock_type, progressive, progressive_input)
    eng.out_from_pyramid = progressive == "output_skip"
    eng.gnconv = {}
    for kind, i, cin, cout, lvl in eng.plan:
        p = f"all_modules.{i}."
        if kind in _RB_KINDS:
            eng._pack_rb(i, lvl)
        elif kind in ("down", "up", "pyrup") or (kind == "pyrdown" and cin != 2):
            eng._pack_resample(i)
        elif kind == "pyrdown":
            # first pyramid level: conv3x3(2 -> C, stride 2) as an im2col GEMM, like the input convolution
            w = sd[p + "Conv_0.weight"].contiguous()
            m = _RB()
            m.cin, m.cout = 2, cout
            m.w = eng._packv(w, 1, cout, 64, sn=(cout, 0, 18), sk=(2, 1, 9), k_valid=18)
            m.wd = eng._packv(w, 1, 32, cout, sn=(2, 1, 9), sk=(cout, 0, 18), n_valid=18)
            m.bias = sd[p + "Conv_0.bias"].contiguous()
            eng.resamp[i] = m
        elif kind == "combine":
            eng.comb[i] = (sd[p + "Conv_0.weight"].reshape(-1, 2).contiguous(), sd[p + "Conv_0.bias"].contiguous())
        elif kind == "attn":
            eng.attn_idx = i
            eng._pack_attn(i)
        elif kind in ("head", "head_add", "final"):
            eng._pack_head(i)
        elif kind == "gnconv":
            m = _RB()
            m.g, m.b = sd[p + "weight"].contiguous(), sd[p + "bias"].contiguous()
            m.c = cin
            m.w, m.wd = eng._pack3x3(sd[f"all_modules.{i + 1}.weight"])
            m.bias = sd[f"all_modules.{i + 1}.bias"].contiguous()
            eng.gnconv[i] = m


# ---------------------------------------------------------------------------------------------------- helpers
def _coef(eng, B, v):
    return torch.full((B,), v, device=eng.device)


def _halfsum(eng, a, b):
    """(a + b) / sqrt(2)   (skip_rescale, ncsnpp.py:362-365, 413-416)."""
    B = a.shape[0]
    c = _coef(eng, B, INV_SQRT2)
    return ops.lincomb3(torch.empty(B, a[0].numel(), device=eng.device), a.view(B, -1), c, b.view(B, -1), c).view(a.shape)


def _scale32(eng, a, s):
    B = a.shape[0]
    return ops.lincomb3(torch.empty(B, a[0].numel(), device=eng.device), a.view(B, -1), _coef(eng, B, s)).view(a.shape)


def _cast(eng, x32, key, scale=1.0, need8=True):
    """fp32 gradient -> tensor-core operand of scale * x32 (times the calibrated power-of-two scale of this site)."""
    B, H, W, C = x32.shape
    op = eng._operand(B, H, W, C, eng._gscale(key), need8=need8)
    ops.cast_operand(x32, op.t16, op.t8, scale=scale * op.gs, split=eng.split)
    eng._record(key, op)
    return op


def _pick_odd(eng, full):
    return upfirdn2d._launch(full, eng._k1, (1, 1), (2, 2), (-1, 0, -1, 0))


def _down2_fwd(eng, i, x2):
    """Downsample conv of the 2-channel input (pad right/bottom, 3x3, stride 2; layerspp.py:147-154)."""
    m = eng.resamp[i]
    B, H, W, _ = x2.shape
    col = eng._operand(B, H, W, 64)
    ops.im2col_c2(x2, col.t16, split=eng.split, col8=col.t8)
    full = torch.empty(B, H, W, m.cout, device=eng.device)
    eng._conv(col, m.w, full, taps=1, n_total=m.cout, bias=m.bias)
    return _pick_odd(eng, full)


def _down2_bwd(eng, i, g, scale):
    m = eng.resamp[i]
    B, h, w, _ = g.shape
    zs = upfirdn2d._launch(g, eng._k1, (2, 2), (1, 1), (1, -1, 1, -1))
    op = _cast(eng, zs, ("dn", i), scale)
    dcol = torch.empty(B, 2 * h, 2 * w, 32, device=eng.device)
    eng._conv(op, m.wd, dcol, taps=1, n_total=32)
    return ops.col2im_c2(dcol, torch.empty(B, 2 * h, 2 * w, 2, device=eng.device))


# ---------------------------------------------------------------------------------------------------- forward
def forward(eng, spec, time_cond, save=True):
    B, H, W, _ = spec.shape
    dev = eng.device
    tb = eng.time_bias(time_cond)
    ctx = {} if save else None
    tape = []
    x0 = Val(spec)
    pyr = x0                    # input pyramid
    hs, h, pyramid, out2 = [], None, None, None
    for kind, i, cin, cout, lvl in eng.plan:
        if kind == "inconv":
            col = eng._operand(B, H, W, 64)
            ops.im2col_c2(spec, col.t16, split=eng.split, col8=col.t8)
            t = torch.empty(B, H, W, NF, device=dev)
            sh = eng._zeros_stats(B, NF)
            eng._conv(col, eng.in_w, t, taps=1, n_total=NF, bias=eng.in_b, stats=sh)
            h = Val(t, sh)
            hs.append(h)
            tape.append(("inconv", i, (x0,), h))
        elif kind in _RB_KINDS:
            xa = hs[-1] if kind in ("rb_d", "rbdown") else h
            xb = hs.pop() if kind == "rbcat" else None
            mode = MODE_DOWN if kind == "rbdown" else (MODE_UP if kind == "rbup" else MODE_NONE)
            t, so = eng._rb_fwd(i, xa.t, _st(xa), xb.t if xb else None, _st(xb) if xb else None, tb, mode, ctx)
            h = Val(t, so)
            if kind == "rb_d":
                hs.append(h)
            tape.append(("rb", i, (xa, xb), h))
        elif kind == "down":
            x = hs[-1]
            h = Val(eng._down_fwd(i, x.t))
            tape.append(("down", i, (x,), h))
        elif kind == "up":
            t, so = eng._up_fwd(i, h.t)
            o = Val(t, so)
            tape.append(("up", i, (h,), o))
            h = o
        elif kind == "combine":
            p = pyr.t
            if eng.fir:       # pyramid_downsample = Downsample(fir=True, with_conv=False): downsample_2d
                p2 = Val(eng._fir_fwd(p, False))
            else:
                p2 = Val(ops.resample_c2(p, 0, torch.empty(B, p.shape[1] // 2, p.shape[2] // 2, 2, device=dev)))
            tape.append(("avgpool", None, (pyr,), p2))
            pyr = p2
            w, b = eng.comb[i]
            hc = Val(ops.combine_fwd(h.t, pyr.t, w, b, torch.empty_like(h.t)))
            tape.append(("combine", i, (h, pyr), hc))
            h = hc
            hs.append(h)
        elif kind == "pyrdown":
            t = _down2_fwd(eng, i, pyr.t) if cin == 2 else eng._down_fwd(i, pyr.t)
            m = Val(_halfsum(eng, t, h.t))
            tape.append(("pyrdown2" if cin == 2 else "pyrdown", i, (pyr, h), m))
            pyr = h = m
            hs.append(h)
        elif kind == "push":
            hs.append(h)
        elif kind == "attn":
            t, so = eng._attn_fwd(h.t, _st(h), ctx)
            o = Val(t, so)
            tape.append(("attn", i, (h,), o))
            h = o
        elif kind in ("head", "head_add", "final"):
            ph = Val(eng._head_fwd(i, h.t, _st(h), ctx))
            tape.append(("head", i, (h,), ph))
            if kind == "head":
                pyramid = ph
            elif kind == "final":
                out2 = ph
            else:
                if eng.fir:   # pyramid_upsample = Upsample(fir=True, with_conv=False): upsample_2d
                    pn = Val(eng._add32(eng._fir_fwd(pyramid.t, True), ph.t))
                else:
                    pn = Val(ops.resample_c2(pyramid.t, 1, torch.empty_like(ph.t), add=ph.t))
                tape.append(("upadd", None, (pyramid, ph), pn))
                pyramid = pn
        elif kind == "gnconv":
            m = eng.gnconv[i]
            C = m.c
            a = eng._operand(*h.t.shape[:3], C)
            ops.gn_apply(h.t, _st(h), m.g, m.b, a.t16, silu=True, split=eng.split, out8=a.t8)
            t = torch.empty(*h.t.shape[:3], C, device=dev)
            eng._conv(a, m.w, t, taps=9, n_total=C, bias=m.bias)
            pyramid = Val(t)
            tape.append(("gnconv", i, (h,), pyramid))
        elif kind == "pyrup":
            t, _ = eng._up_fwd(i, pyramid.t)
            m = Val(_halfsum(eng, t, h.t))
            tape.append(("pyrup", i, (pyramid, h), m))
            pyramid = h = m
        else:
            raise AssertionError(kind)
    assert not hs
    last = pyramid if eng.out_from_pyramid else out2
    out = ops.affine_c2(last.t, eng.out_m, eng.out_b, torch.empty_like(last.t))
    if save:
        ctx.update(tape=tape, last=last, x0=x0, shape=(B, H, W))
    return out, ctx


# ---------------------------------------------------------------------------------------------------- data-gradient
def vjp(eng, ctx, dout):
    B, H, W = ctx["shape"]
    dev = eng.device
    assert dout.shape == (B, H, W, 2) and dout.is_contiguous()
    # unit-rms cotangent per utterance, undone on the result (as in Engine._vjp_impl)
    rs = ops.row_stats(dout.view(B, -1))
    rms = torch.sqrt(rs[:, 1] / (H * W * 2)).float().clamp_min(1e-30)
    dout = ops.lincomb3(torch.empty(B, H * W * 2, device=dev), dout.view(B, -1), (1.0 / rms).contiguous()).view(B, H, W, 2)
    G = {}

    def add(v, g):
        if v is None or g is None:
            return
        k = id(v)
        G[k] = g if k not in G else eng._add32(G[k], g)

    add(ctx["last"], ops.affine_c2(dout, eng.out_mT, [0.0, 0.0], torch.empty_like(dout)))
    for kind, i, ins, out in reversed(ctx["tape"]):
        g = G.pop(id(out), None)
        if g is None:
            continue
        if kind == "head":
            h, = ins
            dx, _ = eng._head_bwd(i, h.t, _st(h), g, None, want32=True)
            add(h, dx)
        elif kind == "upadd":
            pold, ph = ins
            add(ph, g)
            if eng.fir:
                add(pold, eng._fir_bwd(g, True, pold.t.shape[1], pold.t.shape[2]))
            else:
                add(pold, ops.resample_c2(g, 3, torch.empty(B, g.shape[1] // 2, g.shape[2] // 2, 2, device=dev)))
        elif kind == "rb":
            xa, xb = ins
            g16 = _cast(eng, g, ("gx", i), INV_SQRT2, need8=not eng.rb[i].x1b[1])
            dxa, _, dxb = eng._rb_bwd(i, ctx, g16, g, want_a32=True, want_a16=False)
            add(xa, dxa)
            add(xb, dxb)
        elif kind == "attn":
            g16 = _cast(eng, g, ("ga",), INV_SQRT2, need8=False)
            dx, _ = eng._attn_bwd(ctx, g16, g)
            add(ins[0], dx)
        elif kind == "down":
            dx, _ = eng._down_bwd(i, g, None, None, want_g=False)
            add(ins[0], dx)
        elif kind == "up":
            add(ins[0], eng._up_bwd(i, _cast(eng, g, ("gu", i))))
        elif kind == "avgpool":
            if eng.fir:
                add(ins[0], eng._fir_bwd(g, False, ins[0].t.shape[1], ins[0].t.shape[2]))
            else:
                add(ins[0], ops.resample_c2(g, 2, torch.empty(B, 2 * g.shape[1], 2 * g.shape[2], 2, device=dev)))
        elif kind == "combine":
            h, p = ins
            add(h, g)
            add(p, ops.combine_bwd(g, eng.comb[i][0], torch.empty(B, g.shape[1], g.shape[2], 2, device=dev)))
        elif kind in ("pyrdown", "pyrdown2"):
            pin, h = ins
            add(h, _scale32(eng, g, INV_SQRT2))
            if kind == "pyrdown2":
                add(pin, _down2_bwd(eng, i, g, INV_SQRT2))
            else:
                dx, _ = eng._down_bwd(i, g, None, None, scale=INV_SQRT2, want_g=False)
                add(pin, dx)
        elif kind == "gnconv":
            h, = ins
            m = eng.gnconv[i]
            da = torch.empty(*g.shape[:3], m.c, device=dev)
            eng._conv(_cast(eng, g, ("gg", i)), m.wd, da, taps=9, n_total=m.c)
            dx = torch.empty_like(h.t)
            ops.gn_bwd(h.t, _st(h), m.g, m.b, da, eng._scratch_gsum(B), silu=True, dxa=dx)
            add(h, dx)
        elif kind == "pyrup":
            pold, h = ins
            add(h, _scale32(eng, g, INV_SQRT2))
            add(pold, eng._up_bwd(i, _cast(eng, g, ("gp", i), INV_SQRT2)))
        elif kind == "inconv":
            dcol = torch.empty(B, H, W, 32, device=dev)
            eng._conv(_cast(eng, g, ("gi",)), eng.in_wd, dcol, taps=1, n_total=32)
            add(ins[0], ops.col2im_c2(dcol, torch.empty(B, H, W, 2, device=dev)))
        else:
            raise AssertionError(kind)
    dx = G[id(ctx["x0"])]
    return ops.lincomb3(torch.empty(B, H * W * 2, device=dev), dx.view(B, -1), rms.contiguous()).view(B, H, W, 2)
