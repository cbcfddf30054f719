"""NCSN++ score-network engine: forward and data-gradient (VJP) as a sequence of buddy_b200 kernels.

Graph = the shipped configuration of the reference (conf/network/ncsnpp.yaml; networks/ncsnpp.py:281-449,
module map in SURVEY.md App. B): 20 BigGAN ResBlocks, one bottleneck attention block, input pyramid (Combine),
output pyramid heads.  Activations are channels-last fp32 in HBM ([B, F=256, frames, C]); every tensor-core
operand is an fp16 copy produced by the GroupNorm+SiLU kernel; all convolutions / NINs / attention products run
on `ops.conv_gemm` (tcgen05).  Only the GroupNorm inputs (+ their statistics) are kept for the backward pass.

This module only orchestrates kernels (pointers, shapes, order); it contains no arithmetic of its own apart from
the one-time weight repacking.  Host-side tensors are torch CUDA tensors used as device-memory handles.
"""
import math
import os

import torch

from . import ops, upfirdn2d
from .ops import MODE_DOWN, MODE_NONE, MODE_UP

INV_SQRT2 = 1.0 / math.sqrt(2.0)
NF = 128
CH_MULT = (1, 2, 2, 2)


def _f16(t):
    return t.to(torch.float16).contiguous()


def _split_k(w, passes):
    """fp32 [..., N, K] -> fp16 B operand [..., N, passes*K].

    passes 1: [w_hi]; 2: [w_hi | w_hi] (pairs with A = [a_hi | a_lo]); 3: [w_hi | w_hi | w_lo] (A chunks wrap, so the
    third block meets a_hi again): a_hi*w_hi + a_lo*w_hi + a_hi*w_lo accumulated in fp32 by one launch."""
    hi = w.to(torch.float16)
    if passes == 1:
        return hi.contiguous()
    if passes == 2:
        return torch.cat([hi, hi], dim=-1).contiguous()
    lo = (w - hi.float()).to(torch.float16)
    return torch.cat([hi, hi, lo], dim=-1).contiguous()


def _pad_rows(w, rows):
    if w.shape[-2] >= rows:
        return w
    pad = torch.zeros(*w.shape[:-2], rows - w.shape[-2], w.shape[-1], dtype=w.dtype, device=w.device)
    return torch.cat([w, pad], dim=-2)


class _RB:
    pass


class Opnd:
    """Tensor-core operand: fp16 tensor (+ optional e4m3 correction pair) and the extra power-of-two scale `gs`
    it was multiplied by (gradient operands are kept near unit rms; consumers undo it in their epilogue)."""
    __slots__ = ("t16", "t8", "gs")

    def __init__(self, t16, t8=None, gs=1.0):
        self.t16, self.t8, self.gs = t16, t8, gs


class WPack:
    """Packed B operand: fp16 [.., N, p*K] (+ e4m3 [.., N, 2K] = [w_hi*2^5 | w_lo*2^14])."""
    __slots__ = ("w16", "w8")

    def __init__(self, w16, w8=None):
        self.w16, self.w8 = w16, w8


def _e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)


class Engine:
    """Device-resident packed weights + forward / vjp drivers."""

    PRECISIONS = {"fp16": 1, "fp16x2": 2, "fp16x3": 3, "fp16c8": 1, "mixed": 1}
    # "mixed": convolutions whose share of the network's FLOPs is largest run as ONE fp16 pass, everything else as
    # fp16c8.  Each conv contributes about the same rounding error to the output whatever its size, so dropping the
    # corrections of the 5 most expensive convs (49 % of the FLOPs; Cin*Cout/4^level >= 32768) costs 3.1e-4 / 4.6e-4
    # (forward / VJP, CPU emulation scripts/precision_study.py; measured on B200 in tests/test_gpu_network.py) of
    # the 1e-3 budget and removes a quarter of the tensor-core work.
    MIXED_X1_THRESHOLD = 32768
    # Data-gradient convolutions with Cin*Cout/4^level >= 16384 also run single-pass ("E"; "" / "F" / "G" = none / only
    # the full-resolution / only the half-resolution ones, for A-B measurements via BUDDY_X1_BWD).  Measured on B200 at
    # full size (tests/test_gpu_network.py, tests/test_gpu_sampler.py): forward error unchanged (3.6e-4), data-gradient
    # 5.4e-4 -> 7.7e-4, sampler trajectories 4.9e-4 -> 5.7e-4 worst case (they are dominated by the forward error), for
    # 1.52 -> 1.34 fp16-pass equivalents per product.
    BWD_X1_POLICY = "E"

    def __init__(self, state_dict, device, precision="mixed", resblock_type="biggan", progressive="output_skip",
                 progressive_input="input_skip", fir=False, fir_kernel=(1, 3, 3, 1)):
        """resblock_type: "biggan" (shipped configuration) or "ddpm": ResnetBlockDDPMpp blocks — the same two-conv block,
           skip through NIN_0 — and, in the place of the resampling ResBlocks, Downsample / Upsample modules with one 3x3
           convolution on the RAW tensor (layerspp.py:93-216; ncsnpp.py:141-144,200-201,262-263).
           progressive / progressive_input other than the shipped output_skip / input_skip run on the general module
           walk of engine_generic.py (same kernels, gradients kept in fp32 between modules).
           fir: the resampling ResBlocks and the pyramids resample with the FIR `fir_kernel` through upfirdn2d
           (up_or_down_sampling.py:195-256) instead of nearest / 2x2-mean; biggan blocks, no pyramid convolutions.
           precision: operand scheme of the conv GEMMs (all accumulate in fp32 on tcgen05):
             "fp16"   one pass, 11-bit significands (what cuDNN's default TF32 path gives the reference on a GPU)
             "fp16x2" activations split hi+lo (fp16), weights single
             "fp16x3" activations and weights split hi+lo (fp16): three fp16 passes in one launch, fp32-class products
             "fp16c8" fp16 products + the two first-order correction terms a_lo*w_hi + a_hi*w_lo as e4m3 MMAs
                      (twice the K per instruction) folded in with tcgen05's scale-input-d: fp32-class products for
                      the cost of two fp16 passes
             "mixed"  fp16c8, except the few largest convolutions which run single-pass (see MIXED_X1_THRESHOLD)"""
        self.device = torch.device(device)
        assert resblock_type in ("biggan", "ddpm"), resblock_type
        self.ddpm = resblock_type == "ddpm"
        # (BUDDY_GENERIC_WALK=1 runs the shipped graph through the general walk as well: tests compare the two)
        self.fir = bool(fir)
        self.generic = ((progressive, progressive_input) != ("output_skip", "input_skip") or self.fir
                        or os.environ.get("BUDDY_GENERIC_WALK", "0") == "1")
        if self.fir:
            assert not self.ddpm and progressive != "residual" and progressive_input != "residual"
            k1 = torch.tensor([float(v) for v in fir_kernel])
            k2 = torch.outer(k1, k1)
            k2 = k2 / k2.sum()                                    # _setup_kernel (up_or_down_sampling.py:181-188)
            p = k2.shape[0] - 2
            self._fir_k = {True: (k2 * 4.0).to(self.device).contiguous(), False: k2.to(self.device).contiguous()}
            self._fir_pad = {True: ((p + 1) // 2 + 1, p // 2), False: ((p + 1) // 2, p // 2)}
        self.resamp = {}        # ddpm: module index -> Downsample / Upsample convolution
        self._k1 = torch.ones(1, 1, device=self.device)         # upfirdn taps: pick / zero-stuff
        self._k22 = torch.ones(2, 2, device=self.device)        # ... and the 2x2 sum (adjoint of nearest x2)
        self.precision = precision
        self.np = self.PRECISIONS[precision]
        self.c8 = precision in ("fp16c8", "mixed")
        self.mixed = precision == "mixed"
        self.x1_convs = set()   # (module index, conv index) whose DATA-GRADIENT launch runs single-pass (mixed mode)
        self.x1_fwd = set()     # ... whose forward launch does
        # BUDDY_FUSE_GNB=1 folds GroupNorm-backward's statistics pass into the producing dgrad convolution's epilogue
        # where the geometry allows (no resample, single input tensor).  Correct (tests) but OFF by default: measured
        # on B200 it removes 19 ms of GroupNorm time per B=32 step and adds 51 ms to the convolutions (the epilogue's
        # exp/rcp work and x reads no longer hide behind the mainloop at the board's power cap): 398 -> 429 ms.
        self.fuse_gnb = os.environ.get("BUDDY_FUSE_GNB", "0") in ("1", "2", "3")
        # "2" / "3": only under data-gradient convolutions with >= 256 output channels (long mainloop per tile: the
        # epilogue's extra work hides behind it) / additionally only at the coarser levels
        self.fuse_gnb_min_n = {"1": 0, "2": 256, "3": 256}.get(os.environ.get("BUDDY_FUSE_GNB", "0"), 0)
        self.fuse_gnb_coarse = os.environ.get("BUDDY_FUSE_GNB", "0") == "3"
        self._graphs = {}
        self._dense_cat = None
        self.graph_max_batch = 8    # larger batches are GPU-bound: plain launches (no pinned graph memory pool)
        self.graph_cache_size = 4   # captured shapes kept (oldest evicted)
        self.split = 2 if self.c8 else (1 if self.np > 1 else 0)
        self.am = 2 if self.split == 1 else 1
        self.gs = {}            # per-call-site power-of-two scales of the gradient operands (fp16c8)
        self._rms = None
        sd = {k: v.detach().to(self.device, torch.float32) for k, v in state_dict.items()}
        self.sd = sd
        f32 = lambda t: t.contiguous()
        # ---- time embedding
        self.fourier_W = f32(sd["all_modules.0.W"])
        self.lin1 = (f32(sd["all_modules.1.weight"]), f32(sd["all_modules.1.bias"]))
        self.lin2 = (f32(sd["all_modules.2.weight"]), f32(sd["all_modules.2.bias"]))
        # ---- input conv 2 -> NF as an im2col GEMM (K index = tap*2 + ci, padded to 64)
        w3 = sd["all_modules.3.weight"].contiguous()  # [NF, 2, 3, 3]; im2col column = tap*2 + ci (18 used)
        self.in_w = self._packv(w3, 1, NF, 64, sn=(NF, 0, 18), sk=(2, 1, 9), k_valid=18)     # [1, NF, p*64]
        self.in_wd = self._packv(w3, 1, 32, NF, sn=(2, 1, 9), sk=(NF, 0, 18), n_valid=18)    # [1, 32, p*NF]
        self.in_b = f32(sd["all_modules.3.bias"])
        self.rb = {}
        self.comb = {}
        self.heads = {}
        if self.generic:
            from . import engine_generic
            engine_generic.build(self, resblock_type, progressive, progressive_input)
        else:
            # ---- walk the module list exactly as the reference builds it
            i = 4
            top = len(CH_MULT) - 1
            for lvl in range(len(CH_MULT)):
                self._pack_rb(i, lvl)
                i += 1
                if lvl != top:
                    if self.ddpm:
                        self._pack_resample(i)
                    else:
                        self._pack_rb(i, lvl + 1)   # `down` block: its convolutions run at the next (coarser) level
                    i += 1
                    self.comb[i] = (f32(sd[f"all_modules.{i}.Conv_0.weight"].reshape(-1, 2)),
                                    f32(sd[f"all_modules.{i}.Conv_0.bias"]))
                    i += 1
            self._pack_rb(i, top)
            self.attn_idx = i + 1
            self._pack_attn(i + 1)
            self._pack_rb(i + 2, top)
            i += 3
            self.up_levels = []
            for lvl in reversed(range(len(CH_MULT))):
                blocks = [i, i + 1]
                self._pack_rb(i, lvl)
                self._pack_rb(i + 1, lvl)
                i += 2
                self._pack_head(i)
                head = i
                i += 2
                upb = None
                if lvl != 0:
                    if self.ddpm:
                        self._pack_resample(i)
                    else:
                        self._pack_rb(i, lvl - 1)   # `up` block: its convolutions run at the next (finer) level
                    upb = i
                    i += 1
                self.up_levels.append((blocks, head, upb))
            assert i == 36
        ow = sd["output_layer.weight"].reshape(2, 2)
        self.out_m = [float(v) for v in ow.reshape(-1).tolist()]
        self.out_mT = [float(v) for v in ow.t().reshape(-1).tolist()]
        self.out_b = [float(v) for v in sd["output_layer.bias"].tolist()]
        self._gsum = None
        if self.c8:
            self._calibrate()

    # ------------------------------------------------------------------ packing
    def _pack(self, w, x1=False):
        """fp32 [T, N, K] -> WPack (x1: fp16 only -> the launch runs a single pass)."""
        w16 = _split_k(w, self.np)
        if not self.c8 or x1:
            return WPack(w16)
        hi = w.to(torch.float16).float()
        w8 = torch.cat([_e4m3(hi * 32.0), _e4m3((w - hi) * 16384.0)], dim=-1).contiguous()
        return WPack(w16, w8)

    def _packv(self, src, T, N, K, x1=False, **kw):
        """One `buddy_pack_weights` launch on the checkpoint tensor itself (element strides, see ops.pack_weights)."""
        w16, w8 = ops.pack_weights(src, T, N, K, passes=self.np, e4m3=self.c8 and not x1, **kw)
        return WPack(w16, w8)

    def _pack3x3(self, w, x1=False, x1_bwd=None):
        """[Co, Ci, 3, 3] -> forward [9, Co, Ci] (tap = 3*ky + kx) and data-gradient [9, Ci, Co] (taps flipped)."""
        w = w.contiguous()
        co, ci = w.shape[:2]
        fwd = self._packv(w, 9, co, ci, x1, st=1, sn=(co, 0, ci * 9), sk=(ci, 0, 9))
        dgr = self._packv(w, 9, ci, co, x1 if x1_bwd is None else x1_bwd, off0=8, st=-1, sn=(ci, 0, 9),
                          sk=(co, 0, ci * 9))
        return fwd, dgr

    def _bwd_extra_x1(self, level, cin, cout):
        pol = os.environ.get("BUDDY_X1_BWD", self.BWD_X1_POLICY)
        m = cin * cout / 4 ** level
        if pol == "E":
            return m >= 16384
        if pol == "F":
            return m >= 16384 and level == 0
        if pol == "G":
            return m >= 16384 and level == 1
        return False

    def _operand(self, B, H, W, C, gs=1.0, need8=True):
        """need8 = False: every consumer of this operand runs a single fp16 pass (no e4m3 correction pair)."""
        t16 = torch.empty(B, H, W, C * self.am, device=self.device, dtype=torch.float16)
        t8 = torch.empty(B, H, W, 2 * C, device=self.device, dtype=torch.uint8) if (self.c8 and need8) else None
        return Opnd(t16, t8, gs)

    def _x1(self, i, k):
        """True if convolution k (0/1; the fused skip conv follows 1) of ResBlock i runs single-pass."""
        return i is not None and (i, k) in self.x1_convs

    def _conv(self, A, Wp, out, *, taps, n_total, A2=None, W2=None, scale=1.0, **kw):
        """Split-precision conv/GEMM launch: undoes the operand's extra scale in the epilogue."""
        c8 = Wp.w8 is not None      # weights packed without corrections -> single fp16 pass
        assert not c8 or (A.t8 is not None and (A2 is None or A2.t8 is not None)), "operand lacks its e4m3 pair"
        return ops.conv_gemm(A.t16, Wp.w16, out, taps=taps, n_total=n_total, passes=self.np,
                             a2=A2.t16 if A2 is not None else None, w2=W2.w16[0] if W2 is not None else None,
                             a8=A.t8 if c8 else None, w8=Wp.w8, a8_2=A2.t8 if (A2 is not None and c8) else None,
                             w8_2=W2.w8[0] if (W2 is not None and c8) else None,
                             scale=scale / A.gs, **kw)

    def _gscale(self, key):
        return self.gs.get(key, 1.0)

    def _record(self, key, op):
        if self._rms is not None:
            self._rms[key] = (op.t16[..., :op.t16.shape[-1] // self.am].float().pow(2).mean().sqrt().item()) / op.gs

    def _calibrate(self):
        """One tiny forward+VJP on seeded noise: measure the rms of every gradient operand and pick power-of-two
        scales that bring them to ~1 (the e4m3 correction operands have a 2^-6..2^8 window).  The ratios are a
        property of the weights, not of the input (measured: identical within 10 % for sigma = 1e-3 .. 0.5)."""
        g = torch.Generator().manual_seed(1234)
        spec = (torch.randn(1, 256, 80, 2, generator=g) * 10.0).to(self.device)
        tc = torch.full((1,), 0.25 * math.log(0.1), device=self.device)
        dout = torch.randn(1, 256, 80, 2, generator=g).to(self.device)
        self._rms = {}
        _, ctx = self._forward_impl(spec, tc, save=True)
        self._vjp_impl(ctx, dout)
        rms, self._rms = self._rms, None
        self.gs = {k: 2.0 ** round(math.log2(1.0 / max(v, 1e-20))) for k, v in rms.items()}
    def _pack_rb(self, i, level):
        sd, p = self.sd, f"all_modules.{i}."
        r = _RB()
        r.g0, r.b0 = sd[p + "GroupNorm_0.weight"].contiguous(), sd[p + "GroupNorm_0.bias"].contiguous()
        r.g1, r.b1 = sd[p + "GroupNorm_1.weight"].contiguous(), sd[p + "GroupNorm_1.bias"].contiguous()
        r.cin, r.cout = sd[p + "Conv_0.weight"].shape[1], sd[p + "Conv_0.weight"].shape[0]
        # relative cost of the two 3x3 convolutions: Cin*Cout / 4^level (level 0 = full resolution)
        r.x1 = [self.mixed and (ci * r.cout) / 4 ** level >= self.MIXED_X1_THRESHOLD for ci in (r.cin, r.cout)]
        # data-gradient launches may run single-pass where the forward conv keeps its corrections: a rounding error in
        # a dgrad conv only reaches the likelihood gradient, a forward one reaches both outputs (emulation:
        # scripts/precision_study.py, policies E-H)
        r.x1b = [r.x1[k] or (self.mixed and self._bwd_extra_x1(level, ci, r.cout)) for k, ci in enumerate((r.cin, r.cout))]
        for k in (0, 1):
            if r.x1b[k]:
                self.x1_convs.add((i, k))
            if r.x1[k]:
                self.x1_fwd.add((i, k))
        r.w0, r.wd0 = self._pack3x3(sd[p + "Conv_0.weight"], r.x1[0], r.x1b[0])
        r.w1, r.wd1 = self._pack3x3(sd[p + "Conv_1.weight"], r.x1[1], r.x1b[1])
        r.bias0 = sd[p + "Conv_0.bias"].contiguous()
        r.dense = (sd[p + "Dense_0.weight"].contiguous(), sd[p + "Dense_0.bias"].contiguous())
        nin = (p + "NIN_0.W") in sd                                 # ddpm blocks: skip = NIN_0, W stored (in, out)
        r.has_skip_conv = nin or (p + "Conv_2.weight") in sd
        if r.has_skip_conv:
            w2 = sd[p + "NIN_0.W"].t().contiguous() if nin else sd[p + "Conv_2.weight"].contiguous()   # [Cout, Cin(,1,1)]
            r.w2 = self._packv(w2, 1, r.cout, r.cin, r.x1[1], sn=(r.cout, 0, r.cin), sk=(r.cin, 0, 1))   # fused into conv 1
            r.wd2 = self._packv(w2, 1, r.cin, r.cout, r.x1b[1], sn=(r.cin, 0, 1), sk=(r.cout, 0, r.cin))
            r.bias1 = (sd[p + "Conv_1.bias"] + sd[p + ("NIN_0.b" if nin else "Conv_2.bias")]).contiguous()
        else:
            r.bias1 = sd[p + "Conv_1.bias"].contiguous()
        self.rb[i] = r

    def _pack_resample(self, i):
        """ddpm Downsample / Upsample (layerspp.py:93-160, with_conv, fir = False): one 3x3 convolution with bias."""
        sd, p = self.sd, f"all_modules.{i}."
        m = _RB()
        w = sd[p + "Conv_0.weight"]
        m.cout, m.cin = w.shape[:2]
        m.w, m.wd = self._pack3x3(w)
        m.bias = sd[p + "Conv_0.bias"].contiguous()
        self.resamp[i] = m

    def _pack_attn(self, i):
        sd, p = self.sd, f"all_modules.{i}."
        a = _RB()
        a.g, a.b = sd[p + "GroupNorm_0.weight"].contiguous(), sd[p + "GroupNorm_0.bias"].contiguous()
        ws = [sd[p + f"NIN_{n}.W"] for n in range(4)]  # (in, out)
        a.c = ws[0].shape[0]
        a.wqkv = _f16(torch.cat([w.t() for w in ws[:3]], dim=0))[None]       # [1, 3C, C]  rows = outputs
        a.bqkv = torch.cat([sd[p + f"NIN_{n}.b"] for n in range(3)]).contiguous()
        a.wqkv_d = _f16(torch.cat(ws[:3], dim=1))[None]                      # [1, C, 3C]  dgrad: rows = inputs
        a.w3 = _f16(ws[3].t())[None]                                         # [1, C, C]
        a.w3_d = _f16(ws[3])[None]
        a.b3 = sd[p + "NIN_3.b"].contiguous()
        self.attn = a

    def _pack_head(self, i):
        sd = self.sd
        h = _RB()
        h.g, h.b = sd[f"all_modules.{i}.weight"].contiguous(), sd[f"all_modules.{i}.bias"].contiguous()
        w = sd[f"all_modules.{i + 1}.weight"].contiguous()  # [2, C, 3, 3]
        c = w.shape[1]
        # forward as ONE 1x1 GEMM to 18 (-> 32) per-tap partial outputs + a 9-tap gather (col2im_c2):
        #   colh[q][t*2+co] = sum_c a[q][c] * W[co][c][8-t],  out[p][co] = sum_t colh[p - off(t)][t*2+co]
        # (a 3x3 tensor-core conv with N = 2 padded to 16 re-reads the operand patches for nothing: 1.5 ms -> 0.4 ms
        # at full resolution, B = 16).  The bias rides on the centre tap, which is always inside the image.
        # row = t*2 + co (18 of 32 used) <- W[co][:, 8 - t]
        h.w = self._packv(w, 1, 32, c, off0=8, sn=(2, -1, c * 9), sk=(c, 0, 9), n_valid=18)    # [1, 32, p*C]
        bias = torch.zeros(32, device=self.device)
        bias[8:10] = sd[f"all_modules.{i + 1}.bias"]
        h.bias = bias
        # dgrad as an im2col GEMM: dcol[p][tap'*2+co] = dP[p + tap' offset][co];  wd[c][tap'*2+co] = W[co][c][2-ky'][2-kx']
        h.wd = self._packv(w, 1, c, 64, off0=8, sn=(c, 0, 9), sk=(2, -1, c * 9), k_valid=18)   # [1, C, p*64]
        h.c = c
        self.heads[i] = h

    # ------------------------------------------------------------------ helpers
    def _scratch_gsum(self, B):
        # fresh per call (stream-ordered allocator): micro-batches on different streams must not share it
        return torch.empty(B, 32, 2, device=self.device, dtype=torch.float64)

    def _zeros_stats(self, B, C):
        return torch.zeros(B, C // 4, 2, device=self.device, dtype=torch.float64)

    def time_bias(self, time_cond):
        """Per-ResBlock additive bias Dense_0(SiLU(temb)) [B, Cout] for every block (ncsnpp.py:299-318)."""
        B = time_cond.shape[0]
        dev = self.device
        emb = torch.empty(B, 2 * NF, device=dev)
        ops.fourier_features(time_cond.contiguous(), self.fourier_W, emb)
        t1 = torch.empty(B, 4 * NF, device=dev)
        ops.dense(emb, self.lin1[0], self.lin1[1], t1)
        t2 = torch.empty(B, 4 * NF, device=dev)
        ops.dense(t1, self.lin2[0], self.lin2[1], t2, act_in=True)
        if self._dense_cat is None:         # all Dense_0 layers as one [sum Cout, 4 nf] matrix: one launch instead of 21
            ws, bs, seg, r0 = [], [], [], 0
            for i, r in self.rb.items():
                ws.append(r.dense[0])
                bs.append(r.dense[1])
                seg += [[r0, r.cout]] * r.cout
                r0 += r.cout
            self._dense_cat = (torch.cat(ws).contiguous(), torch.cat(bs).contiguous(),
                               torch.tensor(seg, dtype=torch.int32, device=dev))
        W, b, seg = self._dense_cat
        flat = ops.dense_seg(t2, W, b, seg, torch.empty(B * W.shape[0], device=dev), act_in=True)
        out, r0 = {}, 0
        for i, r in self.rb.items():
            out[i] = flat[B * r0:B * (r0 + r.cout)].view(B, r.cout)
            r0 += r.cout
        return out

    # ------------------------------------------------------------------ FIR resampling (fir: True)
    def _fir_fwd(self, x4, up):
        """upsample_2d / downsample_2d by 2 (up_or_down_sampling.py:195-256) of fp32 [B, H, W, C]."""
        p0, p1 = self._fir_pad[up]
        return upfirdn2d._launch(x4, self._fir_k[up], (2, 2) if up else (1, 1), (1, 1) if up else (2, 2),
                                 (p0, p1, p0, p1))

    def _fir_bwd(self, g4, up, in_h, in_w):
        """Adjoint of `_fir_fwd` (op/upfirdn2d.py:104-107,124-139: flipped kernel, up and down exchanged, complementary
        padding): gradient w.r.t. an [B, in_h, in_w, C] input."""
        k = self._fir_k[up]
        kk = k.shape[0]
        p0, _ = self._fir_pad[up]
        u, d = (2, 1) if up else (1, 2)
        oh, ow = g4.shape[1], g4.shape[2]
        g_pad = (kk - p0 - 1, in_w * u - ow * d + p0 - u + 1, kk - p0 - 1, in_h * u - oh * d + p0 - u + 1)
        return upfirdn2d._launch(g4.contiguous(), torch.flip(k, [0, 1]).contiguous(), (d, d), (u, u), g_pad)

    # ------------------------------------------------------------------ ResBlock
    def _rb_fwd(self, i, xa, sa, xb, sb, tb, mode, save):
        r = self.rb[i]
        B, H, W, Ca = xa.shape
        C = Ca + (xb.shape[3] if xb is not None else 0)
        assert C == r.cin, (i, C, r.cin)
        Ho, Wo = (2 * H, 2 * W) if mode == MODE_UP else ((H // 2, W // 2) if mode == MODE_DOWN else (H, W))
        dev = self.device
        a0 = self._operand(B, Ho, Wo, C, need8=not r.x1[0])
        raw = self._operand(B, Ho, Wo, C, need8=not r.x1[1]) if r.has_skip_conv else None
        if self.fir and mode != MODE_NONE:
            # layerspp.py:252-259 with fir: h = resample_fir(act(GroupNorm_0(x))), x = resample_fir(x)
            assert xb is None and raw is not None
            act = self._fir_fwd(ops.gn_act32(xa, sa, r.g0, r.b0, torch.empty_like(xa)), mode == MODE_UP)
            ops.cast_operand(act, a0.t16, a0.t8, split=self.split)
            del act
            ops.cast_operand(self._fir_fwd(xa, mode == MODE_UP), raw.t16, raw.t8, split=self.split)
        else:
            ops.gn_apply(xa, sa, r.g0, r.b0, a0.t16, xb=xb, sb=sb, silu=True, mode=mode,
                         out_raw=raw.t16 if raw else None, split=self.split, out8=a0.t8,
                         out_raw8=raw.t8 if raw else None)
        h1 = torch.empty(B, Ho, Wo, r.cout, device=dev)
        s1 = self._zeros_stats(B, r.cout)
        self._conv(a0, r.w0, h1, taps=9, n_total=r.cout, bias=r.bias0, bias_b=tb[i], stats=s1)
        del a0
        a1 = self._operand(B, Ho, Wo, r.cout, need8=not r.x1[1])
        ops.gn_apply(h1, s1, r.g1, r.b1, a1.t16, silu=True, split=self.split, out8=a1.t8)
        out = torch.empty(B, Ho, Wo, r.cout, device=dev)
        so = self._zeros_stats(B, r.cout)
        if r.has_skip_conv:
            self._conv(a1, r.w1, out, taps=9, n_total=r.cout, A2=raw, W2=r.w2, bias=r.bias1, scale=INV_SQRT2, stats=so)
        else:
            assert mode == MODE_NONE and xb is None
            self._conv(a1, r.w1, out, taps=9, n_total=r.cout, bias=r.bias1, resid=xa, scale=INV_SQRT2, stats=so)
        if save is not None:
            save[i] = (xa, sa, xb, sb, h1, s1, mode)
        return out, so

    def _rb_bwd(self, i, saved, g16, dout32, extra_a=None, want_a32=True, want_a16=True, a16_scale=INV_SQRT2,
                consumer=None):
        """g16 = fp16(dout/sqrt2); returns (dxa32, g16a, dxb32).  `consumer` = ResBlock whose conv-1 / skip dgrad
        reads the returned g16a (decides whether it needs the e4m3 correction pair)."""
        r = self.rb[i]
        xa, sa, xb, sb, h1, s1, mode = saved[i]
        B, Ho, Wo, _ = h1.shape
        dev = self.device
        gsum = self._scratch_gsum(B)
        da1 = torch.empty(B, Ho, Wo, r.cout, device=dev)
        ok_lvl = not self.fuse_gnb_coarse or Ho < 256
        fuse1 = self.fuse_gnb and r.cout >= self.fuse_gnb_min_n and ok_lvl   # GroupNorm_1 sits directly under conv 1
        gsum1 = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64) if fuse1 else gsum
        self._conv(g16, r.wd1, da1, taps=9, n_total=r.cout,
                   gnb=(h1, s1, r.g1, r.b1, gsum1, 32, 1e-6, 1) if fuse1 else None)
        dh1 = self._operand(B, Ho, Wo, r.cout, self._gscale(("h1", i)), need8=not r.x1b[0])
        ops.gn_bwd(h1, s1, r.g1, r.b1, da1, gsum1, silu=True, g16a=dh1.t16, g16_scale=dh1.gs, split=self.split,
                   g8a=dh1.t8, pass0_done=fuse1)
        self._record(("h1", i), dh1)
        del da1
        da0 = torch.empty(B, Ho, Wo, r.cin, device=dev)
        fuse0 = (self.fuse_gnb and mode == MODE_NONE and xb is None and r.cin >= self.fuse_gnb_min_n
                 and ok_lvl)                                              # GroupNorm_0 of a plain block
        gsum0 = torch.zeros(B, 32, 2, device=dev, dtype=torch.float64) if fuse0 else gsum
        self._conv(dh1, r.wd0, da0, taps=9, n_total=r.cin,
                   gnb=(xa, sa, r.g0, r.b0, gsum0, 32, 1e-6, 1) if fuse0 else None)
        del dh1
        if r.has_skip_conv:
            dsk = torch.empty(B, Ho, Wo, r.cin, device=dev)
            self._conv(g16, r.wd2, dsk, taps=1, n_total=r.cin)
            skip_scale = 1.0
        else:
            dsk, skip_scale = dout32, INV_SQRT2
        if self.fir and mode != MODE_NONE:
            # pull both gradients back through the FIR resampler; the GroupNorm backward then sees no resampling
            da0 = self._fir_bwd(da0, mode == MODE_UP, xa.shape[1], xa.shape[2])
            dsk = self._fir_bwd(dsk, mode == MODE_UP, xa.shape[1], xa.shape[2])
            mode = MODE_NONE
        Ca = xa.shape[3]
        dxa = torch.empty_like(xa) if want_a32 else None
        g16a = (self._operand(*xa.shape[:3], Ca, self._gscale(("x", i)), need8=not self._x1(consumer, 1))
                if want_a16 else None)
        dxb = torch.empty_like(xb) if xb is not None else None
        ops.gn_bwd(xa, sa, r.g0, r.b0, da0, gsum0, xb=xb, sb=sb, silu=True, mode=mode, dskip=dsk, skip_scale=skip_scale,
                   extra_a=extra_a, dxa=dxa, dxb=dxb, g16a=g16a.t16 if g16a else None,
                   g16_scale=a16_scale * (g16a.gs if g16a else 1.0), split=self.split, g8a=g16a.t8 if g16a else None,
                   pass0_done=fuse0)
        if g16a:
            self._record(("x", i), g16a)
        return dxa, g16a, dxb

    # ------------------------------------------------------------------ ddpm Downsample / Upsample
    def _add32(self, a, b):
        B = a.shape[0]
        one = torch.ones(B, device=self.device)
        return ops.lincomb3(torch.empty(B, a[0].numel(), device=self.device), a.view(B, -1), one, b.view(B, -1),
                            one).view(a.shape)

    def _down_fwd(self, i, x):
        """Downsample (layerspp.py:147-154): zero-pad right/bottom by one, 3x3 convolution with stride 2.  Output pixel
        (r, c) is pixel (2r+1, 2c+1) of the stride-1 'same' convolution, so the tensor-core kernel runs unchanged at
        the input resolution and upfirdn2d (1 tap, down 2, origin -1) keeps the odd pixels."""
        m = self.resamp[i]
        B, H, W, C = x.shape
        a = self._operand(B, H, W, C)
        ops.cast_operand(x, a.t16, a.t8, split=self.split)
        full = torch.empty(B, H, W, m.cout, device=self.device)
        self._conv(a, m.w, full, taps=9, n_total=m.cout, bias=m.bias)
        return upfirdn2d._launch(full, self._k1, (1, 1), (2, 2), (-1, 0, -1, 0))

    def _down_bwd(self, i, d32, extra, consumer, scale=1.0, want_g=True):
        """d32 fp32 = gradient w.r.t. the module's output (times `scale`); returns the gradient w.r.t. its input
        (+ `extra`) in fp32 and, want_g, as the fp16(/sqrt2) operand of the producing ResBlock."""
        m = self.resamp[i]
        B, h, w, C = d32.shape
        zs = upfirdn2d._launch(d32, self._k1, (2, 2), (1, 1), (1, -1, 1, -1))     # zero-stuffed: (2r+1, 2c+1) <- (r, c)
        op = self._operand(B, 2 * h, 2 * w, C, self._gscale(("dn", i)))
        ops.cast_operand(zs, op.t16, op.t8, scale=scale * op.gs, split=self.split)
        self._record(("dn", i), op)
        del zs
        dx = torch.empty(B, 2 * h, 2 * w, m.cin, device=self.device)
        self._conv(op, m.wd, dx, taps=9, n_total=m.cin)
        if extra is not None:
            dx = self._add32(dx, extra)
        if not want_g:
            return dx, None
        g = self._operand(B, 2 * h, 2 * w, C, self._gscale(("dnx", i)), need8=not self._x1(consumer, 1))
        ops.cast_operand(dx, g.t16, g.t8, scale=INV_SQRT2 * g.gs, split=self.split)
        self._record(("dnx", i), g)
        return dx, g

    def _up_fwd(self, i, x):
        """Upsample (layerspp.py:113-118): nearest x2 (folded into the operand cast), then the 3x3 convolution."""
        m = self.resamp[i]
        B, H, W, C = x.shape
        a = self._operand(B, 2 * H, 2 * W, C)
        ops.cast_operand(x, a.t16, a.t8, upsample=True, split=self.split)
        out = torch.empty(B, 2 * H, 2 * W, m.cout, device=self.device)
        so = self._zeros_stats(B, m.cout)
        self._conv(a, m.w, out, taps=9, n_total=m.cout, bias=m.bias, stats=so)
        return out, so

    def _up_bwd(self, i, g16):
        """g16 = operand of the gradient w.r.t. the module's output (scale 1) -> fp32 gradient w.r.t. its input."""
        m = self.resamp[i]
        B, H2, W2, _ = g16.t16.shape
        da = torch.empty(B, H2, W2, m.cin, device=self.device)
        self._conv(g16, m.wd, da, taps=9, n_total=m.cin)
        return upfirdn2d._launch(da, self._k22, (1, 1), (2, 2), (0, 0, 0, 0))         # 2x2 sums

    # ------------------------------------------------------------------ attention
    # Attention over N = H*W tokens (AttnBlockpp, layerspp.py:75-91).  Up to ATTN_DENSE_BYTES of logits + probabilities
    # (4 s utterances: N = 2112, 27 MB per utterance, 0.86 GB for a micro-batch of 32) the N x N matrices are materialised once and P is kept for the
    # backward pass.  Beyond (30 s: N = 15 040, 1.36 GB per utterance) the block runs over query blocks of ATTN_QBLOCK
    # rows with exact row-wise softmax — at most ATTN_QBLOCK x N logits exist at a time, nothing N x N is stored, and
    # the backward pass recomputes each block's probabilities from q, k (flash-attention style recomputation).
    ATTN_DENSE_BYTES = 1 << 30
    ATTN_QBLOCK = 2048

    def _attn_blocked(self, B, N):
        return B * N * N * 6 > self.ATTN_DENSE_BYTES

    def _attn_probs(self, q, k, r0, nq, C):
        """softmax(q[r0:r0+nq] k^T / sqrt(C)) for one query block: fp16 [B, nq, N] (logits live in fp32 scratch)."""
        B, N = q.shape[0], q.shape[2]
        S = torch.empty(B, 1, nq, N, device=self.device)
        ops.conv_gemm(q[:, :, r0:r0 + nq], k[:, 0], S, taps=1, n_total=N, b_batched=True, scale=float(C) ** -0.5)
        P = torch.empty(B, nq, N, device=self.device, dtype=torch.float16)
        return ops.softmax_fwd(S, P)

    def _attn_fwd_blocked(self, q, k, v, C):
        B, N = q.shape[0], q.shape[2]
        dev = self.device
        vT = torch.empty(B, C, N, device=dev, dtype=torch.float16)
        ops.transpose_h(v[:, 0], vT)
        o = torch.empty(B, 1, N, C, device=dev, dtype=torch.float16)
        for r0 in range(0, N, self.ATTN_QBLOCK):
            nq = min(self.ATTN_QBLOCK, N - r0)
            P = self._attn_probs(q, k, r0, nq, C)
            ob = torch.empty(B, 1, nq, C, device=dev, dtype=torch.float16)
            ops.conv_gemm(P.view(B, 1, nq, N), vT, ob, taps=1, n_total=C, b_batched=True)
            o[:, :, r0:r0 + nq].copy_(ob)
        return o

    def _attn_fwd(self, x, sx, save):
        a = self.attn
        B, H, W, C = x.shape
        N = H * W
        dev = self.device
        hn = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
        ops.gn_apply(x, sx, a.g, a.b, hn, silu=False)
        qkv = torch.empty(B, 1, N, 3 * C, device=dev, dtype=torch.float16)
        ops.conv_gemm(hn.view(B, 1, N, C), a.wqkv, qkv, taps=1, n_total=3 * C, bias=a.bqkv)
        del hn
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        if self._attn_blocked(B, N):
            o = self._attn_fwd_blocked(q, k, v, C)
            out = torch.empty(B, H, W, C, device=dev)
            so = self._zeros_stats(B, C)
            ops.conv_gemm(o.view(B, H, W, C), a.w3, out, taps=1, n_total=C, bias=a.b3, resid=x, scale=INV_SQRT2,
                          stats=so)
            if save is not None:
                save["attn"] = (x, sx, qkv, None)
            return out, so
        S = torch.empty(B, 1, N, N, device=dev)
        ops.conv_gemm(q, k[:, 0], S, taps=1, n_total=N, b_batched=True, scale=float(C) ** -0.5)
        P = torch.empty(B, N, N, device=dev, dtype=torch.float16)
        ops.softmax_fwd(S, P)
        del S
        vT = torch.empty(B, C, N, device=dev, dtype=torch.float16)
        ops.transpose_h(v[:, 0], vT)
        o = torch.empty(B, 1, N, C, device=dev, dtype=torch.float16)
        ops.conv_gemm(P.view(B, 1, N, N), vT, o, taps=1, n_total=C, b_batched=True)
        out = torch.empty(B, H, W, C, device=dev)
        so = self._zeros_stats(B, C)
        ops.conv_gemm(o.view(B, H, W, C), a.w3, out, taps=1, n_total=C, bias=a.b3, resid=x, scale=INV_SQRT2, stats=so)
        if save is not None:
            save["attn"] = (x, sx, qkv, P)
        return out, so

    def _attn_bwd(self, saved, g16, dout32, consumer=None):
        a = self.attn
        x, sx, qkv, P = saved["attn"]
        B, H, W, C = x.shape
        N = H * W
        dev = self.device
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        do = torch.empty(B, 1, N, C, device=dev, dtype=torch.float16)
        ops.conv_gemm(g16.t16.view(B, 1, N, C * self.am)[..., :C], a.w3_d, do, taps=1, n_total=C, scale=1.0 / g16.gs)
        if P is None:
            dqkv = self._attn_bwd_blocked(q, k, v, do, C)
            dhn = torch.empty(B, H, W, C, device=dev)
            ops.conv_gemm(dqkv.view(B, H, W, 3 * C), a.wqkv_d, dhn, taps=1, n_total=C)
            dx = torch.empty_like(x)
            g16x = self._operand(B, H, W, C, self._gscale(("attn",)), need8=not self._x1(consumer, 1))
            ops.gn_bwd(x, sx, a.g, a.b, dhn, self._scratch_gsum(B), silu=False, dskip=dout32, skip_scale=INV_SQRT2,
                       dxa=dx, g16a=g16x.t16, g16_scale=INV_SQRT2 * g16x.gs, split=self.split, g8a=g16x.t8)
            return dx, g16x
        doT = torch.empty(B, C, N, device=dev, dtype=torch.float16)
        ops.transpose_h(do[:, 0], doT)
        PT = torch.empty(B, N, N, device=dev, dtype=torch.float16)
        ops.transpose_h(P, PT)
        dqkv = torch.empty(B, 1, N, 3 * C, device=dev, dtype=torch.float16)
        ops.conv_gemm(PT.view(B, 1, N, N), doT, dqkv, taps=1, n_total=C, b_batched=True, col_off=2 * C, ldc=3 * C)
        del PT, doT
        dP = torch.empty(B, 1, N, N, device=dev)
        ops.conv_gemm(do, v[:, 0], dP, taps=1, n_total=N, b_batched=True)
        dS = torch.empty(B, N, N, device=dev, dtype=torch.float16)
        ops.softmax_bwd(P, dP, float(C) ** -0.5, dS)
        del dP
        kT = torch.empty(B, C, N, device=dev, dtype=torch.float16)
        ops.transpose_h(k[:, 0], kT)
        ops.conv_gemm(dS.view(B, 1, N, N), kT, dqkv, taps=1, n_total=C, b_batched=True, col_off=0, ldc=3 * C)
        dST = torch.empty(B, N, N, device=dev, dtype=torch.float16)
        ops.transpose_h(dS, dST)
        qT = kT
        ops.transpose_h(q[:, 0], qT)
        ops.conv_gemm(dST.view(B, 1, N, N), qT, dqkv, taps=1, n_total=C, b_batched=True, col_off=C, ldc=3 * C)
        del dS, dST
        dhn = torch.empty(B, H, W, C, device=dev)
        ops.conv_gemm(dqkv.view(B, H, W, 3 * C), a.wqkv_d, dhn, taps=1, n_total=C)
        dx = torch.empty_like(x)
        g16x = self._operand(B, H, W, C, self._gscale(("attn",)), need8=not self._x1(consumer, 1))
        ops.gn_bwd(x, sx, a.g, a.b, dhn, self._scratch_gsum(B), silu=False, dskip=dout32, skip_scale=INV_SQRT2, dxa=dx,
                   g16a=g16x.t16, g16_scale=INV_SQRT2 * g16x.gs, split=self.split, g8a=g16x.t8)
        self._record(("attn",), g16x)
        return dx, g16x

    def _attn_bwd_blocked(self, q, k, v, do, C):
        """d(q, k, v) for the query-blocked attention: per block, recompute P, then
        dP = dO V^T, dS = P * (dP - rowsum(P * dP)) / sqrt(C), dQ = dS K, dK += dS^T Q, dV += P^T dO
        (dK, dV accumulate over the blocks in fp32)."""
        B, N = q.shape[0], q.shape[2]
        dev = self.device
        h16 = dict(device=dev, dtype=torch.float16)
        kT = torch.empty(B, C, N, **h16)
        ops.transpose_h(k[:, 0], kT)
        dk32 = torch.zeros(B, 1, N, C, device=dev)
        dv32 = torch.zeros(B, 1, N, C, device=dev)
        dqkv = torch.empty(B, 1, N, 3 * C, **h16)
        for r0 in range(0, N, self.ATTN_QBLOCK):
            nq = min(self.ATTN_QBLOCK, N - r0)
            P = self._attn_probs(q, k, r0, nq, C)
            dP = torch.empty(B, 1, nq, N, device=dev)
            ops.conv_gemm(do[:, :, r0:r0 + nq], v[:, 0], dP, taps=1, n_total=N, b_batched=True)
            dS = torch.empty(B, nq, N, **h16)
            ops.softmax_bwd(P, dP, float(C) ** -0.5, dS)
            del dP
            dqb = torch.empty(B, 1, nq, C, **h16)
            ops.conv_gemm(dS.view(B, 1, nq, N), kT, dqb, taps=1, n_total=C, b_batched=True)
            dqkv[:, :, r0:r0 + nq, :C].copy_(dqb)
            T = torch.empty(B, N, nq, **h16)               # dS^T, then P^T
            bT = torch.empty(B, C, nq, **h16)              # q_block^T, then dO_block^T
            ops.transpose_h(dS, T)
            ops.transpose_h(q[:, 0, r0:r0 + nq], bT)
            ops.conv_gemm(T.view(B, 1, N, nq), bT, dk32, taps=1, n_total=C, b_batched=True, resid=dk32)
            ops.transpose_h(P, T)
            ops.transpose_h(do[:, 0, r0:r0 + nq], bT)
            ops.conv_gemm(T.view(B, 1, N, nq), bT, dv32, taps=1, n_total=C, b_batched=True, resid=dv32)
        dqkv[..., C:2 * C].copy_(dk32)
        dqkv[..., 2 * C:].copy_(dv32)
        return dqkv

    # ------------------------------------------------------------------ pyramid heads
    def _head_fwd(self, i, h, sh, save):
        hd = self.heads[i]
        B, H, W, C = h.shape
        a = self._operand(B, H, W, C)
        ops.gn_apply(h, sh, hd.g, hd.b, a.t16, silu=True, split=self.split, out8=a.t8)
        colh = torch.empty(B, H, W, 32, device=self.device)
        self._conv(a, hd.w, colh, taps=1, n_total=32, bias=hd.bias)
        return ops.col2im_c2(colh, torch.empty(B, H, W, 2, device=self.device))

    def _head_bwd(self, i, h, sh, dP, extra, want32, consumer=None):
        """dP fp32 [B,H,W,2] -> gradient w.r.t. h (plus `extra`), fp32 (optional) and fp16/sqrt2."""
        hd = self.heads[i]
        B, H, W, C = h.shape
        dev = self.device
        col = self._operand(B, H, W, 64, self._gscale(("dP", i)))
        ops.im2col_c2(dP, col.t16, split=self.split, col8=col.t8, in_scale=col.gs)
        self._record(("dP", i), col)
        da = torch.empty(B, H, W, C, device=dev)
        self._conv(col, hd.wd, da, taps=1, n_total=C)
        dx = torch.empty_like(h) if want32 else None
        g16 = self._operand(B, H, W, C, self._gscale(("hx", i)), need8=not self._x1(consumer, 1))
        ops.gn_bwd(h, sh, hd.g, hd.b, da, self._scratch_gsum(B), silu=True, extra_a=extra, dxa=dx, g16a=g16.t16,
                   g16_scale=INV_SQRT2 * g16.gs, split=self.split, g8a=g16.t8)
        self._record(("hx", i), g16)
        return dx, g16

    # ------------------------------------------------------------------ forward
    # ------------------------------------------------------------------ CUDA graphs (small batches are launch-bound)
    def _graph_entry(self, spec, save):
        """Capture forward (and, with save, the VJP) of this (batch, frames) shape once: two CUDA graphs sharing one
        memory pool, static input / output buffers.  ~320 kernel launches per evaluation cost ~33 us of host time
        each; a single 4 s utterance (the reference's own B = 1 usage) is host-bound without this."""
        B, H, W, _ = spec.shape
        key = (B, W, bool(save), torch.cuda.current_stream().cuda_stream)
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        dev = self.device
        ent = {"spec": torch.empty_like(spec), "tc": torch.empty(B, device=dev)}
        ent["spec"].copy_(spec)
        ent["tc"].fill_(-0.5)
        # eager warm-up on the capture inputs (lazy one-off initialisation inside the library must not be captured)
        out, ctx = self._forward_impl(ent["spec"], ent["tc"], save)
        if save:
            self._vjp_impl(ctx, torch.ones_like(out))
        del out, ctx
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        ent["fwd"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ent["fwd"], pool=pool):
            ent["out"], ent["ctx"] = self._forward_impl(ent["spec"], ent["tc"], save)
        if save:
            ent["dout"] = torch.empty_like(ent["out"])
            ent["bwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ent["bwd"], pool=pool):
                ent["dx"] = self._vjp_impl(ent["ctx"], ent["dout"])
        while len(self._graphs) >= self.graph_cache_size:        # each entry pins its activations: keep a few shapes
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = ent
        return ent

    def forward(self, spec, time_cond, save=True, graph=False):
        """spec fp32 [B, 256, Tp, 2] (re, im channels-last), time_cond fp32 [B] -> (out [B,256,Tp,2], ctx).

        graph=True (used by the samplers for batches <= `graph_max_batch`): replay a captured CUDA graph.  The returned
        tensors are the graph's static buffers — valid until the next graphed call of the same shape."""
        # (graphs pin their activations in a private pool: only for launch-bound sizes, up to graph_max_batch 4 s
        # utterances' worth of frames)
        if graph and spec.shape[0] * spec.shape[2] <= self.graph_max_batch * 528 and ops._timer is None:
            ent = self._graph_entry(spec, save)
            ent["spec"].copy_(spec)
            ent["tc"].copy_(time_cond)
            ent["fwd"].replay()
            return ent["out"], ({"_graph": ent} if save else None)
        return self._forward_impl(spec, time_cond, save)

    def vjp(self, ctx, dout):
        """dout fp32 [B,256,Tp,2] (gradient w.r.t. forward's output) -> gradient w.r.t. `spec`."""
        ent = ctx.get("_graph") if isinstance(ctx, dict) else None
        if ent is not None:
            ent["dout"].copy_(dout)
            ent["bwd"].replay()
            return ent["dx"]
        return self._vjp_impl(ctx, dout)

    def _forward_impl(self, spec, time_cond, save=True):
        """spec fp32 [B, 256, Tp, 2] (re, im channels-last), time_cond fp32 [B] -> (out [B,256,Tp,2], ctx)."""
        assert spec.dtype == torch.float32 and spec.is_contiguous() and spec.shape[1] == 256 and spec.shape[3] == 2
        B, H, W, _ = spec.shape
        assert W % 16 == 0, "frame count must be a multiple of 16 (NCSNppTime.stft pads to it)"
        if self.generic:
            from . import engine_generic
            return engine_generic.forward(self, spec, time_cond, save)
        dev = self.device
        ctx = {} if save else None
        tb = self.time_bias(time_cond)
        # input pyramid (pyramid_downsample = 2x2 mean, ncsnpp.py:355-357)
        pyr = [spec]
        for _ in range(3):
            p = pyr[-1]
            pyr.append(ops.resample_c2(p, 0, torch.empty(B, p.shape[1] // 2, p.shape[2] // 2, 2, device=dev)))
        col = self._operand(B, H, W, 64)
        ops.im2col_c2(spec, col.t16, split=self.split, col8=col.t8)
        h = torch.empty(B, H, W, NF, device=dev)
        sh = self._zeros_stats(B, NF)
        self._conv(col, self.in_w, h, taps=1, n_total=NF, bias=self.in_b, stats=sh)
        del col
        hs = [(h, sh)]
        i = 4
        for lvl in range(4):
            h, sh = self._rb_fwd(i, hs[-1][0], hs[-1][1], None, None, tb, MODE_NONE, ctx)
            i += 1
            hs.append((h, sh))
            if lvl != 3:
                if self.ddpm:
                    h = self._down_fwd(i, h)
                else:
                    h, sh = self._rb_fwd(i, h, sh, None, None, tb, MODE_DOWN, ctx)
                i += 1
                w, b = self.comb[i]
                hc = ops.combine_fwd(h, pyr[lvl + 1], w, b, torch.empty_like(h))
                sc = ops.gn_stats(hc)
                i += 1
                hs.append((hc, sc))
        h, sh = self._rb_fwd(i, hs[-1][0], hs[-1][1], None, None, tb, MODE_NONE, ctx)
        h, sh = self._attn_fwd(h, sh, ctx)
        h, sh = self._rb_fwd(i + 2, h, sh, None, None, tb, MODE_NONE, ctx)
        pyramid = None
        head_in = {}
        for (blocks, head, upb) in self.up_levels:
            for bi in blocks:
                xb, sb = hs.pop()
                h, sh = self._rb_fwd(bi, h, sh, xb, sb, tb, MODE_NONE, ctx)
            ph = self._head_fwd(head, h, sh, ctx)
            head_in[head] = (h, sh)
            if pyramid is None:
                pyramid = ph
            else:
                pyramid = ops.resample_c2(pyramid, 1, torch.empty_like(ph), add=ph)
            if upb is not None:
                if self.ddpm:
                    h, sh = self._up_fwd(upb, h)
                else:
                    h, sh = self._rb_fwd(upb, h, sh, None, None, tb, MODE_UP, ctx)
        assert not hs
        out = ops.affine_c2(pyramid, self.out_m, self.out_b, torch.empty_like(pyramid))
        if save:
            ctx["head_in"] = head_in
            ctx["shape"] = (B, H, W)
        return out, ctx

    # ------------------------------------------------------------------ data-gradient
    def _vjp_impl(self, ctx, dout):
        if self.generic:
            from . import engine_generic
            return engine_generic.vjp(self, ctx, dout)
        B, H, W = ctx["shape"]
        dev = self.device
        assert dout.shape == (B, H, W, 2) and dout.is_contiguous()
        # the VJP is linear: bring every utterance's cotangent to unit rms so the fp16 / e4m3 dgrad operands stay
        # in range whatever the magnitude of the loss, and undo it on the result
        rs = ops.row_stats(dout.view(B, -1))
        rms = torch.sqrt(rs[:, 1] / (H * W * 2)).float().clamp_min(1e-30)
        dout = ops.lincomb3(torch.empty(B, H * W * 2, device=dev), dout.view(B, -1), (1.0 / rms).contiguous()).view(
            B, H, W, 2)
        zero_b = [0.0, 0.0]
        dP = [ops.affine_c2(dout, self.out_mT, zero_b, torch.empty_like(dout))]   # level 0 (full res)
        for _ in range(3):
            p = dP[-1]
            dP.append(ops.resample_c2(p, 3, torch.empty(B, p.shape[1] // 2, p.shape[2] // 2, 2, device=dev)))
        partial_hs = {}     # index into hs (0..7) -> fp32 partial gradient from the up path
        hs_idx = 0          # the up path pops hs in reverse: 7,6,...,0 ; walking it backwards we see 0,1,...,7
        carry32 = None      # partial fp32 gradient of the tensor feeding an `up` block (added by the head's gn_bwd)
        g16 = None
        for lvl_pos in reversed(range(4)):          # up_levels[3] is the full-resolution level, processed first
            blocks, head, upb = self.up_levels[lvl_pos]
            if upb is not None:
                # `up` block: consumes h (also seen by this level's head) -> partial gradient only
                if self.ddpm:
                    carry32 = self._up_bwd(upb, g16)
                else:
                    carry32, _, _ = self._rb_bwd(upb, ctx, g16, None, want_a32=True, want_a16=False)
            else:
                carry32 = None
            h, sh = ctx["head_in"][head]
            # the tensor under the head is produced by blocks[1] (has a skip conv): fp16 only is enough
            _, g16 = self._head_bwd(head, h, sh, dP[3 - lvl_pos], carry32, want32=False, consumer=blocks[1])
            for bi in reversed(blocks):
                first_of_level = (bi == blocks[0])
                # producer of the h-part: blocks[0] for blocks[1]; for blocks[0]: the previous level's `up` block
                # (skip conv) or, at the bottleneck level, ResBlock 16 (identity skip -> needs fp32 as well)
                need32 = first_of_level and lvl_pos == 0
                if not first_of_level:
                    cons = blocks[0]
                else:
                    cons = self.up_levels[lvl_pos - 1][2] if lvl_pos > 0 else self.attn_idx + 1
                # ddpm: the h-part of a level's first block comes straight out of the Upsample convolution (no /sqrt2)
                a16s = 1.0 if (self.ddpm and first_of_level and lvl_pos > 0) else INV_SQRT2
                d32, g16, dxb = self._rb_bwd(bi, ctx, g16, None, want_a32=need32, want_a16=True, a16_scale=a16s,
                                             consumer=cons)
                partial_hs[hs_idx] = dxb
                hs_idx += 1
        # bottleneck: RB16 <- attention <- RB14
        i16 = self.attn_idx + 1
        d32, g16, _ = self._rb_bwd(i16, ctx, g16, d32, want_a32=True)      # consumed by the attention block (c8 pair unused there, kept)
        d32, g16 = self._attn_bwd(ctx, g16, d32, consumer=self.attn_idx - 1)
        d32, g16, _ = self._rb_bwd(self.attn_idx - 1, ctx, g16, d32, extra_a=partial_hs[7], want_a32=True,
                                   consumer=self.attn_idx - 2)
        # down path, backwards.  hs index k: 7 = RB13 out, 6 = Combine12, 5 = RB10, 4 = Combine9, 3 = RB7,
        # 2 = Combine6, 1 = RB4, 0 = input conv
        dpyr = {}
        i = self.attn_idx - 2       # 13
        for lvl in (3, 2, 1):
            # plain block at this level: input is the Combine output hs[2*lvl]
            d32, g16, _ = self._rb_bwd(i, ctx, g16, d32, extra_a=partial_hs[2 * lvl], want_a32=True,
                                       want_a16=not self.ddpm, consumer=i - 2)
            w, _ = self.comb[i - 1]
            dpyr[lvl] = ops.combine_bwd(d32, w, torch.empty(B, d32.shape[1], d32.shape[2], 2, device=dev))
            # down block (has skip conv) / ddpm Downsample: input hs[2*lvl-1]
            if self.ddpm:
                d32, g16 = self._down_bwd(i - 2, d32, partial_hs[2 * lvl - 1], consumer=i - 3)
            else:
                d32, g16, _ = self._rb_bwd(i - 2, ctx, g16, None, extra_a=partial_hs[2 * lvl - 1], want_a32=True,
                                           consumer=i - 3)
            i -= 3
        # RB4 (identity skip) : input hs[0]; its producer is the input conv -> fp16 at scale 1
        _, g16, _ = self._rb_bwd(4, ctx, g16, d32, extra_a=partial_hs[0], want_a32=False, a16_scale=1.0)
        dcol = torch.empty(B, H, W, 32, device=dev)
        self._conv(g16, self.in_wd, dcol, taps=1, n_total=32)
        dx = torch.empty(B, H, W, 2, device=dev)
        ops.col2im_c2(dcol, dx)
        # input pyramid adjoint: pyr[l+1] = mean4(pyr[l])
        acc = dpyr[3]
        for lvl in (2, 1):
            acc = ops.resample_c2(acc, 2, dpyr[lvl], accumulate=True)
        ops.resample_c2(acc, 2, dx, accumulate=True)
        return ops.lincomb3(torch.empty(B, H * W * 2, device=dev), dx.view(B, -1), rms.contiguous()).view(B, H, W, 2)
