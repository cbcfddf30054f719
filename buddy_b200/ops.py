"""Thin Python wrappers over the C ABI: one function per kernel entry point.

Tensors are torch CUDA tensors used purely as device-memory handles (data_ptr + strides).
Activations are channels-last [B, H, W, C]; see include/buddy_b200.h for each kernel's contract.
"""
import ctypes

import torch

from . import _capi
from ._capi import GemmDesc, check, lib, ptr, stream_ptr


def _pick_n_tile(n_total, work_pairs=None):
    """MMA N per tile.  Large problems: the widest tile (fewest re-reads of the activation patches).  Small problems
    (one utterance at the low-resolution levels: 9-34 pixel-tile pairs for 74 CTA pairs): narrower tiles, so that
    pairs x n-tiles covers the machine — B = 1 is the reference's own usage (tester.py:153)."""
    if n_total <= 16:
        return 16
    if n_total <= 256:
        nt = (n_total + 15) // 16 * 16
    else:
        nt = next((cand for cand in (256, 192, 128) if n_total % cand == 0), 256)
    if work_pairs is not None and n_total % 64 == 0:
        while nt > 64 and nt % 2 == 0 and n_total % (nt // 2) == 0 and (nt // 2) % 32 == 0 \
                and work_pairs * ((n_total + nt - 1) // nt) < 74:
            nt //= 2
    return nt


def conv_gemm(a, w, out, *, taps, n_total, n_tile=None, a2=None, w2=None, bias=None, bias_b=None, resid=None,
              scale=1.0, stats=None, b_batched=False, col_off=0, ldc=None, max_ctas=0, passes=1, a8=None, w8=None,
              a8_2=None, w8_2=None, direct_epilogue=False, no_pairs=False, debug_flags=0, one_tap_per_stage=False, gnb=None,
              single_tile=None):
    """out[b,h,w,n] = scale*(sum_{tap,k} a[b,h+dy,w+dx,k] w[tap,n,k] + sum_k a2[b,h,w,k] w2[n,k] + bias + bias_b + resid).

    a, a2 : fp16 [B,H,W,C] (channel stride 1, other strides arbitrary multiples of 8 elements)
    w     : fp16 [T, rows, K] (K stride 1); T = taps, or batch when b_batched
    w2    : fp16 [rows, K2]
    out   : fp32 or fp16, row stride ldc (default: out.shape[-1]), written at column col_off
    """
    assert a.dtype == torch.float16 and w.dtype == torch.float16
    assert a.dim() == 4 and a.stride(3) == 1 and w.dim() == 3 and w.stride(2) == 1
    B, H, W, C = a.shape
    d = GemmDesc()
    d.a = ptr(a)
    d.a_c = C
    d.a_stride_w, d.a_stride_h, d.a_stride_b = a.stride(2), a.stride(1), a.stride(0)
    if a2 is not None:
        assert a2.dtype == torch.float16 and a2.shape[:3] == a.shape[:3] and a2.stride(3) == 1
        assert w2 is not None and w2.dtype == torch.float16 and w2.dim() == 2 and w2.stride(1) == 1
        d.a2 = ptr(a2)
        d.a2_c = a2.shape[3]
        d.a2_stride_w, d.a2_stride_h, d.a2_stride_b = a2.stride(2), a2.stride(1), a2.stride(0)
        d.b2 = ptr(w2)
        d.b2_rows = w2.shape[0]
        d.b2_stride_n = w2.stride(0)
        d.k2_total = w2.shape[1]
    d.b = ptr(w)
    d.b_rows = w.shape[1]
    d.b_t = w.shape[0]
    d.b_stride_n, d.b_stride_t = w.stride(1), w.stride(0)
    assert w.shape[2] % 64 == 0 and (w.shape[2] == C or w.shape[2] > C), (w.shape, C)
    d.k_total = w.shape[2]
    d.batch, d.H, d.W = B, H, W
    d.taps = taps
    d.b_batched = 1 if b_batched else 0
    d.n_total = n_total
    if n_tile is None:
        tiles = B * ((H + 15) // 16) * ((W + 7) // 8) if taps == 9 else (B * H * W + 127) // 128
        n_tile = _pick_n_tile(n_total, None if b_batched else (tiles + 1) // 2)
    d.n_tile = n_tile
    d.out = ptr(out)
    d.out_fp16 = 1 if out.dtype == torch.float16 else 0
    assert out.dtype in (torch.float16, torch.float32)
    d.ldc = ldc if ldc is not None else out.shape[-1]
    d.col_off = col_off
    d.bias = ptr(bias)
    d.bias_b = ptr(bias_b)
    d.resid = ptr(resid)
    d.ld_res = resid.shape[-1] if resid is not None else 0
    d.scale = float(scale)
    d.stats = ptr(stats)
    d.max_ctas = max_ctas
    d.no_staged_epilogue = 1 if direct_epilogue else 0
    d.no_cta_pairs = 1 if no_pairs else 0
    d.debug_flags = int(debug_flags)
    d.one_tap_per_stage = 1 if one_tap_per_stage else 0
    d.single_tile_per_cta = {None: 0, True: 1, False: 2}[single_tile]
    if gnb is not None:
        # (x, bundle stats of x, gamma, beta, gsum out, groups, eps, silu): fused GroupNorm-backward statistics
        gx, gstats, ggam, gbet, ggsum, ggroups, geps, gsilu = gnb
        assert gx.dtype == torch.float32 and gx.is_contiguous() and tuple(gx.shape) == (B, H, W, n_total)
        assert ggsum.dtype == torch.float64 and tuple(ggsum.shape) == (B, ggroups, 2)
        d.gnb_x, d.gnb_stats, d.gnb_gamma, d.gnb_beta, d.gnb_gsum = ptr(gx), ptr(gstats), ptr(ggam), ptr(gbet), ptr(ggsum)
        d.gnb_groups, d.gnb_eps, d.gnb_silu = int(ggroups), float(geps), int(gsilu)
    if a8 is not None:
        assert a8.dtype == torch.uint8 and w8.dtype == torch.uint8 and a8.shape[:3] == a.shape[:3]
        assert w8.is_contiguous() and w8.shape[:2] == w.shape[:2] and w8.shape[2] == a8.shape[3]
        d.a8, d.a8_c, d.b8 = ptr(a8), a8.shape[3], ptr(w8)
        d.a8_stride_w, d.a8_stride_h, d.a8_stride_b = a8.stride(2), a8.stride(1), a8.stride(0)
        if a2 is not None:
            assert a8_2.dtype == torch.uint8 and w8_2.dtype == torch.uint8 and w8_2.is_contiguous()
            assert w8_2.shape == (w2.shape[0], a8_2.shape[3])
            d.a8_2, d.a8_2_c, d.b8_2 = ptr(a8_2), a8_2.shape[3], ptr(w8_2)
            d.a8_2_stride_w, d.a8_2_stride_h, d.a8_2_stride_b = a8_2.stride(2), a8_2.stride(1), a8_2.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= n_total
    if bias_b is not None:
        assert bias_b.dtype == torch.float32 and bias_b.shape == (B, n_total)
    if resid is not None:
        assert resid.dtype == torch.float32
    if stats is not None:
        assert stats.dtype == torch.float64
    check(lib().buddy_conv_gemm(ctypes.byref(d), stream_ptr()), "buddy_conv_gemm")
    return out


from ._capi import GnBwdDesc, GnDesc, PackDesc, c_float, c_i64, c_int  # noqa: E402


def pack_weights(src, T, N, K, *, off0=0, st=0, sn=(1, 0, 0), sk=(1, 0, 0), n_valid=None, k_valid=None, passes=1,
                 e4m3=False):
    """One launch: fp32 weight `src` (any contiguous tensor, addressed by element strides) -> packed B operand.
    Element (t, n, k) = src.flatten()[off0 + t*st + (n // ndiv)*sn_outer + (n % ndiv)*sn_inner + (k // kdiv)*sk_outer
    + (k % kdiv)*sk_inner] with sn = (ndiv, sn_outer, sn_inner), sk = (kdiv, sk_outer, sk_inner).
    Returns (w16 [T, N, passes*K] fp16, w8 [T, N, 2K] uint8 or None) — see buddy_pack_desc."""
    assert src.dtype == torch.float32 and src.is_contiguous()
    w16 = torch.empty(T, N, passes * K, device=src.device, dtype=torch.float16)
    w8 = torch.empty(T, N, 2 * K, device=src.device, dtype=torch.uint8) if e4m3 else None
    d = PackDesc()
    d.src, d.off0, d.st = ptr(src), off0, st
    d.ndiv, d.sn_outer, d.sn_inner = sn
    d.kdiv, d.sk_outer, d.sk_inner = sk
    d.T, d.N, d.K = T, N, K
    d.n_valid = N if n_valid is None else n_valid
    d.k_valid = K if k_valid is None else k_valid
    d.passes = passes
    d.w16, d.w8 = ptr(w16), ptr(w8)
    check(lib().buddy_pack_weights(ctypes.byref(d), stream_ptr()), "buddy_pack_weights")
    return w16, w8

MODE_NONE, MODE_UP, MODE_DOWN = 0, 1, 2


def gn_stats(x, stats=None):
    """Per-4-channel-bundle (sum, sumsq) of x fp32 [B,H,W,C] -> fp64 [B, C/4, 2]."""
    B, C = x.shape[0], x.shape[-1]
    P = x.numel() // (B * C)
    if stats is None:
        stats = torch.zeros(B, C // 4, 2, device=x.device, dtype=torch.float64)
    check(lib().buddy_gn_stats(ptr(x), c_int(B), c_i64(P), c_int(C), ptr(stats), stream_ptr()), "buddy_gn_stats")
    return stats


def _gn_desc(xa, sa, xb, sb, gamma, beta, groups, silu, mode, eps, split=False):
    B, H, W, Ca = xa.shape
    d = GnDesc()
    d.xa, d.Ca, d.stats_a = ptr(xa), Ca, ptr(sa)
    if xb is not None:
        assert xb.shape[:3] == xa.shape[:3]
        d.xb, d.Cb, d.stats_b = ptr(xb), xb.shape[3], ptr(sb)
    d.gamma, d.beta = ptr(gamma), ptr(beta)
    d.batch, d.H, d.W = B, H, W
    d.groups, d.eps, d.silu, d.mode = groups, eps, int(silu), mode
    d.split = int(split)
    return d


def gn_apply(xa, sa, gamma, beta, out, *, xb=None, sb=None, groups=32, silu=True, mode=MODE_NONE, out_raw=None,
             eps=1e-6, split=False, out8=None, out_raw8=None):
    """out(fp16) = resample(act(GroupNorm([xa|xb]))); out_raw(fp16) = resample([xa|xb])."""
    assert xa.dtype == torch.float32 and xa.is_contiguous() and out.dtype == torch.float16
    d = _gn_desc(xa, sa, xb, sb, gamma, beta, groups, silu, mode, eps, split)
    d.out, d.out_raw = ptr(out), ptr(out_raw)
    d.out8, d.out_raw8 = ptr(out8), ptr(out_raw8)
    check(lib().buddy_gn_apply(ctypes.byref(d), stream_ptr()), "buddy_gn_apply")
    return out


def gn_act32(x, stats, gamma, beta, out, groups=32, eps=1e-6, silu=True):
    """fp32 GroupNorm(+SiLU) of x [B,H,W,C] (see buddy_gn_act32)."""
    B, C = x.shape[0], x.shape[-1]
    P = x.numel() // (B * C)
    assert x.dtype == torch.float32 and x.is_contiguous() and out.shape == x.shape and out.dtype == torch.float32
    check(lib().buddy_gn_act32(ptr(x), ptr(stats), ptr(gamma), ptr(beta), c_int(B), c_i64(P), c_int(C), c_int(groups),
                               c_float(eps), c_int(int(silu)), ptr(out), stream_ptr()), "buddy_gn_act32")
    return out


def gn_bwd(xa, sa, gamma, beta, da, gsum, *, xb=None, sb=None, groups=32, silu=True, mode=MODE_NONE, dskip=None,
           skip_scale=1.0, extra_a=None, extra_b=None, dxa=None, dxb=None, g16a=None, g16b=None, g16_scale=1.0,
           eps=1e-6, split=False, g8a=None, g8b=None, pass0_done=False):
    d = _gn_desc(xa, sa, xb, sb, gamma, beta, groups, silu, mode, eps, split)
    g = GnBwdDesc()
    g.da, g.dskip, g.skip_scale = ptr(da), ptr(dskip), skip_scale
    g.extra_a, g.extra_b, g.gsum = ptr(extra_a), ptr(extra_b), ptr(gsum)
    g.dxa, g.dxb, g.g16a, g.g16b, g.g16_scale = ptr(dxa), ptr(dxb), ptr(g16a), ptr(g16b), g16_scale
    g.g8a, g.g8b = ptr(g8a), ptr(g8b)
    g.pass0_done = 1 if pass0_done else 0
    assert da.dtype == torch.float32 and gsum.dtype == torch.float64
    check(lib().buddy_gn_bwd(ctypes.byref(d), ctypes.byref(g), stream_ptr()), "buddy_gn_bwd")


def im2col_c2(x, col, split=False, col8=None, in_scale=1.0):
    B, H, W, _ = x.shape
    check(lib().buddy_im2col_c2(ptr(x), c_int(B), c_int(H), c_int(W), ptr(col), c_int(int(split)), ptr(col8),
                                c_float(in_scale), stream_ptr()), "buddy_im2col_c2")
    return col


def col2im_c2(dcol, dx, accumulate=False):
    B, H, W, ld = dcol.shape
    check(lib().buddy_col2im_c2(ptr(dcol), c_int(ld), c_int(B), c_int(H), c_int(W), ptr(dx), c_int(int(accumulate)),
                                stream_ptr()), "buddy_col2im_c2")
    return dx


def resample_c2(x, mode, out, add=None, accumulate=False):
    B, H, W, _ = x.shape
    check(lib().buddy_resample_c2(ptr(x), c_int(B), c_int(H), c_int(W), c_int(mode), ptr(add), ptr(out),
                                  c_int(int(accumulate)), stream_ptr()), "buddy_resample_c2")
    return out


def combine_fwd(h, pyr, w, bias, out):
    C = h.shape[-1]
    P = h.numel() // C
    check(lib().buddy_combine_fwd(ptr(h), ptr(pyr), ptr(w), ptr(bias), c_i64(P), c_int(C), ptr(out), stream_ptr()),
          "buddy_combine_fwd")
    return out


def combine_bwd(dout, w, dpyr):
    C = dout.shape[-1]
    P = dout.numel() // C
    check(lib().buddy_combine_bwd(ptr(dout), ptr(w), c_i64(P), c_int(C), ptr(dpyr), stream_ptr()), "buddy_combine_bwd")
    return dpyr


def affine_c2(x, m4, b2, y):
    P = x.numel() // 2
    m = (c_float * 4)(*[float(v) for v in m4])
    b = (c_float * 2)(*[float(v) for v in b2])
    check(lib().buddy_affine_c2(ptr(x), c_i64(P), m, b, ptr(y), stream_ptr()), "buddy_affine_c2")
    return y


def softmax_fwd(s, p):
    n = s.shape[-1]
    rows = s.numel() // n
    check(lib().buddy_softmax_fwd(ptr(s), c_i64(rows), c_int(n), ptr(p), c_int(p.shape[-1]), stream_ptr()),
          "buddy_softmax_fwd")
    return p


def softmax_bwd(p, dp, scale, ds):
    n = dp.shape[-1]
    rows = dp.numel() // n
    check(lib().buddy_softmax_bwd(ptr(p), c_int(p.shape[-1]), ptr(dp), c_i64(rows), c_int(n), c_float(scale), ptr(ds),
                                  c_int(ds.shape[-1]), stream_ptr()), "buddy_softmax_bwd")
    return ds


def transpose_h(x, out):
    """x fp16 [batch, R, C] (row stride arbitrary) -> out fp16 [batch, C, R]."""
    batch, R, C = x.shape
    assert x.stride(2) == 1 and out.stride(2) == 1 and out.shape == (batch, C, R)
    check(lib().buddy_transpose_h(ptr(x), c_int(batch), c_int(R), c_int(C), c_i64(x.stride(1)), c_i64(x.stride(0)),
                                  ptr(out), c_i64(out.stride(1)), c_i64(out.stride(0)), stream_ptr()),
          "buddy_transpose_h")
    return out


def cast_operand(x, out16, out8=None, scale=1.0, upsample=False, split=0):
    """x fp32 [B,H,W,C] -> operand of scale*x (see buddy_cast_operand); out16 [B,H',W',C (2C for split 1)]."""
    B, H, W, C = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out16.dtype == torch.float16
    check(lib().buddy_cast_operand(ptr(x), c_int(B), c_int(H), c_int(W), c_int(C), c_int(int(upsample)), c_float(scale),
                                   ptr(out16), ptr(out8), c_int(int(split)), stream_ptr()), "buddy_cast_operand")
    return out16


def cast_scale_h(x, scale, y):
    check(lib().buddy_cast_scale_h(ptr(x), c_i64(x.numel()), c_float(scale), ptr(y), stream_ptr()),
          "buddy_cast_scale_h")
    return y


c_u64 = ctypes.c_uint64


def dft_analysis(sig, mat, hop, frames, Tout, out):
    """sig fp32 [B, L]; mat fp32 [2*bins, K]; out fp32 [B, bins, Tout, 2]."""
    B = sig.shape[0]
    M, K = mat.shape
    check(lib().buddy_dft_analysis(ptr(sig), c_i64(sig.stride(0)), c_int(B), ptr(mat), c_int(M), c_int(K), c_int(hop),
                                   c_int(frames), c_int(Tout), ptr(out), stream_ptr()), "buddy_dft_analysis")
    return out


def dft_synthesis(S, mat, frames, fr):
    """S fp32 [B, bins, Tin, 2]; mat [2*bins, K]; fr fp32 [B, frames, K]."""
    B, _, Tin, _ = S.shape
    M, K = mat.shape
    check(lib().buddy_dft_synthesis(ptr(S), c_int(B), c_int(Tin), ptr(mat), c_int(M), c_int(K), c_int(frames), ptr(fr),
                                    stream_ptr()), "buddy_dft_synthesis")
    return fr


class FftMat:
    """The 1024-point STFT matrix mat[2f+c][n] = a[f] * w[n] * (cos, -sin)(2 pi f n / 1024) in factored form."""
    __slots__ = ("a", "w", "tw")

    def __init__(self, a, w, device):
        import math
        self.a = torch.as_tensor(a, dtype=torch.float64).float().to(device).contiguous()
        self.w = torch.as_tensor(w, dtype=torch.float64).float().to(device).contiguous()
        k = torch.arange(512, dtype=torch.float64)
        self.tw = torch.stack([torch.cos(2 * math.pi * k / 1024), -torch.sin(2 * math.pi * k / 1024)], -1).float().to(
            device).contiguous()


def stft_analysis(sig, m, hop, frames, Tout, out):
    """dft_analysis or fft_analysis depending on the matrix representation (`FftMat` = factored 1024-point form)."""
    if isinstance(m, FftMat):
        return fft_analysis(sig, m, hop, frames, Tout, out)
    return dft_analysis(sig, m, hop, frames, Tout, out)


def stft_synthesis(S, m, frames, fr):
    if isinstance(m, FftMat):
        return fft_synthesis(S, m, frames, fr)
    return dft_synthesis(S, m, frames, fr)


def fft_analysis(sig, fm, hop, frames, Tout, out):
    """sig fp32 [B, L]; out fp32 [B, bins, Tout, 2] = a[f] * FFT_1024(w * frame)[f]."""
    B = sig.shape[0]
    check(lib().buddy_fft_analysis(ptr(sig), c_i64(sig.stride(0)), c_int(B), ptr(fm.w), ptr(fm.a), ptr(fm.tw),
                                   c_int(fm.a.numel()), c_int(fm.w.numel()), c_int(hop), c_int(frames), c_int(Tout),
                                   ptr(out), stream_ptr()), "buddy_fft_analysis")
    return out


def fft_synthesis(S, fm, frames, fr):
    """S fp32 [B, bins, Tin, 2]; fr fp32 [B, frames, K] = w[n] * Re(sum_f a[f] S[f] e^{+2 pi i f n / 1024})."""
    B, _, Tin, _ = S.shape
    check(lib().buddy_fft_synthesis(ptr(S), c_int(B), c_int(Tin), ptr(fm.w), ptr(fm.a), ptr(fm.tw),
                                    c_int(fm.a.numel()), c_int(fm.w.numel()), c_int(frames), ptr(fr), stream_ptr()),
          "buddy_fft_synthesis")
    return fr


def ola_gather(fr, hop, off, n_out, out, tab=None, scale_b=None):
    B, frames, K = fr.shape
    check(lib().buddy_ola_gather(ptr(fr), c_int(B), c_int(frames), c_int(K), c_int(hop), c_int(off), c_int(n_out),
                                 ptr(tab), ptr(scale_b), ptr(out), c_i64(out.stride(0)), stream_ptr()),
          "buddy_ola_gather")
    return out


def pad_signal(x, left, total, mode, out, tab=None, scale_b=None):
    B, N = x.shape
    check(lib().buddy_pad_signal(ptr(x), c_i64(x.stride(0)), c_int(B), c_int(N), c_int(left), c_int(total),
                                 c_int(mode), ptr(tab), ptr(scale_b), ptr(out), stream_ptr()), "buddy_pad_signal")
    return out


def reflect_fold(dxp, N, L, out, scale_b=None):
    B = dxp.shape[0]
    check(lib().buddy_reflect_fold(ptr(dxp), c_int(B), c_int(N), c_int(L), ptr(scale_b), ptr(out),
                                   c_i64(out.stride(0)), stream_ptr()), "buddy_reflect_fold")
    return out


def comp_loss(Y, X, frames, compression, weight, loss, grad=None):
    B = Y.shape[0]
    per = Y.numel() // (2 * B)
    check(lib().buddy_comp_loss(ptr(Y), ptr(X), c_int(B), c_i64(per), c_int(frames), c_float(compression),
                                c_float(weight), ptr(loss), ptr(grad), stream_ptr()), "buddy_comp_loss")
    return loss


def row_stats(x, out=None):
    """(sum, sumsq) per row of x fp32 [B, n] -> fp64 [B, 2]."""
    B, n = x.shape
    if out is None:
        out = torch.empty(B, 2, device=x.device, dtype=torch.float64)
    check(lib().buddy_row_stats(ptr(x), c_i64(x.stride(0)), c_int(B), c_int(n), ptr(out), stream_ptr()),
          "buddy_row_stats")
    return out


def fftconv(x, n_in, log2_n2, tw512, work, H, h_batch_stride, mode, y, n_out):
    B = x.shape[0]
    check(lib().buddy_fftconv(ptr(x), c_i64(x.stride(0)), c_int(B), c_int(n_in), c_int(log2_n2), ptr(tw512), ptr(work),
                              ptr(H), c_i64(h_batch_stride), c_int(mode), ptr(y),
                              c_i64(y.stride(0) if y is not None else 0), c_int(n_out), stream_ptr()), "buddy_fftconv")
    return y


def fourier_features(t, W, out):
    B, E = t.shape[0], W.shape[0]
    check(lib().buddy_fourier_features(ptr(t), ptr(W), c_int(B), c_int(E), ptr(out), stream_ptr()),
          "buddy_fourier_features")
    return out


def dense(x, W, bias, out, act_in=False, act_out=False):
    B, In = x.shape
    Out = W.shape[0]
    check(lib().buddy_dense(ptr(x), ptr(W), ptr(bias), c_int(B), c_int(In), c_int(Out), c_int(int(act_in)),
                            c_int(int(act_out)), ptr(out), stream_ptr()), "buddy_dense")
    return out


def dense_seg(x, W, bias, seg, out, act_in=False):
    """x [B, In]; W [Out, In] = several layers' rows; seg int32 [Out, 2]; out flat [B * Out] (see buddy_dense_seg)."""
    B, In = x.shape
    Out = W.shape[0]
    assert seg.dtype == torch.int32 and seg.shape == (Out, 2) and out.numel() == B * Out
    check(lib().buddy_dense_seg(ptr(x), ptr(W), ptr(bias), ptr(seg), c_int(B), c_int(In), c_int(Out),
                                c_int(int(act_in)), ptr(out), stream_ptr()), "buddy_dense_seg")
    return out


def philox_normal(seeds, draw, out):
    B, n = out.shape
    assert seeds.dtype == torch.int64
    check(lib().buddy_philox_normal(ptr(seeds), c_u64(draw), c_int(B), c_int(n), ptr(out), c_i64(out.stride(0)),
                                    stream_ptr()), "buddy_philox_normal")
    return out


def lincomb3(out, x, ca, y=None, cb=None, z=None, cc=None):
    B, n = x.shape
    for t in (x, y, z, out):
        assert t is None or (t.is_contiguous() and t.dtype == torch.float32)
    check(lib().buddy_lincomb3(ptr(x), ptr(y), ptr(z), ptr(ca), ptr(cb), ptr(cc), c_int(B), c_int(n), ptr(out),
                               stream_ptr()), "buddy_lincomb3")
    return out


# ------------------------------------------------------------------------------------------------------------
# optional per-kernel device timing (CUDA events on the launching stream) — used by bench.py for the roofline
# ------------------------------------------------------------------------------------------------------------
class KernelTimer:
    """with KernelTimer() as kt: ...; kt.summary() -> {op: (calls, total_ms, work)}; work = FLOPs or bytes."""

    def __init__(self):
        self.records = []
        self.issued_flops = 0.0     # conv launches: algorithmic FLOPs x fp16-pass equivalents actually issued

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *a):
        global _timer
        _timer = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work, _ in self.records:
            c, ms, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (c + 1, ms + e0.elapsed_time(e1), w + work)
        return out

    def by_tag(self):
        """{(op, tag): (calls, total_ms, work)} — tag = shape key of the launch (conv: B,H,W,K,N,taps,skipK)."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work, tag in self.records:
            c, ms, w = out.get((name, tag), (0, 0.0, 0.0))
            out[(name, tag)] = (c + 1, ms + e0.elapsed_time(e1), w + work)
        return out


_timer = None


def _conv_flops(args, kwargs):
    a, w = args[0], args[1]
    B, H, W, _ = a.shape
    flops = 2.0 * B * H * W * kwargs["n_total"] * w.shape[2] * (1 if kwargs.get("b_batched") else w.shape[0])
    if kwargs.get("b_batched"):
        flops = 2.0 * B * H * W * kwargs["n_total"] * w.shape[2]
    if kwargs.get("w2") is not None:
        flops += 2.0 * B * H * W * kwargs["n_total"] * kwargs["w2"].shape[1]
    return flops / kwargs.get("passes", 1)   # algorithmic FLOPs: split-precision passes are not useful work


def _conv_pass_equiv(kwargs):
    """fp16-pass equivalents issued per algorithmic product (e4m3 correction MMAs carry twice the K per instruction)."""
    return 2.0 if kwargs.get("a8") is not None else float(kwargs.get("passes", 1))


def _conv_tag(args, kwargs):
    a, w = args[0], args[1]
    B, H, W, C = a.shape
    w2 = kwargs.get("w2")
    return (B, H, W, C, kwargs["n_total"], kwargs["taps"], w2.shape[1] if w2 is not None else 0,
            "c8" if kwargs.get("a8") is not None else f"x{kwargs.get('passes', 1)}", str(args[2].dtype)[6:])


def _shape_tag(args, kwargs):
    for t in args:
        if isinstance(t, torch.Tensor):
            return tuple(t.shape)
    return ()


def _timed(fn, work_fn=None, tag_fn=_shape_tag):
    def wrapper(*args, **kwargs):
        if _timer is None:
            return fn(*args, **kwargs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kwargs)
        e1.record()
        work = work_fn(args, kwargs) if work_fn else 0.0
        _timer.records.append((fn.__name__, e0, e1, work, tag_fn(args, kwargs)))
        if fn.__name__ == "conv_gemm":
            _timer.issued_flops += work * _conv_pass_equiv(kwargs)
        return r
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if isinstance(t, torch.Tensor)))


def _gn_apply_bytes(args, kw):
    """algorithmic HBM bytes: every input / output tensor exactly once (statistics and affine vectors are noise)"""
    return _nbytes(args[0], args[4], kw.get("xb"), kw.get("out_raw"), kw.get("out8"), kw.get("out_raw8"))


def _gn_bwd_bytes(args, kw):
    """algorithmic HBM bytes: x, da, dskip, extra and every output once (the two-pass kernel reads x and da twice)"""
    return _nbytes(args[0], args[4], kw.get("xb"), kw.get("dskip"), kw.get("extra_a"), kw.get("extra_b"),
                   kw.get("dxa"), kw.get("dxb"), kw.get("g16a"), kw.get("g16b"), kw.get("g8a"), kw.get("g8b"))


conv_gemm = _timed(conv_gemm, _conv_flops, _conv_tag)
gn_apply = _timed(gn_apply, _gn_apply_bytes)
gn_bwd = _timed(gn_bwd, _gn_bwd_bytes)
for _n in ("gn_stats", "im2col_c2", "col2im_c2", "resample_c2", "combine_fwd", "combine_bwd",
           "affine_c2", "softmax_fwd", "softmax_bwd", "transpose_h", "cast_scale_h", "cast_operand", "dft_analysis", "dft_synthesis", "fft_analysis", "fft_synthesis",
           "ola_gather", "pad_signal", "reflect_fold", "comp_loss", "row_stats", "fftconv", "fourier_features",
           "dense", "philox_normal", "lincomb3"):
    globals()[_n] = _timed(globals()[_n])


# ------------------------------------------------------------------------------------------------------------
# blind operator kernels
# ------------------------------------------------------------------------------------------------------------
def subband_fir(a, h_or_dy, out, *, Nf, pre, mode, shared_h=False, accumulate=False):
    """a, out: fp32 [B, F, T, 2]; mode 0/1: h_or_dy = H [B or 1, F, Nf, 2]; mode 2: h_or_dy = dY, out = dH."""
    B, F, Tx, _ = a.shape
    stride = 0 if shared_h else F * Nf * 2
    check(lib().buddy_subband_fir(ptr(a), ptr(h_or_dy), c_i64(stride), ptr(out), c_int(B), c_int(F), c_int(Tx),
                                  c_int(Nf), c_int(pre), c_int(mode), c_int(int(accumulate)), stream_ptr()),
          "buddy_subband_fir")
    return out


def blind_design_fwd(decays, weights, phases, tabs, A, H0):
    B, F, Nf = phases.shape
    check(lib().buddy_blind_design_fwd(ptr(decays), ptr(weights), ptr(phases), ptr(tabs["kidx"]), ptr(tabs["frac"]),
                                       ptr(tabs["corr"]), ptr(tabs["dpmag"]), c_int(B), c_int(F), c_int(Nf), ptr(A),
                                       ptr(H0), stream_ptr()), "buddy_blind_design_fwd")


def blind_design_bwd(decays, weights, phases, A, tabs, G, dphases, ddecays, dweights):
    B, F, Nf = phases.shape
    scratch = torch.empty(B, 50, Nf, device=phases.device)
    check(lib().buddy_blind_design_bwd(ptr(decays), ptr(weights), ptr(phases), ptr(A), ptr(tabs["kidx"]),
                                       ptr(tabs["frac"]), ptr(tabs["corr"]), ptr(tabs["dpmag"]), ptr(G), c_int(B),
                                       c_int(F), c_int(Nf), ptr(dphases), ptr(ddecays), ptr(dweights), ptr(scratch),
                                       stream_ptr()), "buddy_blind_design_bwd")


def fft_mixed(x, in_real, work, out, N1, sign, tw512):
    B = x.shape[0]
    check(lib().buddy_fft_mixed(ptr(x), c_int(int(in_real)), ptr(work), ptr(out), c_int(B), c_int(N1), c_int(sign),
                                ptr(tw512), stream_ptr()), "buddy_fft_mixed")
    return out


def minphase_pw(mode, B, N, T, c0=None, c1=None, r0=None, r1=None, oc=None, or0=None, or1=None, scale_inv_n=False):
    check(lib().buddy_minphase_pw(c_int(mode), ptr(c0), ptr(c1), ptr(r0), ptr(r1), ptr(oc), ptr(or0), ptr(or1),
                                  c_int(B), c_int(N), c_int(T), c_int(int(scale_inv_n)), stream_ptr()),
          "buddy_minphase_pw")


def adam_project(p, g, m, v, step, lr, beta1, beta2, eps, dmin, dmax, wmin, wmax):
    B, n_per = p.shape
    check(lib().buddy_adam_project(ptr(p), ptr(g), ptr(m), ptr(v), c_int(B), c_int(n_per), c_int(step), c_float(lr),
                                   c_float(beta1), c_float(beta2), c_float(eps), c_float(dmin), c_float(dmax),
                                   c_float(wmin), c_float(wmax), stream_ptr()), "buddy_adam_project")


for _n in ("subband_fir", "blind_design_fwd", "blind_design_bwd", "fft_mixed", "minphase_pw", "adam_project"):
    globals()[_n] = _timed(globals()[_n])


def wpe(Y, taps, delay, iterations, Z=None):
    """Y fp32 [B, F, T, 2] -> WPE-filtered spectra, same layout (see buddy_wpe)."""
    assert Y.dtype == torch.float32 and Y.is_contiguous() and Y.dim() == 4 and Y.shape[3] == 2
    B, F, T, _ = Y.shape
    Z = torch.empty_like(Y) if Z is None else Z
    check(lib().buddy_wpe(ptr(Y), c_int(B), c_int(F), c_int(T), c_int(int(taps)), c_int(int(delay)),
                          c_int(int(iterations)), ptr(Z), stream_ptr()), "buddy_wpe")
    return Z
