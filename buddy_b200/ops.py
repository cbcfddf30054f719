"""Thin Python wrappers over the C ABI: one function per kernel entry point.

Tensors are torch CUDA tensors used purely as device-memory handles (data_ptr + strides).
Activations are channels-last [B, H, W, C]; see include/buddy_b200.h for each kernel's contract.
"""
import ctypes

import torch

from . import _capi
from ._capi import GemmDesc, check, lib, ptr, stream_ptr


def _pick_n_tile(n_total):
    if n_total <= 16:
        return 16
    if n_total <= 256:
        return (n_total + 15) // 16 * 16
    for cand in (256, 192, 128):
        if n_total % cand == 0:
            return cand
    return 256


def conv_gemm(a, w, out, *, taps, n_total, n_tile=None, a2=None, w2=None, bias=None, bias_b=None, resid=None,
              scale=1.0, stats=None, b_batched=False, col_off=0, ldc=None, max_ctas=0):
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
    d.b = ptr(w)
    d.b_rows = w.shape[1]
    d.b_t = w.shape[0]
    d.b_stride_n, d.b_stride_t = w.stride(1), w.stride(0)
    assert w.shape[2] == C, (w.shape, C)
    d.batch, d.H, d.W = B, H, W
    d.taps = taps
    d.b_batched = 1 if b_batched else 0
    d.n_total = n_total
    d.n_tile = n_tile or _pick_n_tile(n_total)
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
