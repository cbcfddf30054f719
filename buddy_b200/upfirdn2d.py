"""upfirdn2d on buddy_b200 kernels — same Python surface as the reference's StyleGAN2 operator
(networks/ncsnpp_utils/op/upfirdn2d.py:86-139: `upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`, `upfirdn1d`,
autograd through the data) and the FIR resampling helpers built on it (up_or_down_sampling.py:181-256).

CUDA tensors only (no CPU fallback: the reference's `upfirdn2d_native` CPU branch is restated in oracle/upfirdn.py as
the test oracle).  The NCSN++ engine uses `_launch` on its channels-last activations for `fir=True` (engine.py
`_fir_fwd` / `_fir_bwd`) and for the ddpm Downsample / Upsample modules; the shipped configuration has `fir: False`, and
the reference as published cannot run `fir=True` at all (its import of this operator is commented out,
up_or_down_sampling.py:10 -> NameError at :140/:176/:223/:256).
"""
import ctypes

import numpy as np
import torch

from ._capi import c_int, check, lib, ptr, stream_ptr


def _launch(x4, kernel, up, down, pad):
    """x4 [major, in_h, in_w, minor] fp32 contiguous."""
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    (up_x, up_y), (down_x, down_y), (px0, px1, py0, py1) = up, down, pad
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1
    out = torch.empty(major, out_h, out_w, minor, device=x4.device)
    check(lib().buddy_upfirdn2d(ptr(x4), ptr(kernel), c_int(major), c_int(in_h), c_int(in_w), c_int(minor), c_int(kh),
                                c_int(kw), c_int(up_x), c_int(up_y), c_int(down_x), c_int(down_y), c_int(px0),
                                c_int(px1), c_int(py0), c_int(py1), ptr(out), stream_ptr()), "buddy_upfirdn2d")
    return out


class UpFirDn2d(torch.autograd.Function):
    """op/upfirdn2d.py:86-139; the backward is the same operator with the flipped kernel, up and down exchanged and
    the complementary padding (:104-107)."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        if not input.is_cuda:
            raise RuntimeError("buddy_b200.upfirdn2d runs on CUDA tensors only (no CPU fallback)")
        up_x, up_y = up
        down_x, down_y = down
        px0, px1, py0, py1 = pad
        kh, kw = kernel.shape
        batch, channel, in_h, in_w = input.shape
        k = kernel.detach().float().contiguous()
        out = _launch(input.detach().float().reshape(-1, in_h, in_w, 1).contiguous(), k, up, down, pad)
        out_h, out_w = out.shape[1], out.shape[2]
        ctx.args = (k, up, down, (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                                  kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1), input.shape)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        k, up, down, g_pad, in_shape = ctx.args
        g = grad_output.detach().float().reshape(-1, grad_output.shape[2], grad_output.shape[3], 1).contiguous()
        gi = _launch(g, torch.flip(k, [0, 1]).contiguous(), down, up, g_pad)
        return gi.view(in_shape), None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


def upfirdn1d(input, kernel, up_x=1, up_y=1, down_x=1, down_y=1, pad_x0=0, pad_x1=0, pad_y0=0, pad_y1=0):
    return UpFirDn2d.apply(input, kernel, (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1))


def _setup_kernel(k):
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = np.outer(k, k)
    k /= np.sum(k)
    assert k.ndim == 2 and k.shape[0] == k.shape[1]
    return k


def upsample_2d(x, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:195-224: FIR upsampling of [N, C, H, W] by `factor` (default kernel: nearest neighbour)."""
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    k = _setup_kernel(k) * (gain * (factor ** 2))
    p = k.shape[0] - factor
    return upfirdn2d(x, torch.tensor(k, device=x.device), up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def downsample_2d(x, k=None, factor=2, gain=1):
    """up_or_down_sampling.py:227-257: FIR downsampling of [N, C, H, W] by `factor` (default kernel: box filter)."""
    assert isinstance(factor, int) and factor >= 1
    if k is None:
        k = [1] * factor
    k = _setup_kernel(k) * gain
    p = k.shape[0] - factor
    return upfirdn2d(x, torch.tensor(k, device=x.device), down=factor, pad=((p + 1) // 2, p // 2))
