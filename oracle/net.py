"""ORACLE (test infrastructure): functional restatement of the reference score network.

Follows networks/ncsnpp.py:281-449 (NCSNpp.forward), :473-506 (NCSNppTime stft/istft/forward) and the blocks in
networks/ncsnpp_utils/layerspp.py (GaussianFourierProjection :39-41, Combine :52-59, AttnBlockpp :75-91,
ResnetBlockBigGANpp :242-274) + up_or_down_sampling.py:59-69, at the shipped config conf/network/ncsnpp.yaml.
Operates directly on a reference-keyed state_dict; plain fp32 PyTorch, differentiable through autograd.
"""
import math

import torch
import torch.nn.functional as F

N_FFT = 510
HOP = 128
N_BINS = N_FFT // 2 + 1  # 256
INV_SQRT2 = 1.0 / math.sqrt(2.0)


def hann510(device):
    return torch.hann_window(N_FFT, periodic=True, device=device)


def net_stft(sig):
    """ncsnpp.py:473-486 — (B,1,T) real -> (B,1,256,F16) complex64, frames zero-padded to a multiple of 16."""
    B, C, T = sig.shape
    spec = torch.stft(sig.reshape(B * C, T), n_fft=N_FFT, hop_length=HOP, center=True, window=hann510(sig.device),
                      return_complex=True)
    spec = spec.reshape(B, C, spec.shape[-2], spec.shape[-1])
    if spec.shape[-1] % 16:
        spec = F.pad(spec, (0, 16 - spec.shape[-1] % 16))
    return spec.to(torch.complex64)


def net_istft(spec, length):
    """ncsnpp.py:489-496 — inverse over ALL (padded) frames, then crop to `length`."""
    B, C, Fq, Tf = spec.shape
    sig = torch.istft(spec.reshape(B * C, Fq, Tf), n_fft=N_FFT, hop_length=HOP, center=True,
                      window=hann510(spec.device), length=length)
    return sig.reshape(B, C, -1)[..., :length]


def _gn(x, sd, pre, groups=32):
    return F.group_norm(x, min(x.shape[1] // 4, groups), sd[pre + ".weight"], sd[pre + ".bias"], eps=1e-6)


def _up2(x):  # naive_upsample_2d, up_or_down_sampling.py:59-63 (== nearest x2)
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _down2(x):  # naive_downsample_2d, :66-69 (== 2x2 mean)
    return F.avg_pool2d(x, 2)


RECORD = None  # debugging aid: set to a dict to capture every block output ({module index: tensor})


def _rec(i, t):
    if RECORD is not None:
        RECORD[i] = t.detach()
    return t


def _resblock(sd, i, x, temb_act, up=False, down=False):
    return _rec(i, _resblock_impl(sd, i, x, temb_act, up, down))


def _resblock_impl(sd, i, x, temb_act, up=False, down=False):
    """ResnetBlockBigGANpp.forward, layerspp.py:242-274 (dropout p=0 is the identity)."""
    p = f"all_modules.{i}"
    h = F.silu(_gn(x, sd, p + ".GroupNorm_0"))
    if up:
        h, x = _up2(h), _up2(x)
    elif down:
        h, x = _down2(h), _down2(x)
    h = F.conv2d(h, sd[p + ".Conv_0.weight"], sd[p + ".Conv_0.bias"], padding=1)
    h = h + F.linear(temb_act, sd[p + ".Dense_0.weight"], sd[p + ".Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(h, sd, p + ".GroupNorm_1"))
    h = F.conv2d(h, sd[p + ".Conv_1.weight"], sd[p + ".Conv_1.bias"], padding=1)
    if (p + ".Conv_2.weight") in sd:
        x = F.conv2d(x, sd[p + ".Conv_2.weight"], sd[p + ".Conv_2.bias"])
    return (x + h) * INV_SQRT2


def _attn(sd, i, x):
    """AttnBlockpp.forward, layerspp.py:75-91: single-head attention over all H*W positions, (x+h)/sqrt2."""
    p = f"all_modules.{i}"
    B, C, H, W = x.shape
    h = _gn(x, sd, p + ".GroupNorm_0").permute(0, 2, 3, 1).reshape(B, H * W, C)
    q = h @ sd[p + ".NIN_0.W"] + sd[p + ".NIN_0.b"]
    k = h @ sd[p + ".NIN_1.W"] + sd[p + ".NIN_1.b"]
    v = h @ sd[p + ".NIN_2.W"] + sd[p + ".NIN_2.b"]
    w = torch.softmax((q @ k.transpose(1, 2)) * (int(C) ** (-0.5)), dim=-1)
    o = (w @ v) @ sd[p + ".NIN_3.W"] + sd[p + ".NIN_3.b"]
    o = o.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return (x + o) * INV_SQRT2


def time_embedding(sd, time_cond):
    """ncsnpp.py:299-318 + layerspp.py:39-41.  Returns SiLU(temb) — the only form the ResBlocks consume."""
    proj = time_cond[:, None] * sd["all_modules.0.W"][None, :] * 2 * math.pi
    emb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
    t = F.linear(emb, sd["all_modules.1.weight"], sd["all_modules.1.bias"])
    t = F.linear(F.silu(t), sd["all_modules.2.weight"], sd["all_modules.2.bias"])
    return F.silu(t)


def ncsnpp_forward(sd, spec, time_cond):
    """NCSNpp.forward (ncsnpp.py:281-449): (B,1,256,F) complex64 -> same shape; module map = SURVEY App. B."""
    x = torch.cat([spec.real, spec.imag], dim=1)  # (B,2,F,T)   :291-297
    ta = time_embedding(sd, time_cond)
    pyr_in = x
    hs = [_rec(3, F.conv2d(x, sd["all_modules.3.weight"], sd["all_modules.3.bias"], padding=1))]
    i = 4
    for lvl in range(4):
        h = _resblock(sd, i, hs[-1], ta)
        i += 1
        hs.append(h)
        if lvl != 3:
            h = _resblock(sd, i, hs[-1], ta, down=True)
            i += 1
            pyr_in = F.avg_pool2d(pyr_in, 2)  # pyramid_downsample (with_conv=False), layerspp.py:156
            h = F.conv2d(pyr_in, sd[f"all_modules.{i}.Conv_0.weight"], sd[f"all_modules.{i}.Conv_0.bias"]) + h
            _rec(i, h)
            i += 1
            hs.append(h)
    h = _resblock(sd, i, hs[-1], ta)
    h = _rec(i + 1, _attn(sd, i + 1, h))
    h = _resblock(sd, i + 2, h, ta)
    i += 3
    pyramid = None
    for lvl in reversed(range(4)):
        for _ in range(2):
            h = _resblock(sd, i, torch.cat([h, hs.pop()], dim=1), ta)
            i += 1
        ph = F.silu(F.group_norm(h, 32, sd[f"all_modules.{i}.weight"], sd[f"all_modules.{i}.bias"], eps=1e-6))
        ph = F.conv2d(ph, sd[f"all_modules.{i + 1}.weight"], sd[f"all_modules.{i + 1}.bias"], padding=1)
        _rec(i + 1, ph)
        i += 2
        pyramid = ph if pyramid is None else F.interpolate(pyramid, scale_factor=2, mode="nearest") + ph
        if lvl != 0:
            h = _resblock(sd, i, h, ta, up=True)
            i += 1
    assert not hs and i == 36
    out = F.conv2d(pyramid, sd["output_layer.weight"], sd["output_layer.bias"])  # :445
    return torch.complex(out[:, 0:1], out[:, 1:2])


def ncsnpp_time_forward(sd, x, time_cond):
    """NCSNppTime.forward (ncsnpp.py:498-506): (B,1,T) -> (B,1,T)."""
    T = x.shape[-1]
    return net_istft(ncsnpp_forward(sd, net_stft(x), time_cond), T)
