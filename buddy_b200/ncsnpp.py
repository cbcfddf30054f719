"""Drop-in replacements for `networks.ncsnpp.NCSNpp` / `NCSNppTime` (reference networks/ncsnpp.py:45-506).

Same constructor keywords (the Hydra config conf/network/ncsnpp.yaml instantiates it by `_target_`), same
`state_dict()` keys, same call signatures:
    NCSNpp.forward(x: complex64 (B,1,F,T), time_cond: (B,)) -> complex64 (B,1,F,T)
    NCSNppTime.forward(x: float32 (B,1,T), time_cond: (B,)) -> float32 (B,1,T)
and differentiable w.r.t. `x` through torch.autograd (the reference DPS sampler calls
`torch.autograd.grad(rec, x)` straight through the network, testing/EulerHeunSamplerDPS.py:65).

All arithmetic runs in the sm_100a kernels of libbuddy_b200.so via `Engine`; there is no CPU/eager fallback —
calling this module with CPU tensors raises.
"""
import math

import torch
import torch.nn as nn

from . import netspec, ops
from .engine import Engine
from .spectral import NetSTFT

_SUPPORTED = dict(nonlinearity='swish', nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=1, attn_resolutions=(0,),
                  resamp_with_conv=True, time_conditional=True, fir=False, skip_rescale=True, resblock_type='biggan',
                  progressive='output_skip', progressive_input='input_skip', progressive_combine='sum',
                  embedding_type='fourier', input_channels=2, spatial_channels=1, dropout=0, centered=True,
                  discriminative=False, image_size=256)


def _default_init(shape, scale, gen=None):
    """DDPM `default_init` (variance_scaling fan_avg uniform, networks/ncsnpp_utils/layers.py:53-91)."""
    scale = 1e-10 if scale == 0 else scale
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    var = scale / ((fan_in + fan_out) / 2)
    return (torch.rand(*shape, generator=gen) * 2.0 - 1.0) * math.sqrt(3 * var)


class _Box(nn.Module):
    """Parameter container giving the reference's dotted state_dict names."""


def _assign(root, dotted, tensor):
    parts = dotted.split(".")
    mod = root
    for name in parts[:-1]:
        if name not in mod._modules:
            mod.add_module(name, _Box())
        mod = mod._modules[name]
    mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class NCSNpp(nn.Module):
    def __init__(self, init_scale=0., fourier_scale=16, fir_kernel=(1, 3, 3, 1), precision=None, **kwargs):
        super().__init__()
        import os
        # Variants of the graph (ncsnpp.py:127-150): resblock_type "biggan" (shipped) | "ddpm" (ResnetBlockDDPMpp blocks,
        # skip through a NIN, separate Downsample / Upsample modules with a 3x3 convolution, layerspp.py:93-216);
        # progressive "output_skip" (shipped) | "residual" | "none"; progressive_input "input_skip" (shipped) |
        # "residual" | "none".  Module indices and state_dict keys follow the reference for every combination.
        variants = dict(resblock_type=netspec.RESBLOCK_TYPES, progressive=netspec.PROGRESSIVE,
                        progressive_input=netspec.PROGRESSIVE_INPUT)
        self.resblock_type = str(kwargs.get("resblock_type", "biggan")).lower()
        self.progressive = str(kwargs.get("progressive", "output_skip")).lower()
        self.progressive_input = str(kwargs.get("progressive_input", "input_skip")).lower()
        for k, v in kwargs.items():
            if k in _SUPPORTED:
                want = _SUPPORTED[k]
                got = tuple(v) if isinstance(want, tuple) else v
                got = got.lower() if isinstance(got, str) else got
                if (k in variants and got in variants[k]) or k == "fir":
                    continue
                if got != want:
                    raise NotImplementedError(
                        f"buddy_b200.NCSNpp implements the shipped BUDDy configuration only: {k}={v!r} (need {want!r})")
        self.variant = (self.resblock_type, self.progressive, self.progressive_input)
        # fir: FIR resampling (upfirdn2d) in the BigGAN blocks and the parameter-free pyramids; the fused FIR convolutions
        # (`up_or_down_sampling.Conv2d`: ddpm Upsample / Downsample, residual pyramids) are not implemented
        self.fir = bool(kwargs.get("fir", False))
        self.fir_kernel = tuple(float(v) for v in fir_kernel)
        if self.fir and (self.resblock_type != "biggan" or "residual" in self.variant[1:]):
            raise NotImplementedError("buddy_b200.NCSNpp: fir=True is implemented for resblock_type='biggan' with "
                                      "progressive / progressive_input in {output_skip, input_skip, none}")
        if self.fir and (len(self.fir_kernel) % 2 or len(self.fir_kernel) < 2):
            raise NotImplementedError("fir_kernel must have an even number of taps")
        gn_modules, scaled_convs = netspec.init_roles(*self.variant)
        # the `mixed` single-pass policy is tuned (and measured) on the shipped progressive graph; the other progressive
        # variants keep the e4m3 corrections on every convolution unless told otherwise
        shipped = self.variant[1:] == ("output_skip", "input_skip") and not self.fir
        self.precision = precision or os.environ.get("BUDDY_PRECISION", "mixed" if shipped else "fp16c8")
        self.time_conditional = True
        self.spatial_channels, self.input_channels = 1, 2
        self.FORCE_STFT_OUT = False
        for key, shape in netspec.param_spec(*self.variant):
            leaf = key.split(".")[-1]
            owner = key.split(".")[-2]
            if key == "all_modules.0.W":
                t = torch.randn(shape) * fourier_scale
            elif "GroupNorm" in key or (owner.isdigit() and int(owner) in gn_modules):
                t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif leaf in ("bias", "b"):
                t = torch.zeros(shape)
            elif key == "output_layer.weight":
                t = (torch.rand(shape) * 2 - 1) * math.sqrt(1.0 / 2)   # nn.Conv2d default (kaiming-uniform, fan_in 2)
            elif leaf == "W":      # NIN (in, out): default_init(scale=0.1), NIN_3: init_scale
                sc = init_scale if owner == "NIN_3" else 0.1
                t = _default_init((shape[1], shape[0]), sc).t().contiguous()
            else:
                sc = 1.0
                if owner == "Conv_1" or (owner.isdigit() and int(owner) in scaled_convs):
                    sc = init_scale
                t = _default_init(shape, sc)
            _assign(self, key, t)
        if "output_layer.bias" in dict(self.named_parameters()):
            with torch.no_grad():
                self.output_layer.bias.uniform_(-math.sqrt(0.5), math.sqrt(0.5))
        self._engine = None
        self._engine_key = None

    # -- engine lifecycle ---------------------------------------------------------------------------
    def engine(self):
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("buddy_b200.NCSNpp runs on CUDA (sm_100a) only; move the module to a GPU "
                               "(there is no CPU fallback)")
        key = (str(dev), self.precision, tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if self._engine is None or key != self._engine_key:
            self._engine = Engine(self.state_dict(), dev, precision=self.precision, resblock_type=self.resblock_type,
                                  progressive=self.progressive, progressive_input=self.progressive_input,
                                  fir=self.fir, fir_kernel=self.fir_kernel)
            self._engine_key = key
        return self._engine

    def forward(self, x, time_cond=None):
        """x complex64 (B,1,F,T) -> complex64 (B,1,F,T)   (reference ncsnpp.py:281-449)."""
        assert time_cond is not None, "the shipped model is time-conditional"
        spec = torch.view_as_real(x[:, 0].contiguous())
        out = _NetFn.apply(self, spec, time_cond.float())
        return torch.view_as_complex(out.contiguous())[:, None]


class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, spec, time_cond):
        eng = module.engine()
        need = spec.requires_grad
        out, saved = eng.forward(spec.detach().contiguous().float(), time_cond.detach(), save=need)
        ctx.eng, ctx.saved = eng, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise RuntimeError("buddy_b200: forward was run without requires_grad on the input")
        dx = ctx.eng.vjp(ctx.saved, dout.contiguous().float())
        ctx.saved = None
        return None, dx, None


class NCSNppTime(NCSNpp):
    """NCSNpp wrapped with STFT / iSTFT (reference ncsnpp.py:455-506)."""

    def __init__(self, stft=None, **kwargs):
        assert stft is not None, "stft must be provided"
        super().__init__(**kwargs)
        n_fft = stft["n_fft"] if isinstance(stft, dict) else stft.n_fft
        hop = stft["hop_length"] if isinstance(stft, dict) else stft.hop_length
        if (n_fft, hop) != (NetSTFT.N_FFT, NetSTFT.HOP):
            raise NotImplementedError(f"buddy_b200.NCSNppTime: stft n_fft={n_fft}, hop={hop} unsupported (need 510/128)")
        self.stft_kwargs = stft
        self._stft = None

    def stft_engine(self):
        dev = next(self.parameters()).device
        if self._stft is None or self._stft.device != dev:
            self._stft = NetSTFT(dev)
        return self._stft

    def stft(self, sig):
        """(B,C=1,T) float32 -> (B,1,256,frames16) complex64 (ncsnpp.py:473-486)."""
        B, C, T = sig.shape
        spec = self.stft_engine().forward(sig.reshape(B * C, T).contiguous().float())
        return torch.view_as_complex(spec).reshape(B, C, spec.shape[1], spec.shape[2])

    def istft(self, spec, length=None):
        B, C, F, Tf = spec.shape
        s = torch.view_as_real(spec.reshape(B * C, F, Tf).contiguous())
        return self.stft_engine().inverse(s, length).reshape(B, C, length)

    def forward(self, x, time_cond=None):
        """x float32 (B,1,T) -> float32 (B,1,T)   (ncsnpp.py:498-506)."""
        B, C, T = x.shape
        return _TimeNetFn.apply(self, x.reshape(B * C, T), time_cond.float()).reshape(B, C, T)


class _TimeNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, time_cond):
        eng, st = module.engine(), module.stft_engine()
        need = x.requires_grad
        xd = x.detach().contiguous().float()
        out_spec, saved = eng.forward(st.forward(xd), time_cond.detach(), save=need)
        ctx.eng, ctx.st, ctx.saved, ctx.n = eng, st, saved, xd.shape[1]
        return st.inverse(out_spec, xd.shape[1])

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise RuntimeError("buddy_b200: forward was run without requires_grad on the input")
        g = dout.contiguous().float()
        dspec = ctx.eng.vjp(ctx.saved, ctx.st.inverse_adjoint(g))   # (the engine normalises the cotangent itself)
        ctx.saved = None
        return None, ctx.st.forward_adjoint(dspec, ctx.n), None
