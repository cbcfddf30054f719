"""Deterministic NCSN++ parameter sets for parity tests (test infrastructure).

Shapes/keys follow the reference's `NCSNppTime.state_dict()` at the shipped config
(conf/network/ncsnpp.yaml; networks/ncsnpp.py:47-274).  The reference's own random init is degenerate
(`init_scale: 0` -> 1e-10 variance on every Conv_1 / NIN_3 / pyramid head, layers.py:88-91), which makes the
network output ~0 and parity tests blind.  `make_state_dict` therefore draws every tensor at
"trained-like" scale: fan-avg uniform with scale 1 for all matrices, GroupNorm affine ~ N(1,0.1)/N(0,0.1),
small random biases.
"""
import math

import torch

NF = 128
CH_MULT = (1, 2, 2, 2)
GROUPS = 32


def _rb(keys, i, cin, cout, resample):
    p = f"all_modules.{i}."
    keys += [(p + "GroupNorm_0.weight", (cin,)), (p + "GroupNorm_0.bias", (cin,)),
             (p + "Conv_0.weight", (cout, cin, 3, 3)), (p + "Conv_0.bias", (cout,)),
             (p + "Dense_0.weight", (cout, 4 * NF)), (p + "Dense_0.bias", (cout,)),
             (p + "GroupNorm_1.weight", (cout,)), (p + "GroupNorm_1.bias", (cout,)),
             (p + "Conv_1.weight", (cout, cout, 3, 3)), (p + "Conv_1.bias", (cout,))]
    if cin != cout or resample:
        keys += [(p + "Conv_2.weight", (cout, cin, 1, 1)), (p + "Conv_2.bias", (cout,))]


def param_spec():
    """[(key, shape)] in the reference's state_dict order (271 entries, 27.74 M parameters)."""
    keys = [("output_layer.weight", (2, 2, 1, 1)), ("output_layer.bias", (2,)),
            ("all_modules.0.W", (NF,)),
            ("all_modules.1.weight", (4 * NF, 2 * NF)), ("all_modules.1.bias", (4 * NF,)),
            ("all_modules.2.weight", (4 * NF, 4 * NF)), ("all_modules.2.bias", (4 * NF,)),
            ("all_modules.3.weight", (NF, 2, 3, 3)), ("all_modules.3.bias", (NF,))]
    i = 4
    c = NF
    hs = [NF]
    for lvl, m in enumerate(CH_MULT):
        _rb(keys, i, c, NF * m, False)
        c = NF * m
        i += 1
        hs.append(c)
        if lvl != len(CH_MULT) - 1:
            _rb(keys, i, c, c, True)
            i += 1
            keys += [(f"all_modules.{i}.Conv_0.weight", (c, 2, 1, 1)), (f"all_modules.{i}.Conv_0.bias", (c,))]
            i += 1
            hs.append(c)
    _rb(keys, i, c, c, False)
    i += 1
    p = f"all_modules.{i}."
    keys += [(p + "GroupNorm_0.weight", (c,)), (p + "GroupNorm_0.bias", (c,))]
    for n in range(4):
        keys += [(p + f"NIN_{n}.W", (c, c)), (p + f"NIN_{n}.b", (c,))]
    i += 1
    _rb(keys, i, c, c, False)
    i += 1
    for lvl in reversed(range(len(CH_MULT))):
        cout = NF * CH_MULT[lvl]
        for _ in range(2):
            _rb(keys, i, c + hs.pop(), cout, False)
            c = cout
            i += 1
        keys += [(f"all_modules.{i}.weight", (c,)), (f"all_modules.{i}.bias", (c,))]
        i += 1
        keys += [(f"all_modules.{i}.weight", (2, c, 3, 3)), (f"all_modules.{i}.bias", (2,))]
        i += 1
        if lvl != 0:
            _rb(keys, i, c, c, True)
            i += 1
    assert not hs and i == 36
    return keys


def make_state_dict(seed=0, dtype=torch.float32, spec=None):
    """spec: [(key, shape)] of another variant (e.g. the reference module's own state_dict layout); default shipped."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in (spec or param_spec()):
        if key == "all_modules.0.W":
            t = torch.randn(shape, generator=g) * 16.0  # GaussianFourierProjection, fourier_scale 16
        elif "GroupNorm" in key or key.split(".")[-2] in ("19", "24", "29", "34"):
            if key.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith(".W") and len(shape) == 2:  # NIN: (in, out)
            lim = math.sqrt(3.0 / ((shape[0] + shape[1]) / 2))
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif len(shape) >= 2:
            rf = 1
            for s in shape[2:]:
                rf *= s
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            lim = math.sqrt(3.0 / ((fan_in + fan_out) / 2))
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        else:  # biases
            t = 0.02 * torch.randn(shape, generator=g)
        sd[key] = t.to(dtype)
    return sd
