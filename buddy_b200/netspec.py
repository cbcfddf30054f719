"""Parameter layout of the reference NCSN++ (`NCSNppTime.state_dict()`, networks/ncsnpp.py:47-274) at the shipped
configuration: 271 tensors, 27.74 M parameters.  Keys and shapes must match exactly so that reference checkpoints
load unchanged (utils/training_utils.py:6-27 loads `state_dict['ema']` into the network)."""

NF = 128
CH_MULT = (1, 2, 2, 2)


def _rb(keys, i, cin, cout, resample, ddpm=False):
    """ResnetBlockBigGANpp (layerspp.py:219-274) or, ddpm, ResnetBlockDDPMpp (:163-216; skip through NIN_0) /
    Downsample / Upsample with a 3x3 convolution (:93-160) in the place of the resampling block."""
    p = f"all_modules.{i}."
    if ddpm and resample:
        keys += [(p + "Conv_0.weight", (cout, cin, 3, 3)), (p + "Conv_0.bias", (cout,))]
        return
    keys += [(p + "GroupNorm_0.weight", (cin,)), (p + "GroupNorm_0.bias", (cin,)),
             (p + "Conv_0.weight", (cout, cin, 3, 3)), (p + "Conv_0.bias", (cout,)),
             (p + "Dense_0.weight", (cout, 4 * NF)), (p + "Dense_0.bias", (cout,)),
             (p + "GroupNorm_1.weight", (cout,)), (p + "GroupNorm_1.bias", (cout,)),
             (p + "Conv_1.weight", (cout, cout, 3, 3)), (p + "Conv_1.bias", (cout,))]
    if ddpm:
        if cin != cout:
            keys += [(p + "NIN_0.W", (cin, cout)), (p + "NIN_0.b", (cout,))]
    elif cin != cout or resample:
        keys += [(p + "Conv_2.weight", (cout, cin, 1, 1)), (p + "Conv_2.bias", (cout,))]


def param_spec(resblock_type="biggan"):
    ddpm = resblock_type == "ddpm"
    keys = [("output_layer.weight", (2, 2, 1, 1)), ("output_layer.bias", (2,)),
            ("all_modules.0.W", (NF,)),
            ("all_modules.1.weight", (4 * NF, 2 * NF)), ("all_modules.1.bias", (4 * NF,)),
            ("all_modules.2.weight", (4 * NF, 4 * NF)), ("all_modules.2.bias", (4 * NF,)),
            ("all_modules.3.weight", (NF, 2, 3, 3)), ("all_modules.3.bias", (NF,))]
    i, c, hs = 4, NF, [NF]
    for lvl, m in enumerate(CH_MULT):
        _rb(keys, i, c, NF * m, False, ddpm)
        c = NF * m
        i += 1
        hs.append(c)
        if lvl != len(CH_MULT) - 1:
            _rb(keys, i, c, c, True, ddpm)
            i += 1
            keys += [(f"all_modules.{i}.Conv_0.weight", (c, 2, 1, 1)), (f"all_modules.{i}.Conv_0.bias", (c,))]
            i += 1
            hs.append(c)
    _rb(keys, i, c, c, False, ddpm)
    i += 1
    p = f"all_modules.{i}."
    keys += [(p + "GroupNorm_0.weight", (c,)), (p + "GroupNorm_0.bias", (c,))]
    for n in range(4):
        keys += [(p + f"NIN_{n}.W", (c, c)), (p + f"NIN_{n}.b", (c,))]
    i += 1
    _rb(keys, i, c, c, False, ddpm)
    i += 1
    for lvl in reversed(range(len(CH_MULT))):
        cout = NF * CH_MULT[lvl]
        for _ in range(2):
            _rb(keys, i, c + hs.pop(), cout, False, ddpm)
            c = cout
            i += 1
        keys += [(f"all_modules.{i}.weight", (c,)), (f"all_modules.{i}.bias", (c,))]
        i += 1
        keys += [(f"all_modules.{i}.weight", (2, c, 3, 3)), (f"all_modules.{i}.bias", (2,))]
        i += 1
        if lvl != 0:
            _rb(keys, i, c, c, True, ddpm)
            i += 1
    assert not hs and i == 36
    return keys
