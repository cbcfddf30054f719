"""Module plan and parameter layout of the reference NCSN++ (`NCSNpp.__init__`, networks/ncsnpp.py:47-274) at
nf = 128, ch_mult = (1, 2, 2, 2), one ResBlock per level, attention at the bottleneck only.

`plan()` lists the modules in the order the reference appends them to `all_modules` (= the order its forward pass runs
them, ncsnpp.py:281-449); `param_spec()` derives the state_dict keys and shapes from it.  Keys and shapes must match
exactly so that reference checkpoints load unchanged (utils/training_utils.py:6-27 loads `state_dict['ema']` into the
network).  Shipped configuration (biggan / output_skip / input_skip): 36 modules, 271 tensors, 27.74 M parameters.
"""

NF = 128
CH_MULT = (1, 2, 2, 2)

RESBLOCK_TYPES = ("biggan", "ddpm")
PROGRESSIVE = ("output_skip", "residual", "none")
PROGRESSIVE_INPUT = ("input_skip", "residual", "none")


def plan(resblock_type="biggan", progressive="output_skip", progressive_input="input_skip"):
    """[(kind, module index, cin, cout, level)], level = resolution the module's convolutions run at (0 = full).

    kinds: rb_d / rb_m / rbcat  ResBlock on hs[-1] (result pushed) / on h / on cat(h, hs.pop())
           rbdown / rbup        BigGAN resampling ResBlock;  down / up: ddpm Downsample / Upsample (3x3 conv)
           combine              input_skip: 1x1 conv of the mean-pooled input pyramid, added to h, pushed
           pyrdown              progressive_input residual: Downsample conv of the pyramid, (pyr + h)/sqrt2, pushed
           push                 progressive_input none: h pushed as it is
           attn                 bottleneck attention
           head / head_add      output_skip: GroupNorm + SiLU + conv3x3(C -> 2) (two modules), first / summed into the
                                nearest-upsampled pyramid;  final: the same after the last level (progressive != output_skip)
           gnconv               progressive residual, coarsest level: GroupNorm + SiLU + conv3x3(C -> C) (two modules)
           pyrup                progressive residual: Upsample conv of the pyramid, h = (pyr + h)/sqrt2
    """
    assert resblock_type in RESBLOCK_TYPES and progressive in PROGRESSIVE and progressive_input in PROGRESSIVE_INPUT
    ddpm = resblock_type == "ddpm"
    ops = [("inconv", 3, 2, NF, 0)]
    i, c, hs, pyr_ch = 4, NF, [NF], 2
    top = len(CH_MULT) - 1
    for lvl, m in enumerate(CH_MULT):
        ops.append(("rb_d", i, c, NF * m, lvl))
        c = NF * m
        i += 1
        hs.append(c)
        if lvl != top:
            ops.append(("down" if ddpm else "rbdown", i, c, c, lvl if ddpm else lvl + 1))
            i += 1
            if progressive_input == "input_skip":
                ops.append(("combine", i, 2, c, lvl + 1))
                i += 1
            elif progressive_input == "residual":
                ops.append(("pyrdown", i, pyr_ch, c, lvl))
                pyr_ch = c
                i += 1
            else:
                ops.append(("push", None, c, c, lvl + 1))
            hs.append(c)
    ops += [("rb_m", i, c, c, top), ("attn", i + 1, c, c, top), ("rb_m", i + 2, c, c, top)]
    i += 3
    pyramid_ch = 0
    for lvl in reversed(range(len(CH_MULT))):
        cout = NF * CH_MULT[lvl]
        for _ in range(2):
            ops.append(("rbcat", i, c + hs.pop(), cout, lvl))
            c = cout
            i += 1
        if progressive == "output_skip":
            ops.append(("head" if lvl == top else "head_add", i, c, 2, lvl))
            i += 2
        elif progressive == "residual":
            if lvl == top:
                ops.append(("gnconv", i, c, c, lvl))
                i += 2
            else:
                ops.append(("pyrup", i, pyramid_ch, c, lvl))
                i += 1
            pyramid_ch = c
        if lvl != 0:
            ops.append(("up" if ddpm else "rbup", i, c, c, lvl if ddpm else lvl - 1))
            i += 1
    assert not hs
    if progressive != "output_skip":
        ops.append(("final", i, c, 2, 0))
        i += 2
    return ops, i


def _rb_keys(keys, i, cin, cout, resample, ddpm):
    """ResnetBlockBigGANpp (layerspp.py:219-274) or ResnetBlockDDPMpp (:163-216; skip through NIN_0)."""
    p = f"all_modules.{i}."
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


def param_spec(resblock_type="biggan", progressive="output_skip", progressive_input="input_skip"):
    """[(key, shape)] in the reference's state_dict order."""
    ddpm = resblock_type == "ddpm"
    keys = [("output_layer.weight", (2, 2, 1, 1)), ("output_layer.bias", (2,)),
            ("all_modules.0.W", (NF,)),
            ("all_modules.1.weight", (4 * NF, 2 * NF)), ("all_modules.1.bias", (4 * NF,)),
            ("all_modules.2.weight", (4 * NF, 4 * NF)), ("all_modules.2.bias", (4 * NF,))]
    ops, n_modules = plan(resblock_type, progressive, progressive_input)
    for kind, i, cin, cout, _ in ops:
        p = f"all_modules.{i}."
        if kind == "inconv":
            keys += [(p + "weight", (cout, cin, 3, 3)), (p + "bias", (cout,))]
        elif kind in ("rb_d", "rb_m", "rbcat", "rbdown", "rbup"):
            _rb_keys(keys, i, cin, cout, kind in ("rbdown", "rbup"), ddpm)
        elif kind in ("down", "up", "pyrdown", "pyrup"):
            keys += [(p + "Conv_0.weight", (cout, cin, 3, 3)), (p + "Conv_0.bias", (cout,))]
        elif kind == "combine":
            keys += [(p + "Conv_0.weight", (cout, cin, 1, 1)), (p + "Conv_0.bias", (cout,))]
        elif kind == "attn":
            keys += [(p + "GroupNorm_0.weight", (cin,)), (p + "GroupNorm_0.bias", (cin,))]
            for n in range(4):
                keys += [(p + f"NIN_{n}.W", (cin, cin)), (p + f"NIN_{n}.b", (cin,))]
        elif kind in ("head", "head_add", "final", "gnconv"):
            keys += [(p + "weight", (cin,)), (p + "bias", (cin,)),
                     (f"all_modules.{i + 1}.weight", (cout, cin, 3, 3)), (f"all_modules.{i + 1}.bias", (cout,))]
    return keys


def init_roles(resblock_type="biggan", progressive="output_skip", progressive_input="input_skip"):
    """(bare GroupNorm module indices, module indices of the convolutions drawn with `init_scale`) — ncsnpp.py:232-271."""
    ops, _ = plan(resblock_type, progressive, progressive_input)
    gn = {i for kind, i, *_ in ops if kind in ("head", "head_add", "final", "gnconv")}
    scaled = {i + 1 for kind, i, *_ in ops if kind in ("head", "head_add", "final")}
    return gn, scaled
