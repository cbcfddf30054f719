"""Checkpoint loading for the drop-in network (reference testing/tester.py:60-98 `load_checkpoint` /
`load_checkpoint_legacy`, utils/training_utils.py:5-98 `load_state_dict`).

The published BUDDy checkpoints are dictionaries {'it', 'network', 'ema', 'optimizer', ...}; the tester loads the EMA
weights into the network (`tr_utils.load_state_dict(state_dict, ema=self.network)`).  Because
`buddy_b200.ncsnpp.NCSNppTime` has the reference's 271 state_dict keys, the reference's own loaders work on it
unchanged; this module is the same strategy chain for use without the reference on sys.path:
  1. the chosen weight set ('ema' by default, else 'network' / 'model'), strict;
  2. the same, non-strict (missing / unexpected keys reported);
  3. only the tensors whose name AND shape match;
  4. legacy layout: keys of 'model' zipped with the list 'ema_weights' (all entries, or the trainable ones only).
The packed tensor-core operands are rebuilt lazily by `NCSNpp.engine()` at the next forward (parameter versions change).
"""
import torch


class CheckpointError(RuntimeError):
    pass


def _weights(ckpt, prefer):
    for key in prefer:
        if key in ckpt and isinstance(ckpt[key], dict) and ckpt[key]:
            return key, ckpt[key]
    return None, None


def load_checkpoint(ckpt, network, device="cpu", prefer=("ema", "network", "model"), log=print):
    """ckpt: path or already-loaded dictionary.  Returns {'it': iteration, 'source': key used, 'strategy': 1..4,
    'loaded': number of tensors, 'missing': [...], 'unexpected': [...]}; raises CheckpointError if nothing fits."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location=device, weights_only=False)
    info = {"it": int(ckpt.get("it", 0)) if not torch.is_tensor(ckpt.get("it", 0)) else int(ckpt["it"].item()),
            "missing": [], "unexpected": []}
    own = network.state_dict()
    src, sd = _weights(ckpt, prefer)
    legacy = "ema_weights" in ckpt and isinstance(ckpt.get("model"), dict)
    if legacy and src != "ema" and prefer and prefer[0] == "ema":
        sd = None            # averaged weights exist only in the legacy layout: they win over the raw 'model'
    if sd is not None:
        info["source"] = src
        try:
            network.load_state_dict(sd, strict=True)
            info.update(strategy=1, loaded=len(own))
            return info
        except RuntimeError as e:
            log(f"checkpoint['{src}'] does not load strictly: {str(e).splitlines()[0]}")
        same_shape = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
        if len(same_shape) == len([k for k in sd if k in own]) and same_shape:
            res = network.load_state_dict(sd, strict=False)
            info.update(strategy=2, loaded=len(same_shape), missing=list(res.missing_keys),
                        unexpected=list(res.unexpected_keys))
            return info
        if same_shape:
            merged = dict(own)
            merged.update(same_shape)
            network.load_state_dict(merged, strict=True)
            info.update(strategy=3, loaded=len(same_shape), missing=[k for k in own if k not in same_shape],
                        unexpected=[k for k in sd if k not in same_shape])
            return info
    if legacy:
        keys, ema = list(ckpt["model"].keys()), list(ckpt["ema_weights"])
        if len(ema) == len(keys):
            sd = dict(zip(keys, ema))
        else:   # only the trainable tensors were averaged: the others come from 'model'
            sd, i = {}, 0
            for k, v in ckpt["model"].items():
                if getattr(v, "requires_grad", False):
                    sd[k] = ema[i]
                    i += 1
                else:
                    sd[k] = v
        network.load_state_dict(sd, strict=True)
        info.update(source="ema_weights", strategy=4, loaded=len(sd))
        return info
    raise CheckpointError("no loadable weights in the checkpoint (keys: %s)" % sorted(ckpt.keys()))
