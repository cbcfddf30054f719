"""buddy_b200 — B200-native (sm_100a) hot path for BUDDy reverse-diffusion dereverberation.

The plug-in classes are also reachable from the package root (`network._target_=buddy_b200.NCSNppTime`,
`tester.sampler._target_=buddy_b200.EulerHeunSamplerDPS`); they are imported on first use so that `import buddy_b200`
itself stays free of torch / CUDA initialisation."""
__version__ = "0.2.0"

_EXPORTS = {
    "NCSNpp": "ncsnpp", "NCSNppTime": "ncsnpp",
    "Sampler": "samplers", "NoSampler": "samplers", "EulerHeunSampler": "samplers", "EulerHeunSamplerDPS": "samplers",
    "EDM": "edm",
    "Operator": "operators", "RIROperator": "operators", "SubbandFiltering": "operators",
    "BlindSubbandFiltering": "operators",
    "BatchedDereverb": "tester", "AsyncWavWriter": "tester", "PairedWavSet": "tester",
    "load_checkpoint": "checkpoint",
    "fast_apply_RIR": "functional", "get_loss": "functional", "hilbert": "functional",
    "minimum_phase_version": "functional",
}


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'buddy_b200' has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS))
