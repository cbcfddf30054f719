"""Batched front-end for the dereverberation test loop.

The reference's `Tester.test_dereverberation` (testing/tester.py:123-163) feeds the sampler ONE utterance at a time:
scale to sigma_data (:135), build the RIR operator and the observation (:143-145), `predict_conditional` (:153), write
wavs.  Every utterance is an independent problem, so utterances of EQUAL length can share a batch without changing any
result (per-utterance GroupNorm statistics, norms, noise streams; see samplers.py).  Utterances of different lengths
cannot be padded into one batch exactly — GroupNorm statistics and the attention span the whole spectrogram — so the
front-end buckets by exact length and runs each bucket in chunks of `max_batch`.

Only tensors in, tensors out: file I/O (soundfile) stays with the caller, as in the reference's `utils/log.py`.
"""
import torch

from .operators import RIROperator


def length_buckets(lengths, max_batch):
    """[(indices)] — indices grouped by equal length (first-seen order), each group split into chunks <= max_batch."""
    groups = {}
    for i, n in enumerate(lengths):
        groups.setdefault(int(n), []).append(i)
    out = []
    for idx in groups.values():
        for s in range(0, len(idx), max_batch):
            out.append(idx[s:s + max_batch])
    return out


class BatchedDereverb:
    """sampler: a `buddy_b200.samplers.EulerHeunSamplerDPS`; `scaling_factor` = tester.posterior_sampling.
    warm_initialization.scaling_factor (sigma_data of the dataset, tester.py:135)."""

    def __init__(self, sampler, max_batch=32, scaling_factor=None):
        self.sampler = sampler
        self.max_batch = int(max_batch)
        ps = sampler.args.tester.posterior_sampling
        self.scaling_factor = float(ps.warm_initialization.scaling_factor if scaling_factor is None else scaling_factor)

    def observe(self, clean, rir):
        """tester.py:134-145 for one utterance: (seg = scaling_factor * x / std(x), y = RIR(seg)); 1-D CUDA tensors."""
        seg = clean.float()
        seg = self.scaling_factor * seg / seg.std()
        op = RIROperator(time_kernel_size=rir.shape[-1])
        op.update_params(rir.float())
        return seg, op.degradation(seg[None])[0]

    def informed(self, ys, rirs):
        """ys: list of 1-D reverberant signals; rirs: list of 1-D RIRs (any lengths).  Returns the list of
        reconstructions in input order — each identical to `predict_conditional(y[None], operator)` run alone."""
        assert len(ys) == len(rirs)
        preds = [None] * len(ys)
        first = getattr(self.sampler, "utterance_offset", 0)
        restore = self.sampler.seed_base
        if restore is None:
            # one run seed for the whole call: utterance i keeps stream seed + first + i whatever bucket it lands in
            self.sampler.seed_base = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        for idx in length_buckets([y.shape[-1] for y in ys], self.max_batch):
            y = torch.stack([ys[i].float() for i in idx])
            m = max(rirs[i].shape[-1] for i in idx)
            h = torch.zeros(len(idx), m, device=y.device)
            for r, i in enumerate(idx):
                h[r, :rirs[i].shape[-1]] = rirs[i].float()      # zero tail: the same convolution
            op = RIROperator(time_kernel_size=m)
            op.update_params(h)
            self.sampler.utterance_ids = [first + i for i in idx]   # noise stream of utterance i = seed_base + first + i
            out = self.sampler.predict_conditional(y, op, shape=tuple(y.shape))
            for r, i in enumerate(idx):
                preds[i] = out[r]
        self.sampler.utterance_ids = None
        self.sampler.seed_base = restore
        return preds
