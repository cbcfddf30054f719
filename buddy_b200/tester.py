"""Batched front-end for the dereverberation test loop.

The reference's `Tester.test_dereverberation` (testing/tester.py:123-163) feeds the sampler ONE utterance at a time:
scale to sigma_data (:135), build the RIR operator and the observation (:143-145), `predict_conditional` (:153), write
wavs.  Every utterance is an independent problem, so utterances of EQUAL length can share a batch without changing any
result (per-utterance GroupNorm statistics, norms, noise streams; see samplers.py).  Utterances of different lengths
cannot be padded into one batch exactly — GroupNorm statistics and the attention span the whole spectrogram — so the
front-end buckets by exact length and runs each bucket in chunks of `max_batch`.

`BatchedDereverb.informed / .blind` take and return tensors; `AsyncWavWriter` is the I/O half
(`utils/log.py:90-110 write_audio_file`, five synchronous wav writes per utterance in the reference loop): device ->
pinned host copies on a side stream and the file writes on worker threads, so that writing utterance i overlaps the
sampling of the next batch.
"""
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

from .operators import RIROperator


class AsyncWavWriter:
    """write_audio_file(x, sr, name, path) without stalling the sampler (reference: utils/log.py:90-110 via soundfile,
    synchronous `.cpu().numpy()` + `sf.write`).  Mono, 32-bit float WAV (what soundfile writes for float input is
    16-bit PCM by default; `pcm16=True` selects that).  `close()` waits for everything outstanding."""

    def __init__(self, workers=4, pcm16=True):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pcm16 = pcm16
        self.futures = []
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._lock = threading.Lock()

    @staticmethod
    def _wav_bytes(x, sr, pcm16):
        x = x.flatten()
        if pcm16:
            data = (x.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16).numpy().tobytes()
            fmt, bits = 1, 16
        else:
            data = x.to(torch.float32).numpy().tobytes()
            fmt, bits = 3, 32
        hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack(
            "<IHHIIHH", 16, fmt, 1, sr, sr * bits // 8, bits // 8, bits) + b"data" + struct.pack("<I", len(data))
        return hdr + data

    def write(self, x, sr, name, path="tmp", normalize=False):
        """Returns the file path at once; the tensor may be overwritten by the caller as soon as this returns only if
        it is a CPU tensor — CUDA tensors are copied on a side stream ordered after the caller's current stream."""
        os.makedirs(path, exist_ok=True)
        full = os.path.join(path, name + ".wav")
        x = x.detach()
        if normalize:
            x = x / x.abs().max()
        if x.is_cuda:
            host = torch.empty(x.numel(), dtype=torch.float32).pin_memory()
            ev = torch.cuda.Event()
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                host.copy_(x.flatten().float(), non_blocking=True)
                ev.record(self.stream)
            x.record_stream(self.stream)
        else:
            host, ev = x.flatten().float().clone(), None

        def job():
            if ev is not None:
                ev.synchronize()
            with open(full, "wb") as f:
                f.write(self._wav_bytes(host, int(sr), self.pcm16))
            return full

        with self._lock:
            self.futures.append(self.pool.submit(job))
        return full

    def close(self):
        with self._lock:
            fs, self.futures = self.futures, []
        return [f.result() for f in fs]


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

    def init_blind_operator(self, B, device, generator=None):
        """The state tester.py:147-151 builds per utterance — `BlindSubbandFiltering(op_hp)` then
        `update_H(use_noise=True)`: T60 = 0.1 s decays, weight 2, phases of the STFT of white noise made consistent
        (minimum phase, direct path) — for B utterances at once.  Returns a duck-typed operator object with `params`,
        `params_phases`, `H` batched over utterances (what `EulerHeunSamplerDPS.predict_conditional(blind=True)` reads)."""
        from .blind import BlindEngine
        hp = self.sampler.args.tester.informed_dereverberation.op_hp if hasattr(
            self.sampler.args.tester, "informed_dereverberation") else None
        be = BlindEngine(BlindEngine.LEN_RIR, device, op_hp=hp, sample_rate=self.sampler.args.exp.sample_rate)
        g = lambda k, d: (hp[k] if isinstance(hp, dict) else getattr(hp, k)) if hp is not None else d
        ip = g("init_params", None)
        t60 = float((ip["T60_breakpoints"] if isinstance(ip, dict) else ip.T60_breakpoints)[0]) if ip is not None else 0.1
        wt = float((ip["multiexp_weighting"] if isinstance(ip, dict) else ip.multiexp_weighting)[0]) if ip is not None else 2.0
        noise = torch.randn(B, be.LEN_RIR, generator=generator).to(device)
        ph0 = torch.angle(torch.view_as_complex(be.loss_stft.forward(noise)[:, :, 1:be.NF + 1].contiguous()))
        decay = 6.908 / (t60 * (self.sampler.args.exp.sample_rate / be.HOP))
        be.init_state(B, torch.full((1, 25), decay), torch.full((1, 25), wt), ph0,
                      torch.zeros(B, be.F, be.NF, dtype=torch.complex64))
        be.select(slice(0, B))
        H0 = torch.view_as_complex(be.update_H().contiguous()).clone()

        class _BlindState:
            op_hp = hp
            num_exponentials = 1
        op = _BlindState()
        op.params = [be.full["decays"].clone(), be.full["weights"].clone()]
        op.params_phases = [torch.angle(H0)]
        op.H = H0
        op._engine = be
        return op

    def blind(self, ys, generator=None):
        """ys: list of 1-D reverberant signals.  Returns (reconstructions, estimated time-domain RIRs) in input order —
        tester.py:147-161 (`predict_conditional(..., blind=True)` then `sampler.operator.get_time_RIR()`) for whole
        buckets of equal-length utterances at once."""
        preds, rirs = [None] * len(ys), [None] * len(ys)
        first = getattr(self.sampler, "utterance_offset", 0)
        restore = self.sampler.seed_base
        if restore is None:
            self.sampler.seed_base = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        for idx in length_buckets([y.shape[-1] for y in ys], self.max_batch):
            y = torch.stack([ys[i].float() for i in idx])
            op = self.init_blind_operator(len(idx), y.device, generator)
            self.sampler.utterance_ids = [first + i for i in idx]
            out = self.sampler.predict_conditional(y, op, shape=tuple(y.shape), blind=True)
            be = op._engine
            Hr = torch.view_as_real(op.H_batch.contiguous()).contiguous()
            be.init_state(len(idx), op.params_batch[0], op.params_batch[1], op.params_batch[2], op.H_batch)
            be.select(slice(0, len(idx)))
            rir = be.get_time_RIR(Hr)
            for r, i in enumerate(idx):
                preds[i], rirs[i] = out[r], rir[r]
        self.sampler.utterance_ids = None
        self.sampler.seed_base = restore
        return preds, rirs
