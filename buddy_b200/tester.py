"""Batched front-end for the dereverberation test loop.

The reference's `Tester.test_dereverberation` (testing/tester.py:123-163) feeds the sampler ONE utterance at a time:
scale to sigma_data (:135), build the RIR operator and the observation (:143-145), `predict_conditional` (:153), write
wavs.  Every utterance is an independent problem, so utterances of EQUAL length can share a batch without changing any
result (per-utterance GroupNorm statistics, norms, noise streams; see samplers.py).  Utterances of different lengths
cannot be padded into one batch exactly — GroupNorm statistics and the attention span the whole spectrogram — so the
front-end buckets by exact length and runs each bucket in chunks of `max_batch`.

`BatchedDereverb.informed / .blind` take and return tensors; `AsyncWavWriter` is the output half of the I/O
(`utils/log.py:90-110 write_audio_file`, five synchronous wav writes per utterance in the reference loop): device ->
pinned host copies on a side stream and the file writes on worker threads, so that writing utterance i overlaps the
sampling of the next batch.  `read_wav` / `PairedWavSet` are the input half (`datasets/vctk.py:148-226`
`VCTKTestPaired`: clean/<speaker>/<id>.wav paired with rir/<speaker>/<id>.wav, RIR cropped at its direct path and
peak-normalised) with the files decoded on worker threads, and `BatchedDereverb.test_dereverberation` is the loop of
tester.py:123-163 over such a set, writing the reference's directory layout (tester.py:167-203).
"""
import glob
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
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


def read_wav(path):
    """(float32 tensor [frames] or [frames, channels] in [-1, 1), sample rate) of a RIFF/WAVE file: PCM 8/16/24/32-bit
    or IEEE float 32/64, plain or WAVE_FORMAT_EXTENSIBLE (soundfile's `sf.read` on the files the reference uses)."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 12 or raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:           # extensible: the sub-format GUID starts with the real tag
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if tag == 3 and bits in (32, 64):
        x = np.frombuffer(data, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    elif tag == 1 and bits == 16:
        x = np.frombuffer(data, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 32:
        x = (np.frombuffer(data, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == 1 and bits == 24:
        b = np.frombuffer(data[:len(data) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
    elif tag == 1 and bits == 8:
        x = (np.frombuffer(data, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}, {bits} bits")
    x = x[:len(x) // ch * ch]
    t = torch.from_numpy(np.ascontiguousarray(x))
    return (t if ch == 1 else t.view(-1, ch)), int(sr)


class PairedWavSet:
    """datasets/vctk.py:148-226 (`VCTKTestPaired`): every clean/<speaker>/<id>.wav of the test speakers with its
    rir/<speaker>/<id>.wav.  Items are (clean float32 [n], rir float32 [m], file name); the RIR starts at its direct
    path (arg-max of |h|) and is divided by its peak (:213-216).  Files are decoded on `workers` threads ahead of use."""

    def __init__(self, path, speakers_test=None, speakers_discard=(), fs=16000, num_examples=-1, workers=4):
        self.fs = int(fs)
        self.samples, self.rirs = [], []
        for spk in sorted(os.listdir(os.path.join(path, "clean"))):
            if spk in speakers_discard or (speakers_test is not None and spk not in speakers_test):
                continue
            for f in sorted(glob.glob(os.path.join(path, "clean", spk, "*.wav"))):
                self.samples.append(f)
                self.rirs.append(os.path.join(path, "rir", spk, os.path.basename(f)))
        if num_examples > 0:
            assert len(self.samples) >= num_examples, "error in dataloading: not enough examples"
            self.samples, self.rirs = self.samples[:num_examples], self.rirs[:num_examples]
        self.filenames = [os.path.basename(f) for f in self.samples]
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._items = [self._pool.submit(self._load, i) for i in range(len(self.samples))]

    def _load(self, i):
        x, sr = read_wav(self.samples[i])
        h, sr_h = read_wav(self.rirs[i])
        assert sr == self.fs and sr_h == self.fs, "wrong sampling rate"
        assert x.dim() == 1 and h.dim() == 1, "wrong number of channels"
        h = h[int(h.abs().argmax()):]
        return x, h / h.abs().max(), self.filenames[i]

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return self._items[i].result()


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
        try:
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
        finally:                                  # a failed bucket (NaN guard, bad input) must not leave the sampler pinned
            self.sampler.utterance_ids = None
            self.sampler.seed_base = restore
        return preds

    def test_dereverberation(self, test_set, out_dir, mode="informed_dereverberation", blind=False, device="cuda",
                             writer=None):
        """tester.py:123-163 over a whole test set: scale every utterance to sigma_data, reverberate it with its RIR,
        reconstruct (informed: with the true RIR; blind: operator estimated along the way), and write
        <out_dir>/{original, degraded, reconstructed, true_rir[, estimated_rir]}/<name>.wav (tester.py:167-203) through an
        `AsyncWavWriter`.  Returns the list of reconstructed-file paths in set order."""
        sr = self.sampler.args.exp.sample_rate
        writer = AsyncWavWriter() if writer is None else writer
        items = [test_set[i] for i in range(len(test_set))]
        segs, ys, rirs = [], [], []
        for x, h, _ in items:
            seg, y = self.observe(torch.as_tensor(x).to(device), torch.as_tensor(h).to(device))
            segs.append(seg)
            ys.append(y)
            rirs.append(torch.as_tensor(h).to(device))
        sub = lambda k: os.path.join(out_dir, k)
        for (x, h, name), seg, y in zip(items, segs, ys):
            stem = os.path.basename(name)[:-4]
            writer.write(seg, sr, stem, sub("original"))
            writer.write(y, sr, stem, sub("degraded"))
            writer.write(torch.as_tensor(h), sr, stem, sub("true_rir"))
        if blind:
            preds, est = self.blind(ys)
        else:
            preds, est = self.informed(ys, rirs), None
        paths = []
        for k, (_, _, name) in enumerate(items):
            stem = os.path.basename(name)[:-4]
            paths.append(writer.write(preds[k], sr, stem, sub("reconstructed")))
            if blind:
                writer.write(est[k], sr, stem, sub("estimated_rir"))
        writer.close()        # wait for every file of this call (the writer itself stays usable): the paths exist on return
        return paths

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
        try:
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
        finally:
            self.sampler.utterance_ids = None
            self.sampler.seed_base = restore
        return preds, rirs
