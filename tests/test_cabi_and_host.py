"""CPU-only checks: the C-ABI library loads without a GPU and exports every symbol include/buddy_b200.h declares;
host-side logic (schedule, sharding with a 2-rank gloo group, state_dict layout, loud failure without CUDA)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from buddy_b200 import build
    lib = ctypes.CDLL(build.build())
    hdr = open(os.path.join(ROOT, "include", "buddy_b200.h")).read()
    names = set(re.findall(r"\b(buddy_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/buddy_b200.h but not exported"
    lib.buddy_last_error.restype = ctypes.c_char_p
    assert lib.buddy_version() >= 100


def test_ctypes_struct_layout_matches_header():
    """sizeof() of the mirrored structs must equal the C compiler's (guards against silent ABI drift)."""
    from buddy_b200 import _capi
    src = r'''
    #include <stdio.h>
    #include "buddy_b200.h"
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(buddy_gemm_desc), sizeof(buddy_gn_desc), sizeof(buddy_gn_bwd_desc), sizeof(buddy_pack_desc)); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(_capi.GemmDesc), ctypes.sizeof(_capi.GnDesc),
                                     ctypes.sizeof(_capi.GnBwdDesc), ctypes.sizeof(_capi.PackDesc)]


def test_no_cpu_fallback():
    from buddy_b200.ncsnpp import NCSNppTime
    m = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 8192), torch.zeros(1))
    from buddy_b200 import ops
    with pytest.raises(AssertionError):
        ops.gn_stats(torch.zeros(1, 4, 4, 128))


def test_unsupported_config_raises():
    from buddy_b200.ncsnpp import NCSNppTime
    with pytest.raises(NotImplementedError):
        NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), fir=True, resblock_type="ddpm")
    with pytest.raises(NotImplementedError):
        NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=64)
    with pytest.raises(NotImplementedError):
        NCSNppTime(stft=dict(n_fft=512, hop_length=128, center=True))


def test_shard_range_partitions():
    from buddy_b200.dist import shard_range
    for total in (1024, 1000, 7):
        for world in (1, 2, 4, 8):
            r = [shard_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from buddy_b200.dist import shard_range, max_over_ranks, gather_utterances, world_info
rank, world, _ = world_info()
dist.init_process_group("gloo", rank=rank, world_size=world)
lo, hi = shard_range(rank, world, 5)
local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
full = gather_utterances(local, 5)
assert torch.equal(full[:, 0], torch.arange(5, dtype=torch.float32)), full
assert max_over_ranks(10.0 + rank) == 10.0 + world - 1
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out


def test_sampler_schedule_host_logic():
    """Schedule/gamma of the product sampler equal the reference fixture — pure host arithmetic, no GPU needed."""
    from buddy_b200.edm import EDM
    from buddy_b200.samplers import EulerHeunSampler
    from oracle import ref_harness as rh
    g = torch.load(os.path.join(ROOT, "tests", "golden", "schedule.pt"), weights_only=False)

    class M(torch.nn.Module):
        pass
    for key in ("informed_35", "blind_60"):
        mode, T = key.split("_")
        s = EulerHeunSampler(M(), EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)),
                             rh.make_args(mode, int(T)))
        t = s.create_schedule()
        assert torch.equal(t, g[key]["t"]) and torch.equal(s.get_gamma(t), g[key]["gamma"])


def test_length_buckets():
    """Front-end bucketing (tester.py:132 loop, batched): equal lengths share a batch, chunks <= max_batch, every index
    exactly once, input order kept inside a bucket."""
    from buddy_b200.tester import length_buckets
    lengths = [100, 200, 100, 100, 300, 200, 100]
    b = length_buckets(lengths, 2)
    assert b == [[0, 2], [3, 6], [1, 5], [4]]
    assert sorted(i for g in b for i in g) == list(range(len(lengths)))
    assert length_buckets([], 4) == [] and length_buckets([5], 4) == [[0]]


def test_factored_stft_matrices_equal_the_dft_matrices():
    """Host side of the FFT-based 1024-point STFT kernels: every matrix the likelihood / blind operator uses is
    mat[2f+c][n] = a[f] * w[n] * (cos, -sin)(2 pi f n / 1024) with the (a, w) the code passes to `ops.FftMat`
    (spectral.LossSTFT, blind.BlindEngine) — checked against the explicit DFT matrices in fp64."""
    import math
    from buddy_b200.spectral import _dft_mats, irfft_weights
    n_fft, bins, win = 1024, 513, 512
    w = torch.hann_window(win, dtype=torch.float64)
    norm = math.sqrt(float((w ** 2).sum()))
    ana, syn = _dft_mats(n_fft, bins, w, win, "cpu")
    n = torch.arange(win, dtype=torch.float64)
    f = torch.arange(bins, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(f, n) / n_fft

    def factored(a):
        m = torch.empty(2 * bins, win, dtype=torch.float64)
        m[0::2] = a[:, None] * torch.cos(ang) * w
        m[1::2] = -a[:, None] * torch.sin(ang) * w
        return m

    ones = torch.ones(bins, dtype=torch.float64)
    aw = irfft_weights(bins, n_fft)
    assert aw[0] == 1 / n_fft and aw[-1] == 1 / n_fft and aw[1] == 2 / n_fft
    for got, want in ((factored(ones), ana.double()), (factored(ones / norm), ana.double() / norm),
                      (factored(aw), syn.double()), (factored(aw * norm), syn.double() * norm)):
        assert (got - want).abs().max().item() < 1e-7 * want.abs().max().item()


def test_package_root_exports():
    """the plug-in classes selectable as `buddy_b200.<Class>` Hydra targets."""
    import importlib
    import buddy_b200
    for name, mod in buddy_b200._EXPORTS.items():
        assert getattr(buddy_b200, name) is getattr(importlib.import_module("buddy_b200." + mod), name)
    with pytest.raises(AttributeError):
        buddy_b200.no_such_thing


def test_netspec_variant_layouts():
    """Module plan / state_dict layout of every NCSN++ graph variant.  The (tensor count, module count) pairs were read
    off the instantiated reference (`NCSNppTime(resblock_type=..., progressive=..., progressive_input=...)`,
    networks/ncsnpp.py:196-274) in the build container, where `param_spec` was compared key by key and shape by shape."""
    from buddy_b200 import netspec
    from buddy_b200.ncsnpp import NCSNppTime
    want = {("biggan", "output_skip", "input_skip"): (271, 36), ("biggan", "output_skip", "residual"): (271, 36),
            ("biggan", "output_skip", "none"): (265, 33), ("biggan", "residual", "input_skip"): (269, 35),
            ("biggan", "residual", "residual"): (269, 35), ("biggan", "residual", "none"): (263, 32),
            ("biggan", "none", "input_skip"): (259, 30), ("biggan", "none", "residual"): (259, 30),
            ("biggan", "none", "none"): (253, 27), ("ddpm", "output_skip", "input_skip"): (211, 36),
            ("ddpm", "output_skip", "residual"): (211, 36), ("ddpm", "output_skip", "none"): (205, 33),
            ("ddpm", "residual", "input_skip"): (209, 35), ("ddpm", "residual", "residual"): (209, 35),
            ("ddpm", "residual", "none"): (203, 32), ("ddpm", "none", "input_skip"): (199, 30),
            ("ddpm", "none", "residual"): (199, 30), ("ddpm", "none", "none"): (193, 27)}
    for v, (n_keys, n_mod) in want.items():
        spec = netspec.param_spec(*v)
        assert (len(spec), netspec.plan(*v)[1]) == (n_keys, n_mod), v
        assert len({k for k, _ in spec}) == n_keys
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2],
                     resblock_type="ddpm", progressive="residual", progressive_input="residual")
    sd = net.state_dict()
    assert [(k, tuple(t.shape)) for k, t in sd.items()] == netspec.param_spec("ddpm", "residual", "residual")
    assert sum(t.numel() for t in NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128,
                                             ch_mult=[1, 2, 2, 2]).state_dict().values()) == 27_736_590


def test_wav_reader_and_paired_set(tmp_path):
    """Input half of the tester front-end: `read_wav` against scipy's reader on every sample format, and
    `PairedWavSet` (datasets/vctk.py:148-226: clean/<spk>/<id>.wav + rir/<spk>/<id>.wav, RIR cropped at its direct
    path and peak-normalised)."""
    import numpy as np
    from scipy.io import wavfile
    from buddy_b200.tester import AsyncWavWriter, PairedWavSet, read_wav
    rs = np.random.RandomState(0)
    x = np.clip(rs.randn(3001) * 0.2, -0.99, 0.99).astype(np.float32)
    cases = dict(f32=x, f64=x.astype(np.float64), i16=(x * 32767).astype(np.int16),
                 i32=(x.astype(np.float64) * 2147483647).astype(np.int32), u8=(x * 127 + 128).astype(np.uint8))
    for name, arr in cases.items():
        p = str(tmp_path / (name + ".wav"))
        wavfile.write(p, 16000, arr)
        y, sr = read_wav(p)
        sr2, z = wavfile.read(p)
        z = {"i16": lambda a: a / 32768.0, "i32": lambda a: a / 2147483648.0,
             "u8": lambda a: (a.astype(np.float32) - 128.0) / 128.0}.get(name, lambda a: a)(z)
        assert sr == sr2 == 16000 and y.dtype == torch.float32 and np.abs(y.numpy() - z).max() < 1e-7, name
    # 24-bit PCM and two channels, written by hand
    v = (x[:100].astype(np.float64) * 8388607).astype(np.int32)
    b = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], 1).astype(np.uint8).tobytes()
    hdr = b"RIFF" + (36 + len(b)).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little") \
        + (1).to_bytes(2, "little") + (2).to_bytes(2, "little") + (8000).to_bytes(4, "little") \
        + (8000 * 6).to_bytes(4, "little") + (6).to_bytes(2, "little") + (24).to_bytes(2, "little") + b"data" \
        + len(b).to_bytes(4, "little")
    (tmp_path / "s24.wav").write_bytes(hdr + b)
    y, sr = read_wav(str(tmp_path / "s24.wav"))
    assert sr == 8000 and y.shape == (50, 2) and np.abs(y.numpy().ravel() - v / 8388608.0).max() < 1e-7
    with pytest.raises(ValueError):
        (tmp_path / "bad.wav").write_bytes(b"not a wave file at all")
        read_wav(str(tmp_path / "bad.wav"))
    # paired set in the reference's directory layout (written with our own writer: float32 files)
    w = AsyncWavWriter(pcm16=False)
    h = np.zeros(500, np.float32)
    h[37], h[60], h[200] = -0.5, 0.2, 0.1
    for spk, uid in (("p1", "p1_001"), ("p1", "p1_002"), ("p2", "p2_001")):
        w.write(torch.from_numpy(x), 16000, uid, str(tmp_path / "set" / "clean" / spk))
        w.write(torch.from_numpy(h), 16000, uid, str(tmp_path / "set" / "rir" / spk))
    w.close()
    ds = PairedWavSet(str(tmp_path / "set"), speakers_test=["p1"])
    assert len(ds) == 2 and ds.filenames == ["p1_001.wav", "p1_002.wav"]
    c, r, name = ds[1]
    assert name == "p1_002.wav" and torch.equal(c, torch.from_numpy(x))
    assert r.shape == (463,) and r[0] == -1.0 and abs(float(r[23]) - 0.4) < 1e-7      # cropped at |h| max, / peak
    assert len(PairedWavSet(str(tmp_path / "set"))) == 3


def test_async_wav_writer_roundtrip(tmp_path):
    """I/O half of the tester front-end (reference utils/log.py:90-110): 16-bit PCM mono files, written off-thread."""
    import wave

    import torch
    from buddy_b200.tester import AsyncWavWriter
    w = AsyncWavWriter(workers=2)
    xs = [torch.sin(torch.arange(4000 + 100 * i) * 0.03) * 0.5 for i in range(5)]
    paths = [w.write(x, 16000, f"utt{i}", str(tmp_path)) for i, x in enumerate(xs)]
    assert w.close() == paths
    for p, x in zip(paths, xs):
        with wave.open(p) as f:
            assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 16000, x.numel())
            got = torch.frombuffer(bytearray(f.readframes(x.numel())), dtype=torch.int16).float() / 32767.0
        assert (got - x).abs().max() < 1e-4


def test_error_behaviour_without_gpu():
    """Same error classes as the reference for the same mistakes (SURVEY.md §8b "Errors"), and loud failures instead of
    a CPU fallback."""
    import pytest
    import torch
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.samplers import EulerHeunSampler, EulerHeunSamplerDPS
    from oracle import ref_harness as rh
    edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))

    class _Net(torch.nn.Module):
        pass
    s = EulerHeunSamplerDPS(_Net(), edm, rh.make_args("informed", 3))
    with pytest.raises(ValueError):                       # EulerHeunSamplerDPS.py:181
        s.predict_unconditional((1, 8192), "cpu")
    with pytest.raises(RuntimeError):                     # CPU observation: no fallback
        s.predict_conditional(torch.zeros(1, 8192), object(), shape=(1, 8192))
    args = rh.make_args("informed", 3)
    args.tester.sampling_params["schedule"] = "song"
    with pytest.raises(NotImplementedError):              # Sampler.py:58-65
        EulerHeunSampler(_Net(), edm, args).create_schedule()
    for bad in (dict(fir=True, resblock_type="ddpm"), dict(fir=True, progressive="residual"), dict(resblock_type="ddpmpp"),
                dict(progressive_combine="cat"), dict(num_res_blocks=2)):
        with pytest.raises(NotImplementedError):
            NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2], **bad)
    with pytest.raises(NotImplementedError):              # other STFT sizes
        NCSNppTime(stft=dict(n_fft=512, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    with pytest.raises(RuntimeError):                     # module still on the CPU
        net(torch.zeros(1, 1, 8192), torch.zeros(1))
    # operator hyper-parameters the kernels do not implement are refused when the operator is bound
    class _Op:
        op_hp = dict(NFFT=2048, win_length=512, hop=128, window="hann")
        params = torch.zeros(100)
    with pytest.raises(NotImplementedError):
        s._validate_operator(_Op(), 1, False)


def test_wpe_oracle_conventions():
    """oracle/wpe.py (restated nara_wpe conventions, parity unpinned): the Blackman/fading STFT pair reconstructs
    perfectly, the frame count is the one the reference's call produces for a 4 s utterance, and WPE removes late
    reverberation from a synthetic exponentially decaying room."""
    import numpy as np
    from oracle import wpe as ow
    rng = np.random.default_rng(0)
    x = rng.standard_normal(65536)
    X = ow.stft(x)
    assert X.shape == (515, 257)
    assert np.abs(ow.istft(X)[:65536] - x).max() < 1e-12
    n = 8192
    s = np.convolve(np.convolve(rng.standard_normal(n), np.ones(8) / 8)[:n], [1, -0.5])[:n]
    h = rng.standard_normal(3000) * np.exp(-np.arange(3000) / 500.0)
    h[0] = 1
    y = np.convolve(s, h)[:n]
    d = ow.wpe_dereverb(y)
    assert d.shape == (n,) and np.linalg.norm(d - s) < 0.7 * np.linalg.norm(y - s)


def test_checkpoint_loader_strategies(tmp_path):
    """buddy_b200.checkpoint.load_checkpoint: the strategy chain of the reference's loaders (tester.py:60-98,
    training_utils.py:5-98) on reference-format checkpoint dictionaries — EMA preferred, strict, non-strict,
    shape-matched, legacy 'ema_weights' — into the drop-in network (same 271 keys as the reference's)."""
    import pytest
    import torch
    from buddy_b200.checkpoint import CheckpointError, load_checkpoint
    from buddy_b200.ncsnpp import NCSNppTime
    from oracle.weights import make_state_dict
    mk = lambda: NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2])
    ema, other = make_state_dict(0), make_state_dict(1)
    path = str(tmp_path / "ckpt.pt")
    torch.save({"it": 190000, "network": other, "ema": ema, "optimizer": {}}, path)
    net = mk()
    info = load_checkpoint(path, net, log=lambda *a: None)
    assert info["strategy"] == 1 and info["source"] == "ema" and info["it"] == 190000 and info["loaded"] == 271
    assert all(torch.equal(v, ema[k]) for k, v in net.state_dict().items())
    # non-strict: one tensor missing, one unexpected
    part = {k: v for k, v in ema.items() if k != "output_layer.bias"}
    part["extra.weight"] = torch.zeros(3)
    info = load_checkpoint({"ema": part}, mk(), log=lambda *a: None)
    assert info["strategy"] == 2 and info["missing"] == ["output_layer.bias"] and info["unexpected"] == ["extra.weight"]
    # shape-matched: a tensor of the wrong shape is skipped, the rest is loaded
    bad = dict(ema)
    bad["output_layer.bias"] = torch.zeros(5)
    net = mk()
    info = load_checkpoint({"network": bad}, net, log=lambda *a: None)
    assert info["strategy"] == 3 and info["loaded"] == 270 and info["source"] == "network"
    assert torch.equal(net.state_dict()["all_modules.3.weight"], ema["all_modules.3.weight"])
    # legacy layout
    info = load_checkpoint({"model": other, "ema_weights": [ema[k] for k in other]}, net, log=lambda *a: None)
    assert info["strategy"] == 4 and torch.equal(net.state_dict()["all_modules.1.weight"], ema["all_modules.1.weight"])
    with pytest.raises(CheckpointError):
        load_checkpoint({"it": 3}, mk(), log=lambda *a: None)


def test_sampler_placeholders_and_function_mirrors_without_gpu():
    """Sampler base-class placeholders / NoSampler (testing/Sampler.py:23-37,74-86) return None; `get_loss`
    (utils/losses.py:17-95) resolves names like the reference; operator hyper-parameters outside the kernels' geometry
    are refused at construction."""
    import torch
    from buddy_b200 import functional as F
    from buddy_b200.edm import EDM
    from buddy_b200.operators import RIROperator
    from buddy_b200.samplers import NoSampler, Sampler
    from oracle import ref_harness as rh
    edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))

    class _Net(torch.nn.Module):
        pass
    for cls in (Sampler, NoSampler):
        s = cls(_Net(), edm, rh.make_args("unconditional", 3))
        assert s.predict() is None and s.predict_unconditional((1, 8), "cpu") is None
        assert s.predict_conditional(None, None) is None and s.step(None, 0, 0, 0) is None
        assert s.T == 3 and s.step_counter == 0 and len(s.create_schedule()) == 4
    assert F.get_loss(rh.AD(name="none")) is None
    ok = rh.AD(name="l2_comp_stft_summean", weight=512, compression_factor=0.667)
    assert callable(F.get_loss(ok)) and callable(F.get_loss(rh.AD(name="hybrid", loss_1=ok, loss_2=ok)))
    for bad in (rh.AD(name="l2_sum"), rh.AD(name="l2_stft_sum"),
                rh.AD(name="l2_comp_stft_sum", compression_factor=0.5, freq_weighting="sqrt")):
        with pytest.raises(NotImplementedError):
            F.get_loss(bad)
    with pytest.raises(AssertionError):                   # losses.py:49: compression factor outside (0, 1]
        F.get_loss(rh.AD(name="l2_comp_stft_sum", compression_factor=1.5))

    class _Op:
        n_fft, win_length, hop_length = 2048, 512, 128
    with pytest.raises(NotImplementedError):
        F.get_loss(ok, operator=_Op())
    with pytest.raises(NotImplementedError):              # reverb.py:29
        RIROperator(rh.AD(NFFT=1024, win_length=512, hop=128, window="hamming"))
    with pytest.raises(NotImplementedError):
        RIROperator(rh.AD(NFFT=2048, win_length=512, hop=128, window="hann"))
    op = RIROperator(rh.op_hp(), time_kernel_size=10)
    with pytest.raises(AssertionError):                   # reverb.py:34 "filter is None"
        op.degradation(torch.zeros(1, 16))
    with pytest.raises(ValueError):                       # reverb.py:61
        op.apply_stft(torch.zeros(1, 1, 16))
    with pytest.raises(ValueError):
        F.fast_apply_RIR(torch.zeros(16), torch.zeros(4))
    with pytest.raises(NotImplementedError):              # the mixed-radix FFT chain exists at the blind operator's size
        F.minimum_phase_version(torch.zeros(4096))
    with pytest.raises(NotImplementedError):
        F.hilbert(torch.zeros(2, 8192))


def test_front_end_restores_the_sampler_after_a_failed_bucket():
    """BatchedDereverb pins `seed_base` / `utterance_ids` for the duration of a call; a bucket that raises (NaN guard,
    bad input) must hand the sampler back as it was."""
    import torch
    from buddy_b200.tester import BatchedDereverb
    from oracle import ref_harness as rh

    class _Sampler:
        args = rh.make_args("informed", 3)
        seed_base, utterance_ids, utterance_offset = None, None, 0
        seen = []

        def predict_conditional(self, y, op, shape=None, blind=False):
            self.seen.append((self.seed_base, list(self.utterance_ids), tuple(y.shape), tuple(op.params.shape)))
            if y.shape[1] == 24:
                raise FloatingPointError("utterance 0 is NaN")
            return y.clone()

    s = _Sampler()
    fe = BatchedDereverb(s, max_batch=2)
    ys = [torch.zeros(16), torch.zeros(16), torch.zeros(16), torch.zeros(24)]
    rirs = [torch.ones(3), torch.ones(5), torch.ones(4), torch.ones(2)]
    with pytest.raises(FloatingPointError):
        fe.informed(ys, rirs)
    assert s.seed_base is None and s.utterance_ids is None
    # buckets by exact length, chunks of max_batch, one run seed for the whole call, RIRs zero-padded per bucket
    assert [(ids, ys_, hs) for _, ids, ys_, hs in s.seen] == [([0, 1], (2, 16), (2, 5)), ([2], (1, 16), (1, 4)),
                                                             ([3], (1, 24), (1, 2))]
    assert len({seed for seed, *_ in s.seen}) == 1 and s.seen[0][0] is not None
    s.seed_base = 77
    out = fe.informed(ys[:3], rirs[:3])
    assert len(out) == 3 and s.seed_base == 77 and s.seen[-1][0] == 77
