#!/usr/bin/env python
"""bench.py — sampler steps/s of the BUDDy reverse-diffusion dereverberation hot path on N x B200.

Workload (BASELINE.json configs[1]): batch = 32 synthetic utterances of 4.096 s @ 16 kHz per GPU, informed DPS
(EulerHeunSamplerDPS + RIROperator, conf/tester/informed_dereverberation_DPS.yaml), T = 35, order 2, NCSN++ with
random (non-degenerate) weights, default operand precision ("mixed", DESIGN.md §4: <= 1e-3 vs the fp32 reference).  One bench "step" = one Euler-Heun sampler step over the whole batch = 2 DPS
network evaluations (forward + data-gradient) per utterance.  value = utterance-steps per second, whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling: 32 utt per GPU)

`--impl reference` times the reference's own algorithm on the host CPU cores (the oracle port — the Python
reference cannot travel to the GPU box), B = 1, same config.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_SAMPLES = 65536
T_STEPS = 35
BATCH_PER_GPU = 32
BLIND_MICRO_BATCH = 64        # utterances per network evaluation in the blind configuration (122 GB of activations)
EVALS_PER_UTT = 2 * (T_STEPS - 1) + 1          # 69 (order 2, last step is Euler)
GFLOP_PER_EVAL = 2578.7                        # fwd 1289.3 + VJP 1289.3


class AD(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def informed_args(T):
    sde = AD(sigma_data=0.05, sigma_min=1e-4, sigma_max=0.5, rho=10)
    loss = AD(name="l2_comp_stft_summean", weight=512, frequency_weighting="none", compression_factor=0.667,
              multiple_compression_factors=False)
    return AD(exp=AD(audio_len=65536, sample_rate=16000),   # (normguide divides by sqrt(exp.audio_len) whatever the length)
              tester=AD(sampling_params=AD(same_as_training=False, sde_hp=sde, Schurn=10, Snoise=1, Stmin=0, Stmax=10,
                                           order=2, T=T, schedule="edm"),
                        posterior_sampling=AD(zeta=2.75, rec_loss=loss, normalization_type="grad_norm",
                                              warm_initialization=AD(mode="reverb_scaled", scaling_factor=0.05),
                                              constraint_speech_magnitude=AD(use=False))))


def synth_batch(first, count, N_SAMPLES=N_SAMPLES):
    """Speech surrogate + RIR per utterance: seeds 1000+b / 2000+b, host side, fp32."""
    import numpy as np
    from scipy.signal import lfilter
    s = torch.empty(count, N_SAMPLES)
    h = torch.empty(count, 16000)
    for i in range(count):
        b = first + i
        n = torch.randn(N_SAMPLES, generator=torch.Generator().manual_seed(1000 + b)).numpy()
        lp = torch.from_numpy(lfilter([1.0], [1.0, -0.95], n).astype(np.float32))
        s[i] = 0.05 * lp / lp.std()
        g = torch.Generator().manual_seed(2000 + b)
        t60 = 0.3 + 0.7 * torch.rand(1, generator=g).item()
        r = torch.randn(16000, generator=g) * torch.exp(-6.908 * torch.arange(16000) / (t60 * 16000))
        r[0] = 1.0
        h[i] = r / r.abs().max()
    return s, h


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report nothing rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_time(steps, warmup, threads=None):
    """The reference's own sampler on the host cores: B = 1, one Euler-Heun DPS step per 'step'.

    kind "reference": the UNMODIFIED reference (`testing.EulerHeunSamplerDPS.step` driving `networks.ncsnpp.NCSNppTime`,
    `testing.operators.reverb.RIROperator`, `utils.losses`) imported from oracle/_ref (staged by oracle/stage_ref.py;
    /root/reference in the build container).  kind "port": the oracle restatement, when no reference copy travelled.
    Must run in a process that sees no GPU (the reference's operators pick "cuda" whenever it is available)."""
    from oracle import ref_harness as rh
    from oracle.weights import make_state_dict
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = make_state_dict(0)
    s, h = synth_batch(0, 1)
    g = torch.Generator().manual_seed(3000)
    times = []
    if rh.available() and not torch.cuda.is_available():
        rh.install()
        from testing.EulerHeunSamplerDPS import EulerHeunSamplerDPS
        from testing.operators.reverb import RIROperator
        from utils.losses import get_loss
        args = rh.make_args("informed", T_STEPS)
        net, edm = rh.build_network(sd), rh.build_edm()
        op = RIROperator(args.tester.informed_dereverberation.op_hp, time_kernel_size=h.shape[-1], sample_rate=16000)
        op.update_params(h[0])
        y = op.degradation(s)
        smp = EulerHeunSamplerDPS(net, edm, args)
        smp.operator, smp.y = op, y
        smp.rec_loss = get_loss(args.tester.posterior_sampling.rec_loss, operator=op)
        t = smp.create_schedule()
        gamma = smp.get_gamma(t)
        x = smp.initialize_x((1, N_SAMPLES), "cpu", t)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x, _ = smp.step(x, t[i], t[i + 1], gamma[i], blind=False)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return sum(times) / len(times), threads, "reference"
    from oracle import operators as oop
    from oracle import sampler as osm
    y = oop.fast_apply_rir(s, h[0])
    t = osm.create_schedule(T_STEPS)
    gamma = osm.get_gamma(t, 10)
    x = 0.05 * y / y.std() + t[0] * torch.randn(1, N_SAMPLES, generator=g)
    degrade = lambda v: oop.fast_apply_rir(v, h[0])
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x_hat, t_hat = osm._perturb(x, t[i], gamma[i], torch.randn(1, N_SAMPLES, generator=g))
        d, x_den, x_hat = osm._likelihood(sd, x_hat, t_hat, y, degrade, 2.75, N_SAMPLES, False)
        dt = t[i + 1] - t_hat
        x_p = x_hat + dt * d
        d2, x_den, _ = osm._likelihood(sd, x_p, t[i + 1], y, degrade, 2.75, N_SAMPLES, False)
        x = x_hat + dt * 0.5 * (d + d2)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads, "port"


_SAMPLE = {"reference": "B=1, one Euler-Heun DPS step (2 network fwd+VJP evaluations) per step, fp32: the UNMODIFIED "
                        "reference (testing.EulerHeunSamplerDPS.step, networks.ncsnpp.NCSNppTime, RIROperator; staged "
                        "copy oracle/_ref) on all host cores",
           "port": "B=1, one Euler-Heun DPS step (2 network fwd+VJP evaluations) per step, fp32, oracle port of the "
                   "reference run on all host cores (no reference copy on this box)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads, kind = cpu_reference_step_time(args.steps, args.warmup)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": "sampler_steps_per_sec", "value": val, "unit": "utterance-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(1, "cpu"),
        "utterances_per_sec": val * 2 / EVALS_PER_UTT,
        "cpu_baseline": {"value": val, "unit": "utterance-steps/s", "cores": threads, "kind": kind,
                         "sample": _SAMPLE[kind]},
        "e2e": {"value": val, "unit": "utterance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(steps=1, warmup=0):
    """The cpu_baseline leg of the GPU arm: the reference arm in a child process that sees no GPU."""
    import subprocess
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps),
                          "--warmup", str(warmup)], env=env, capture_output=True, text=True, timeout=900)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("cpu baseline subprocess failed: " + out.stderr[-500:])


def blind_args(T):
    a = informed_args(T)
    sp = a.tester.sampling_params
    sp["Schurn"], sp["order"] = 50, 1
    loss = a.tester.posterior_sampling.rec_loss
    a.tester["posterior_sampling"] = AD(
        zeta=0.5, rec_loss=loss, rec_loss_params=loss,
        RIR_noise_regularization=AD(use=True, crop_sigma_max=0.01, crop_sigma_min=5e-4,
                                    loss=AD(name="l2_comp_stft_summean", weight=2560, compression_factor=0.667)),
        blind_hp=AD(optimizer="adam", lr_op=0.1, beta1=0.9, beta2=0.99, weight_decay=0, op_updates_per_step=10),
        warm_initialization=AD(mode="reverb_scaled", scaling_factor=0.05),
        constraint_speech_magnitude=AD(use=True, speech_scaling=0.05))
    return a


def config_dict(batch, precision):
    return {"workload": "informed EulerHeunSamplerDPS T=35 order 2, synthetic 4.096 s @ 16 kHz (65536 samples), "
                        "NCSN++ 27.7M random-init (non-degenerate) weights",
            "batch_per_gpu": batch, "samples": N_SAMPLES, "T": T_STEPS, "order": 2,
            "evals_per_utterance": EVALS_PER_UTT, "gflop_per_eval": GFLOP_PER_EVAL, "precision": precision,
            "step_definition": "one Euler-Heun sampler step over the batch = 2 network fwd+VJP evaluations per "
                               "utterance", "l2_policy": "inputs_larger_than_L2 (>= 1.3 GB of activations per "
                                                         "utterance per evaluation vs 126 MB L2)"}


# --------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from buddy_b200 import _capi, ops
    from buddy_b200.edm import EDM
    from buddy_b200.ncsnpp import NCSNppTime
    from buddy_b200.operators import RIROperator
    from buddy_b200.samplers import EulerHeunSamplerDPS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    torch.manual_seed(0)
    net = NCSNppTime(stft=dict(n_fft=510, hop_length=128, center=True), nf=128, ch_mult=[1, 2, 2, 2], init_scale=1.0,
                     precision=args.precision)
    with torch.no_grad():   # GroupNorm affine away from the identity so every path carries signal
        for k, p in net.named_parameters():
            if "GroupNorm" in k or k.split(".")[1] in ("19", "24", "29", "34") and p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    net = net.to(dev).eval()
    if world > 1:   # replicate rank 0's weights (the only collective besides the timing barriers)
        for p in net.parameters():
            dist.broadcast(p.data, src=0)

    long_form = args.mode == "long"
    N_SAMPLES = 480000 if long_form else 65536        # run_ours-local: 30 s long-form (BASELINE configs[4]) or 4.096 s
    s_host, h_host = synth_batch(rank * B, B, N_SAMPLES)
    op = RIROperator()
    op.update_params(h_host.to(dev))
    y = op.degradation(s_host.to(dev))
    y_host = y.cpu().pin_memory()

    blind = args.mode == "blind"
    T_run = 60 if blind else T_STEPS
    smp = EulerHeunSamplerDPS(net, EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10)),
                              blind_args(T_run) if blind else informed_args(T_STEPS))
    smp.seed_base = 3000          # noise stream of utterance b = 3000 + global index
    smp.utterance_offset = rank * B
    smp.micro_batch = args.micro_batch
    smp.n_streams = args.streams
    smp.use_graphs = not args.no_graphs
    if blind:
        # reference initialisation (tester.py:149-151): T60 = 0.1 s, weight 2, phases of coherent noise, per utterance
        from buddy_b200.blind import BlindEngine
        be = BlindEngine(N_SAMPLES, dev)
        g = torch.Generator().manual_seed(4000 + rank)
        ph0 = torch.angle(torch.view_as_complex(be.loss_stft.forward(
            torch.randn(B, be.LEN_RIR, generator=g).to(dev))[:, :, 1:101].contiguous()))
        be.init_state(B, torch.full((1, 25), 6.908 / (0.1 * 125)), torch.full((1, 25), 2.0), ph0,
                      torch.zeros(B, 513, 100, dtype=torch.complex64))
        be.select(slice(0, B))
        H0 = torch.view_as_complex(be.update_H())

        class _Op:
            pass
        bop = _Op()
        bop.params = [be.full["decays"].clone(), be.full["weights"].clone()]
        bop.params_phases = [torch.angle(H0)]
        bop.H = H0
        del be
        smp.operator, smp.y = bop, y
        smp._bind_operator(bop, y, True)
    else:
        smp.operator, smp.y = op, y
        smp._bind_operator(op, y, False)
    t = smp.create_schedule()
    gamma = smp.get_gamma(t)
    smp._start_run()                                   # run seed = seed_base (what predict*() does first)
    x = smp.initialize_x((B, N_SAMPLES), dev, t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_i = 0

    def one_step(xc):
        nonlocal step_i
        i = step_i % (T_run - 1)        # regular steps only (the final step of a trajectory is an Euler step)
        step_i += 1
        xn, _ = smp.step(xc, t[i], t[i + 1], gamma[i], blind)
        return xn

    for _ in range(args.warmup):
        x = one_step(x)
    clk = ClockSampler(local)
    barrier()
    clk.start()
    _capi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.KernelTimer() as kt:
        e0.record()
        for _ in range(args.steps):
            x = one_step(x)
        e1.record()
        barrier()
    launches = _capi.launch_count()
    clk.stop_flag = True
    clk.join()
    ms = e0.elapsed_time(e1)
    summ = kt.summary()
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: the public step() call with host buffers, H2D + D2H inside the timed region
    x_host = x.cpu().pin_memory()
    out_host = torch.empty(2, B, N_SAMPLES).pin_memory()
    barrier()
    e0.record()
    for _ in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        smp.y = y_host.to(dev, non_blocking=True)
        smp._Y = smp._loss_stft.forward(smp.y)      # observation spectrum recomputed from the fresh copy
        i = step_i % (T_run - 1)
        step_i += 1
        xn, xden = smp.step(xd, t[i], t[i + 1], gamma[i], blind)
        out_host[0].copy_(xn, non_blocking=True)
        out_host[1].copy_(xden, non_blocking=True)
        torch.cuda.synchronize()
        x_host.copy_(out_host[0])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = tms.item()
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if not blind and not long_form and not args.no_extras:
            extras(args, net, smp, op, y_host, h_host, dev, world, rank, B)     # collective-free, but keeps ranks in step
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_gbs, peak_src = peaks()
    calls, conv_ms, conv_flops = summ.get("conv_gemm", (0, 0.0, 0.0))
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None
    eng = net.engine()
    np_ = (kt.issued_flops / conv_flops) if conv_flops else 1.0   # fp16-pass equivalents issued per algorithmic product
    if eng.mixed:
        dtype_s = (f"f16 tensor-core operands, f32 accumulate: the {len(eng.x1_fwd)} largest forward and {len(eng.x1_convs)} "
                   "largest data-gradient convolutions single-pass, the rest + 2 e4m3 correction passes; f32 "
                   "activations / statistics / sampler")
    elif eng.c8:
        dtype_s = "f16 products + 2 e4m3 correction passes (2 fp16-pass equivalents) / f32 accumulate+activations"
    else:
        dtype_s = f"f16 operands x{eng.np} passes / f32 accumulate+activations"
    # roofline.traffic: DRAM bytes of the dominant launch cannot be measured outside a profiler; the number reported
    # is the `ncu --set full` capture of THIS kernel revision committed under profiles/ (file and revision named), per
    # launch, scaled to this run's utterances per launch; null when no capture of the current revision is present.
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "conv_gemm_r02_ncu_full.json")
    if os.path.exists(tp) and not blind and not long_form:
        try:
            cap = json.load(open(tp))
            mbs = min(B, args.micro_batch) / float(cap["utterances_per_launch"])
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * mbs
            traffic_note = (f"ncu --set full capture {cap['file']} (kernel revision {cap['kernel_rev']}): dominant launch "
                            f"{cap['shape']} at {cap['utterances_per_launch']} utterances per launch, dram read "
                            f"{cap['dram_bytes_read']:.4g} B + write {cap['dram_bytes_write']:.4g} B vs algorithmic "
                            f"{cap['algorithmic_bytes']:.4g} B; scaled x{mbs:g} to this run (not measured in this run)")
        except Exception as e:
            traffic_note = f"profiles/conv_gemm_r02_ncu_full.json unreadable: {e}"
    total_ms_ops = sum(v[1] for v in summ.values())
    line = {
        "metric": "sampler_steps_per_sec", "value": value, "unit": "utterance-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_s,
        "data": "synthetic", "config": config_dict(B, args.precision),
        "utterances_per_sec": (value / 60) if blind else (value * 2 / EVALS_PER_UTT),
        "evals_per_sec": value * (1 if blind else 2),
        "algorithmic_tflops": value * (1 if blind else 2) * GFLOP_PER_EVAL / 1e3,
        "e2e": {"value": e2e_value, "unit": "utterance-steps/s", "h2d_bytes_per_step": 2 * B * N_SAMPLES * 4,
                "d2h_bytes_per_step": 2 * B * N_SAMPLES * 4},
        "gpu_launches": int(launches), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
        "clocks": clk.result(),
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv, all conv/NIN/attention "
                     "launches of the timed region)", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                     "traffic_note": traffic_note, "peak_source": peak_src,
                     "achieved_definition": "algorithmic FLOPs (2*M*N*K of the convolution, split-precision passes "
                                            "NOT counted) / summed CUDA-event duration of the launches",
                     "issued_mma_tflops": (achieved * np_) if achieved else None,
                     "fp16_pass_equivalents_per_product": np_, "launches": calls,
                     "share_of_step": conv_ms / total_ms_ops if total_ms_ops else None},
        "kernel_time_share": {k: round(v[1] / total_ms_ops, 4) for k, v in
                              sorted(summ.items(), key=lambda kv: -kv[1][1])[:(16 if blind else 6)]},
        "roofline_secondary": [
            {"kernel": k, "bound": "hbm", "achieved": summ[k][2] / (summ[k][1] * 1e-3) / 1e9, "peak": peak_gbs,
             "unit": "GB/s", "frac": summ[k][2] / (summ[k][1] * 1e-3) / 1e9 / peak_gbs, "launches": summ[k][0],
             "share_of_step": summ[k][1] / total_ms_ops,
             "achieved_definition": "algorithmic bytes (every input / output tensor of the launch once) / summed "
                                    "CUDA-event duration; gn_bwd is a two-pass kernel and really moves x and da twice"}
            for k in ("gn_bwd", "gn_apply") if k in summ and summ[k][1] > 0],
    }
    if not blind and not long_form and not args.no_extras:
        line["extra"] = extras(args, net, smp, op, y_host, h_host, dev, world, rank, B)
    if long_form:
        line["config"]["workload"] = ("LONG-FORM informed EulerHeunSamplerDPS T=35 order 2, synthetic 30 s @ 16 kHz "
                                      "(480000 samples -> 256 x 3760 spectrogram, attention over 15040 tokens, "
                                      "block-wise RIR convolution), BASELINE configs[4]; whole utterances (exact global GroupNorm), "
                                      "attention over query blocks of 2048 rows with recomputation in the backward pass")
        line["config"]["samples"] = N_SAMPLES
        line["config"]["gflop_per_eval"] = 18760.0
        line["algorithmic_tflops"] = value * 2 * 18760.0 / 1e3
    if blind:
        line["config"]["workload"] = ("blind EulerHeunSamplerDPS T=60 order 1 + 10 operator-Adam iterations per step "
                                      "(BASELINE configs[2]), synthetic 4.096 s @ 16 kHz")
        line["config"]["step_definition"] = "one Euler sampler step over the batch = 1 network fwd+VJP + 10 operator updates"
        line["config"]["T"], line["config"]["order"], line["config"]["evals_per_utterance"] = 60, 1, 60
    if world == 1 and not args.no_cpu_baseline and not blind and not long_form:
        line["cpu_baseline"] = cpu_baseline_subprocess(2, 1)
        line["cpu_baseline"]["sample"] += "; mean of 2 steps after 1 warm-up step (about 10-15 s of CPU work)"
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extras(args, net, smp, op, y_host, h_host, dev, world, rank, B):
    """Extra legs of the default bench line, each on this rank's shard and timed with CUDA events:
       e2e_full   : ONE complete `predict_conditional` (T = 35, order 2: 69 evaluations per utterance) through the
                    public API with host buffers (H2D of y inside, D2H of the result inside);
       b1_latency : the reference's own usage, one utterance at a time (CUDA-graph replay), ms per sampler step;
       blind      : BASELINE configs[2] (batch 128 per GPU, order 1 + 10 operator-Adam iterations per step), 3 steps."""
    from buddy_b200.edm import EDM
    from buddy_b200.operators import RIROperator
    from buddy_b200.samplers import EulerHeunSamplerDPS
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edm = EDM("ve_karras", dict(sigma_data=0.05, sigma_min=1e-5, sigma_max=10, rho=10))
    # ---- full trajectory
    res_host = torch.empty(B, N_SAMPLES).pin_memory()
    torch.cuda.synchronize()
    e0.record()
    yd = y_host.to(dev, non_blocking=True)
    pred = smp.predict_conditional(yd, op, shape=(B, N_SAMPLES))
    res_host.copy_(pred, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    out["e2e_full"] = {"seconds": sec, "utterances": B, "T": T_STEPS, "evals_per_utterance": EVALS_PER_UTT,
                       "utterances_per_sec": B / sec, "utterance_steps_per_sec": B * T_STEPS / sec,
                       "h2d_bytes": B * N_SAMPLES * 4, "d2h_bytes": B * N_SAMPLES * 4,
                       "what": "one complete predict_conditional(y, RIROperator) per GPU, host buffers in / out"}
    del pred, yd
    # ---- B = 1 latency
    s1 = EulerHeunSamplerDPS(net, edm, informed_args(T_STEPS))
    s1.seed_base = 3000
    op1 = RIROperator()
    op1.update_params(h_host[0].to(dev))
    y1 = y_host[:1].to(dev)
    s1.operator, s1.y = op1, y1
    s1._bind_operator(op1, y1, False)
    t = s1.create_schedule()
    gamma = s1.get_gamma(t)
    s1._start_run()
    x = s1.initialize_x((1, N_SAMPLES), dev, t)
    for i in range(3):
        x, _ = s1.step(x, t[i], t[i + 1], gamma[i])
    torch.cuda.synchronize()
    e0.record()
    for i in range(3, 9):
        x, _ = s1.step(x, t[i], t[i + 1], gamma[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    out["b1_latency"] = {"ms_per_step": ms, "ms_per_evaluation": ms / 2, "utterance_steps_per_sec": 1e3 / ms,
                         "what": "B = 1 (the reference's own usage, tester.py:153), CUDA-graph replay of forward / VJP"}
    del s1, x
    # ---- blind leg (configs[2])
    torch.cuda.empty_cache()
    Bb, steps = 128, 3
    from buddy_b200.blind import BlindEngine
    sb, hb = synth_batch(rank * Bb, Bb)
    opb = RIROperator()
    opb.update_params(hb.to(dev))
    yb = opb.degradation(sb.to(dev))
    smb = EulerHeunSamplerDPS(net, edm, blind_args(60))
    # 64 utterances per network evaluation (122 GB of saved activations): the operator-update kernels between the
    # forward and the data-gradient pass are launch / latency bound and fill the machine better (measured +3.7 %)
    free_b, _ = torch.cuda.mem_get_info()
    mb_blind = BLIND_MICRO_BATCH if free_b > 160 * 2 ** 30 else args.micro_batch
    torch.cuda.reset_peak_memory_stats()
    smb.seed_base, smb.utterance_offset, smb.micro_batch = 3000, rank * Bb, mb_blind
    be = BlindEngine(N_SAMPLES, dev)
    g = torch.Generator().manual_seed(4000 + rank)
    ph0 = torch.angle(torch.view_as_complex(be.loss_stft.forward(
        torch.randn(Bb, be.LEN_RIR, generator=g).to(dev))[:, :, 1:101].contiguous()))
    be.init_state(Bb, torch.full((1, 25), 6.908 / (0.1 * 125)), torch.full((1, 25), 2.0), ph0,
                  torch.zeros(Bb, 513, 100, dtype=torch.complex64))
    be.select(slice(0, Bb))
    H0 = torch.view_as_complex(be.update_H())

    class _Op:
        pass
    bop = _Op()
    bop.params = [be.full["decays"].clone(), be.full["weights"].clone()]
    bop.params_phases = [torch.angle(H0)]
    bop.H = H0
    del be
    smb.operator, smb.y = bop, yb
    smb._bind_operator(bop, yb, True)
    tb = smb.create_schedule()
    gb = smb.get_gamma(tb)
    smb._start_run()
    x = smb.initialize_x((Bb, N_SAMPLES), dev, tb)
    x, _ = smb.step(x, tb[0], tb[1], gb[0], True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(1, 1 + steps):
        x, _ = smb.step(x, tb[i], tb[i + 1], gb[i], True)
    e1.record()
    torch.cuda.synchronize()
    msb = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        tms = torch.tensor([msb], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        msb = tms.item()
    out["blind"] = {"value": world * Bb * steps / (msb * 1e-3), "unit": "utterance-steps/s", "batch_per_gpu": Bb,
                    "steps": steps, "warmup": 1, "ms_per_step": msb / steps, "micro_batch": mb_blind,
                    "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
                    "utterances_per_sec": world * Bb * steps / (msb * 1e-3) / 60,
                    "workload": "BASELINE configs[2]/[3]: blind EulerHeunSamplerDPS T=60 order 1 + 10 operator-Adam "
                                "iterations per step, 128 utterances per GPU, 4.096 s @ 16 kHz"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--streams", type=int, default=1, help="micro-batches in flight on separate CUDA streams")
    ap.add_argument("--precision", default=os.environ.get("BUDDY_PRECISION", "mixed"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (full trajectory, B=1 latency, blind)")
    ap.add_argument("--no-graphs", action="store_true", help="plain launches even for batches <= 8 (profiling)")
    ap.add_argument("--mode", default="informed", choices=["informed", "blind", "long"],
                    help="informed = BASELINE configs[1] (default, the bench line); blind = configs[2] (order 1, T=60, "
                         "10 operator-Adam iterations per step); long = configs[4] (30 s utterances, batch 16)")
    args = ap.parse_args()
    if args.mode == "blind":
        if args.batch == BATCH_PER_GPU:
            args.batch = 128          # BASELINE configs[2]
        if args.micro_batch == 32:
            args.micro_batch = BLIND_MICRO_BATCH
    if args.mode == "long":
        if args.batch == BATCH_PER_GPU:
            args.batch = 16
        if args.micro_batch == 32:
            args.micro_batch = 8          # 14 GB per 30 s utterance with query-blocked attention (round 1: 4 at 18 GB)
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""       # the reference's operators take "cuda" whenever it is available
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (buddy_b200 has no CPU fallback; use --impl reference for the "
                             "CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
