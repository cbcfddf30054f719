"""Round-2 additions to tests/golden/ (same rules as oracle/make_golden.py: the UNMODIFIED reference, seeded inputs).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_r2

  * sampler_informed_T2_rescale.pt — informed DPS, order 2, constraint_speech_magnitude.use = True: pins that the
    magnitude constraint is applied after the FIRST evaluation of a Heun step only (EulerHeunSamplerDPS.py:128-129
    vs :136-150), a combination neither shipped config exercises (informed: use = False; blind: order 1).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import NS, randn, save, synth_rir, synth_utterance  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    rh.install()
    net = rh.build_network(make_state_dict(0))
    edm = rh.build_edm()
    from testing.EulerHeunSamplerDPS import EulerHeunSamplerDPS
    from testing.operators.reverb import RIROperator
    T = 2
    args = rh.make_args("informed", T, rescale=True)
    s = synth_utterance(71, NS)
    h = synth_rir(72, 2000, 0.5)
    op = RIROperator(args.tester.informed_dereverberation.op_hp, time_kernel_size=h.shape[-1], sample_rate=16000)
    op.update_params(h)
    y = op.degradation(s[None])
    noise = [randn(500 + i, 1, NS) for i in range(T + 1)]
    smp = EulerHeunSamplerDPS(net, edm, args)
    with rh.injected_noise(noise):
        pred = smp.predict_conditional(y, op, shape=(1, NS), blind=False)
    save("sampler_informed_T2_rescale.pt", {"T": T, "n": NS, "noise_seed0": 500, "h": h, "y": y, "pred": pred})


if __name__ == "__main__":
    main()
