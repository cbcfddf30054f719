"""Round-2 additions to tests/golden/ (same rules as oracle/make_golden.py: the UNMODIFIED reference, seeded inputs).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_r2

  * sampler_informed_T2_rescale.pt — informed DPS, order 2, constraint_speech_magnitude.use = True: pins that the
    magnitude constraint is applied after the FIRST evaluation of a Heun step only (EulerHeunSamplerDPS.py:128-129
    vs :136-150), a combination neither shipped config exercises (informed: use = False; blind: order 1).
  * upfirdn2d.pt — outputs of the reference's own `upfirdn2d_native` (op/upfirdn2d.py:157-200).  The module cannot be
    imported (its top level JIT-compiles the CUDA extension), so the function's source is taken from the file as is
    and executed on its own.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.make_golden import NS, randn, save, synth_rir, synth_utterance  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402


UPFIR_CASES = [  # (N, C, H, W, kernel taps, up, down, pad (x0, x1, y0, y1))
    (2, 3, 8, 10, 4, (2, 2), (1, 1), (2, 1, 2, 1)), (1, 2, 16, 12, 4, (1, 1), (2, 2), (1, 1, 1, 1)),
    (1, 1, 7, 9, 3, (3, 2), (2, 3), (0, 2, 1, 0)), (2, 2, 9, 8, 5, (1, 1), (1, 1), (-1, 2, 3, -2)),
    (1, 4, 6, 6, 2, (2, 2), (1, 1), (1, 0, 1, 0)),
]


def upfirdn_golden():
    import ast
    import torch.nn.functional as F
    path = os.path.join(rh.REF_ROOT, "networks", "ncsnpp_utils", "op", "upfirdn2d.py")
    src = open(path).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "upfirdn2d_native"][0]
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    out = []
    for i, (n, c, h, w, taps, up, down, pad) in enumerate(UPFIR_CASES):
        x = randn(800 + i, n, c, h, w)
        k = randn(820 + i, taps, taps)
        out.append({"x_seed": 800 + i, "k_seed": 820 + i, "shape": (n, c, h, w), "taps": taps, "up": up, "down": down,
                    "pad": pad, "out": ns["upfirdn2d_native"](x, k, up[0], up[1], down[0], down[1], *pad)})
    save("upfirdn2d.pt", out)


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
    upfirdn_golden()


if __name__ == "__main__":
    main()
