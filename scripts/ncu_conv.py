"""Dev tool (GPU): launches a fixed list of conv_gemm shapes (for `ncu --set full` captures and quick timing).

    python scripts/ncu_conv.py [B] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from buddy_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = "cuda"
SHAPES = [  # (H, W, Cin, Cout, mode) — the mixed policy's dominant launches first
    (256, 528, 256, 256, "x1"), (256, 528, 384, 128, "x1"), (128, 264, 256, 256, "c8"), (256, 528, 128, 128, "c8"),
    (256, 528, 256, 256, "c8"), (128, 264, 512, 256, "x1"), (256, 528, 256, 256, "x3"),
    (256, 528, 128, 128, "c8", 384), (256, 528, 256, 256, "x1", 256),
]


def e4m3(t):
    return t.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)


if os.environ.get("SHAPESET") == "n128":
    SHAPES = [(256, 528, 128, 128, "c8"), (256, 528, 128, 128, "x1"), (256, 528, 128, 128, "c8", 384),
              (256, 528, 256, 128, "x1"), (256, 528, 256, 256, "x1"), (128, 264, 256, 256, "c8")]

for shp in SHAPES:
    H, W, ci, co, mode = shp[:5]
    cs = shp[5] if len(shp) > 5 else 0
    g = torch.Generator(device=dev).manual_seed(0)
    p = {"c8": 1, "x1": 1, "x3": 3}[mode]
    am = 1 if mode != "x3" else 2
    a = torch.randn(B, H, W, ci * am, device=dev, generator=g).half()
    w = (torch.randn(9, co, ci * p, device=dev, generator=g) * 0.05).half()
    out = torch.empty(B, H, W, co, device=dev)
    stats = torch.zeros(B, co // 4, 2, device=dev, dtype=torch.float64)
    kw = {}
    if mode == "c8":
        kw = dict(a8=e4m3(torch.randn(B, H, W, 2 * ci, device=dev, generator=g)),
                  w8=e4m3(torch.randn(9, co, 2 * ci, device=dev, generator=g)))
    if cs:
        kw["a2"] = torch.randn(B, H, W, cs, device=dev, generator=g).half()
        kw["w2"] = (torch.randn(co, cs, device=dev, generator=g) * 0.05).half()
        if mode == "c8":
            kw["a8_2"] = e4m3(torch.randn(B, H, W, 2 * cs, device=dev, generator=g))
            kw["w8_2"] = e4m3(torch.randn(co, 2 * cs, device=dev, generator=g))
    bias = torch.zeros(co, device=dev)
    if os.environ.get("NOSTATS"):
        stats = None
    if os.environ.get("ONETAP"):
        kw["one_tap_per_stage"] = True
    if os.environ.get("DBG"):
        kw["debug_flags"] = int(os.environ["DBG"])
    if os.environ.get("SINGLE"):
        kw["single_tile"] = True
    if os.environ.get("NOPAIR"):
        kw["no_pairs"] = True
    if os.environ.get("DIRECT"):
        kw["direct_epilogue"] = True
    run = lambda: ops.conv_gemm(a, w, out, taps=9, n_total=co, passes=p, bias=bias, stats=stats, **kw)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 2.0 * B * H * W * co * (ci * 9 + cs)
    pe = {"c8": 2, "x1": 1, "x3": 3}[mode]
    print(f"{H}x{W} {ci}->{co}{"+skip" + str(cs) if cs else ""} {mode}: {ms:.3f} ms  alg {alg / ms / 1e9:.0f} TF/s  issued(fp16-equiv) {alg * pe / ms / 1e9:.0f} TF/s",
          flush=True)
