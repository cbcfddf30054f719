"""Dev tool (GPU): effective HBM bandwidth of the GroupNorm kernels on the largest layer shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buddy_b200 import ops

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (B, H, W, Ca, Cb, mode, split, raw) in [(4, 256, 528, 256, 0, 0, True, True), (4, 256, 528, 128, 0, 0, True, False),
                                            (4, 256, 528, 256, 128, 0, True, True), (4, 128, 264, 256, 0, 1, True, True),
                                            (4, 256, 528, 128, 0, 2, True, True)]:
    C = Ca + Cb
    xa = torch.randn(B, H, W, Ca, device="cuda")
    xb = torch.randn(B, H, W, Cb, device="cuda") if Cb else None
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    sa = ops.gn_stats(xa); sb = ops.gn_stats(xb) if Cb else None
    Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
    am = 2 if split else 1
    out = torch.empty(B, Ho, Wo, C * am, device="cuda", dtype=torch.float16)
    rawt = torch.empty_like(out) if raw else None
    ms = timeit(lambda: ops.gn_apply(xa, sa, gamma, beta, out, xb=xb, sb=sb, mode=mode, out_raw=rawt, split=split))
    byt = xa.numel() * 4 + (xb.numel() * 4 if Cb else 0) + out.numel() * 2 * (2 if raw else 1)
    print(f"gn_apply  B{B} {H}x{W} C{Ca}+{Cb} mode{mode}: {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s")
    da = torch.randn(B, Ho, Wo, C, device="cuda")
    dsk = torch.randn(B, Ho, Wo, C, device="cuda")
    gsum = torch.empty(B, 32, 2, device="cuda", dtype=torch.float64)
    dxa = torch.empty_like(xa); g16a = torch.empty(B, H, W, Ca * am, device="cuda", dtype=torch.float16)
    dxb = torch.empty_like(xb) if Cb else None
    ms = timeit(lambda: ops.gn_bwd(xa, sa, gamma, beta, da, gsum, xb=xb, sb=sb, mode=mode, dskip=dsk, skip_scale=1.0,
                                   dxa=dxa, dxb=dxb, g16a=g16a, g16_scale=0.7, split=split))
    rd = (xa.numel() + (xb.numel() if Cb else 0)) * 4 * 2 + da.numel() * 4 * 2 + dsk.numel() * 4
    wr = dxa.numel() * 4 + (dxb.numel() * 4 if Cb else 0) + g16a.numel() * 2
    print(f"gn_bwd    B{B} {H}x{W} C{Ca}+{Cb} mode{mode}: {ms:.3f} ms  {(rd + wr) / ms / 1e6:.0f} GB/s")
