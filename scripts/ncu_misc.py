"""Dev tool (GPU): one launch of every non-conv kernel class of the hot path at bench-like shapes, for
`ncu --set full` captures (profiles/*_r02_ncu.txt) and quick CUDA-event timings.

    python scripts/ncu_misc.py [B]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from buddy_b200 import ops
from buddy_b200.blind import BlindEngine
from buddy_b200.spectral import LossSTFT, NetSTFT, RirConv
from buddy_b200.wpe import WpeDereverb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda"
N = 65536
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)


def timed(name, fn, nbytes=None):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:40s} {ms:8.3f} ms" + (f"  {nbytes / ms / 1e6:7.0f} GB/s (algorithmic bytes)" if nbytes else ""), flush=True)


# ---- GroupNorm apply / backward at the dominant shape (full resolution, 128 channels, c8 operands)
H, W, C = 256, 528, 128
x = rn(B, H, W, C)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
sa = ops.gn_stats(x)
a16 = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
a8 = torch.empty(B, H, W, 2 * C, device=dev, dtype=torch.uint8)
timed("gn_apply 256x528 C128 (+e4m3 pair)", lambda: ops.gn_apply(x, sa, gamma, beta, a16, split=2, out8=a8),
      x.numel() * 4 + a16.numel() * 2 + a8.numel())
da, dsk = rn(B, H, W, C), rn(B, H, W, C)
gsum = torch.empty(B, 32, 2, device=dev, dtype=torch.float64)
dx = torch.empty_like(x)
timed("gn_bwd 256x528 C128 (dskip, dx32, g16+g8)",
      lambda: ops.gn_bwd(x, sa, gamma, beta, da, gsum, dskip=dsk, skip_scale=1.0, dxa=dx, g16a=a16, g16_scale=0.7,
                         split=2, g8a=a8), x.numel() * 16 + a16.numel() * 2 + a8.numel())
del x, da, dsk, dx, a16, a8
# ---- network STFT / iSTFT (n_fft 510 DFT-matrix kernels)
st = NetSTFT(dev)
sig = rn(B, N)
spec = st.forward(sig)
timed("dft_analysis (net STFT 510/128)", lambda: st.forward(sig), sig.numel() * 4 + spec.numel() * 4)
timed("dft_synthesis+ola (net iSTFT)", lambda: st.inverse(spec, N), sig.numel() * 4 + spec.numel() * 4)
# ---- likelihood STFT (1024-point shared-memory FFTs) + compressed loss
ls = LossSTFT(dev)
Y = ls.forward(sig)
timed("fft_analysis (loss STFT 1024/128)", lambda: ls.forward(sig), sig.numel() * 4 + Y.numel() * 4)
timed("fft_synthesis+ola (loss STFT adjoint)", lambda: ls.adjoint(Y, N), sig.numel() * 4 + Y.numel() * 4)
Y2 = ls.forward(rn(B, N))
loss = torch.empty(B, device=dev, dtype=torch.float64)
G = torch.empty_like(Y)
timed("comp_loss (loss + gradient)", lambda: ops.comp_loss(Y, Y2, Y.shape[2], 0.667, 512.0, loss, G), Y.numel() * 12)
# ---- informed operator: 2^17-point FFT convolution
h = rn(B, 16000) * torch.exp(-torch.arange(16000, device=dev) / 3000.0)
rc = RirConv(h, N, dev)
timed("fftconv (RIR forward, 2^17)", lambda: rc.forward(sig), sig.numel() * 8)
timed("fftconv (RIR adjoint)", lambda: rc.adjoint(sig), sig.numel() * 8)
# ---- blind operator: sub-band FIR (forward, d/dX, d/dH), filter design chain
be = BlindEngine(N, dev)
be.init_state(B, torch.full((1, 25), 0.4), torch.full((1, 25), 2.0), (torch.rand(B, 513, 100, device=dev) * 6.28 - 3.14),
              torch.zeros(B, 513, 100, dtype=torch.complex64))
be.select(slice(0, B))
Hb = be.update_H()
timed("subband_fir forward", lambda: ops.subband_fir(Y, Hb, torch.empty_like(Y), Nf=100, pre=1, mode=0), Y.numel() * 8)
timed("subband_fir d/dX", lambda: ops.subband_fir(Y, Hb, torch.empty_like(Y), Nf=100, pre=1, mode=1), Y.numel() * 8)
timed("subband_fir d/dH", lambda: ops.subband_fir(Y, Y2, torch.empty_like(Hb), Nf=100, pre=1, mode=2), Y.numel() * 8)
timed("blind update_H (design + min-phase chain)", lambda: be.update_H())
# ---- WPE warm start
wd = WpeDereverb(dev)
Yw = wd.stft(sig)
timed("wpe (50 taps, 5 iterations, 257 bins)", lambda: ops.wpe(Yw, 50, 2, 5))
# ---- upfirdn2d (the reference's native op): FIR x2 upsampling of a channels-last activation
from buddy_b200 import upfirdn2d as bu
xu = rn(B, 128, 264, 128)
k = torch.tensor([[1., 3, 3, 1]], device=dev)
k = (k.t() @ k) / 16.0
timed("upfirdn2d x2 up, [1,3,3,1]", lambda: bu._launch(xu, k, (2, 2), (1, 1), (2, 1, 2, 1)), xu.numel() * 4 * 5)
