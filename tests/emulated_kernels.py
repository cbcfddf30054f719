"""TEST INFRASTRUCTURE ONLY — plain-torch CPU stand-ins for the spectral / pointwise entry points of the C ABI.

`tests/test_host_composition.py` swaps them into `buddy_b200.ops` (pytest `monkeypatch`) to check, without a GPU, the
HOST-side composition the product wraps around its kernels: frame counts, padding offsets, overlap-add envelopes,
adjoint pairs, the block-wise overlap-add of long RIR convolutions, batch normalisation of the loss factory.  Each
stand-in restates the contract written next to the entry point in `include/buddy_b200.h` / the kernel source (cited per
function); none of them is a product code path: nothing under `buddy_b200/`, `bench.py` or `__graft_entry__.py` imports
this module, and the CUDA kernels themselves are checked against the oracle by the `-m gpu` tests.
"""
import math

import torch


def pad_signal(x, left, total, mode, out, tab=None, scale_b=None):
    """spectral.cu `pad_signal_kernel`: padded[b][j] = tab[j] * scale_b[b] * x[b][src(j - left)]; mode 0 zero outside
    [0, N), mode 1 reflect."""
    B, N = x.shape
    s = torch.arange(total) - left
    if mode == 1:
        s = s.abs()
        s = torch.where(s >= N, 2 * (N - 1) - s, s)
    ok = (s >= 0) & (s < N)
    v = torch.where(ok[None], x[:, s.clamp(0, N - 1)], torch.zeros(()))
    if tab is not None:
        v = v * tab[:total]
    if scale_b is not None:
        v = v * scale_b[:, None]
    out.copy_(v)
    return out


def reflect_fold(dxp, N, L, out, scale_b=None):
    """spectral.cu `reflect_fold_kernel`: adjoint of the reflect pad of width L."""
    v = dxp[:, L:L + N].clone()
    s = torch.arange(1, L + 1)
    v[:, s] += dxp[:, L - s]
    s = torch.arange(N - 1 - L, N - 1)
    v[:, s] += dxp[:, 2 * N - 2 + L - s]
    if scale_b is not None:
        v = v * scale_b[:, None]
    out.copy_(v)
    return out


def _frames(sig, hop, frames, K):
    idx = torch.arange(frames)[:, None] * hop + torch.arange(K)[None]
    return sig[:, idx]                                               # [B, frames, K]


def dft_analysis(sig, mat, hop, frames, Tout, out):
    """out[b][f][t][c] = sum_{n<K} mat[2f+c][n] * sig[b][t*hop + n] for t < frames, zero for frames <= t < Tout."""
    M, K = mat.shape
    o = torch.einsum("btk,mk->bmt", _frames(sig, hop, frames, K).double(), mat.double()).float()
    out.zero_()
    out[:, :, :frames, 0] = o[:, 0::2]
    out[:, :, :frames, 1] = o[:, 1::2]
    return out


def dft_synthesis(S, mat, frames, fr):
    """fr[b][t][n] = sum_{m<M} S[b][m/2][t][m%2] * mat[m][n]   (t < frames, n < K)."""
    M, K = mat.shape
    Sm = torch.stack([S[:, :, :frames, 0], S[:, :, :frames, 1]], 2).reshape(S.shape[0], M, frames)
    fr.copy_(torch.einsum("bmt,mk->btk", Sm.double(), mat.double()).float())
    return fr


def _dense(fm):
    """ops.FftMat: mat[2f+c][n] = a[f] * w[n] * (cos, -sin)(2 pi f n / 1024)."""
    f = torch.arange(fm.a.numel(), dtype=torch.float64)
    n = torch.arange(fm.w.numel(), dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(f, n) / 1024
    mat = torch.empty(2 * fm.a.numel(), fm.w.numel(), dtype=torch.float64)
    mat[0::2] = fm.a.double()[:, None] * torch.cos(ang) * fm.w.double()
    mat[1::2] = -fm.a.double()[:, None] * torch.sin(ang) * fm.w.double()
    return mat


def fft_analysis(sig, fm, hop, frames, Tout, out):
    """out = a[f] * FFT_1024(w * frame)[f]: the same linear map as dft_analysis with the factored matrix."""
    return dft_analysis(sig, _dense(fm), hop, frames, Tout, out)


def fft_synthesis(S, fm, frames, fr):
    """fr = w[n] * Re(sum_f a[f] S[f] e^{+2 pi i f n / 1024}): dft_synthesis with the factored matrix."""
    return dft_synthesis(S, _dense(fm), frames, fr)


def ola_gather(fr, hop, off, n_out, out, tab=None, scale_b=None):
    """spectral.cu `ola_gather_kernel`: out[b][s] = tab[s+off] * scale_b[b] * sum_t fr[b][t][s + off - t*hop]."""
    B, frames, K = fr.shape
    total = (frames - 1) * hop + K
    buf = torch.zeros(B, max(total, off + n_out))
    for t in range(frames):
        buf[:, t * hop:t * hop + K] += fr[:, t]
    if tab is not None:
        assert off + n_out <= tab.numel(), "the kernel would read `tab` out of bounds"
        buf[:, :tab.numel()] *= tab
    v = buf[:, off:off + n_out]
    if scale_b is not None:
        v = v * scale_b[:, None]
    out[:, :n_out].copy_(v)            # `out` may be a wider row (the kernel takes its leading dimension)
    return out


def lincomb3(out, x, ca, y=None, cb=None, z=None, cc=None):
    """out[b] = ca[b] x[b] + cb[b] y[b] + cc[b] z[b]."""
    v = x * ca[:, None]
    if y is not None:
        v = v + y * cb[:, None]
    if z is not None:
        v = v + z * cc[:, None]
    out.copy_(v)
    return out


def row_stats(x, out=None):
    """(sum, sum of squares) per row in fp64."""
    xd = x.double()
    return torch.stack([xd.sum(1), (xd * xd).sum(1)], 1)


def comp_loss(Y, X, frames, compression, weight, loss, grad=None):
    """loss[b] = weight/frames * sum |Yc - Xc|^2, Zc = (|Z| + 1e-8)^c e^{j angle Z}; grad = dloss[b]/dX (re, im)."""
    with torch.enable_grad():
        Yc = torch.view_as_complex(Y.contiguous())
        Xr = X.detach().clone().requires_grad_(True)
        Xc = torch.view_as_complex(Xr)
        comp = lambda Z: (Z.abs() + 1e-8) ** compression * torch.exp(1j * Z.angle())
        per = (weight / frames) * ((comp(Yc) - comp(Xc)).abs() ** 2).sum(dim=(1, 2))
        if grad is not None:
            (g,) = torch.autograd.grad(per.sum(), Xr)
            grad.copy_(g)
    loss.copy_(per.detach().double())
    return loss


def fftconv(x, n_in, log2_n2, tw512, work, H, h_batch_stride, mode, y, n_out):
    """L = 256 * 2^log2_n2 points.  mode 0: work <- spectrum of x[:, :n_in] (reusable as `H`); mode 1:
    y = real(ifft(fft(x) H))[:n_out]; mode 2: the same with conj(H) (the adjoint)."""
    L = 256 << log2_n2
    X = torch.fft.fft(x[:, :n_in].double(), L)
    if mode == 0:
        work.copy_(torch.view_as_real(X.to(torch.complex64)))
        return y
    Hc = torch.view_as_complex(H.contiguous()).to(torch.complex128)
    if h_batch_stride == 0:
        Hc = Hc[:1]
    if mode == 2:
        Hc = Hc.conj()
    y[:, :n_out].copy_(torch.fft.ifft(X * Hc, L).real[:, :n_out].float())   # y may be a wider scratch row
    return y


def fft_mixed(x, in_real, work, out, N1, sign, tw512):
    """Unnormalised 256*N1-point DFT (sign -1) / inverse DFT (sign +1) of real or complex rows."""
    xc = x.double() if in_real else torch.view_as_complex(x.contiguous()).to(torch.complex128)
    N = xc.shape[-1]
    F = torch.fft.fft(xc) if sign < 0 else torch.fft.ifft(xc) * N
    out.copy_(torch.view_as_real(F.to(torch.complex64)))
    return out


def minphase_pw(mode, B, N, T, c0=None, c1=None, r0=None, r1=None, oc=None, or0=None, or1=None, scale_inv_n=False):
    """blind.cu `minphase_pw_kernel`: pointwise stages of minimum_phase_version (0-3) and of its backward (4-7)."""
    invN = 1.0 / N
    if mode == 0:      # m = |Hf|, Lc = (log(m + 1e-8), 0)
        m = torch.view_as_complex(c0).abs()
        or0.copy_(m)
        oc[..., 0] = torch.log(m + 1e-8)
        oc[..., 1] = 0
    elif mode == 1:    # D = C * (2 for k < N/2 else 0) (* 1/N)
        w = torch.zeros(N)
        w[:N // 2] = 2.0 * (invN if scale_inv_n else 1.0)
        oc.copy_(c0 * w[None, :, None])
    elif mode == 2:    # phi = -Im(c)/N, E = m e^{j phi}
        phi = -c0[..., 1] * invN
        or0.copy_(phi)
        oc[..., 0] = r0 * torch.cos(phi)
        oc[..., 1] = r0 * torch.sin(phi)
    elif mode == 3:    # hm = Re/N for k < T, sample 0 := r0[0]
        or0.copy_(c0[:, :T, 0] * invN)
        or0[:, 0] = r0[0]
    elif mode == 4:    # backward start: G_z = dh2[k] for 1 <= k < T else 0
        or0.zero_()
        or0[:, 1:T] = r0[:, 1:T]
    elif mode == 5:    # G_E = raw/N; t = e^{-j phi} G_E; g_m1 = Re t; G_c = (0, -m Im t)      (r0 = m, r1 = phi)
        gx, gy = c0[..., 0] * invN, c0[..., 1] * invN
        cs, sn = torch.cos(r1), torch.sin(r1)
        tr, ti = cs * gx + sn * gy, cs * gy - sn * gx
        or0.copy_(tr)
        oc[..., 0] = 0
        oc[..., 1] = -r0 * ti
    elif mode == 6:    # G_L = Re(raw3); g_m = g_m1 + G_L/(m + 1e-8); G_Hf = g_m * Hf / m   (r0 = m, r1 = g_m1, c1 = Hf)
        gm = r1 + c0[..., 0] / (r0 + 1e-8)
        ok = r0 > 0
        safe = torch.where(ok, r0, torch.ones_like(r0))
        oc[..., 0] = torch.where(ok, gm * c1[..., 0] / safe, torch.zeros_like(gm))
        oc[..., 1] = torch.where(ok, gm * c1[..., 1] / safe, torch.zeros_like(gm))
    elif mode == 7:    # g_u = Re(raw4) -> [B][T]
        or0.copy_(c0[:, :T, 0])
    else:
        raise NotImplementedError(mode)


def _shift(Z, d):
    """Z[..., t + d] with zeros outside the row."""
    T = Z.shape[-1]
    out = torch.zeros_like(Z)
    if d >= 0:
        if d < T:
            out[..., :T - d] = Z[..., d:]
    elif -d < T:
        out[..., -d:] = Z[..., :T + d]
    return out


def subband_fir(a, h_or_dy, out, *, Nf, pre, mode, shared_h=False, accumulate=False):
    """include/buddy_b200.h: mode 0 Y[t] = sum_n H[n] X[t+pre-n]; mode 1 dX[s] = sum_n conj(H[n]) dY[s-pre+n];
    mode 2 dH[n] (+)= sum_t conj(X[t+pre-n]) dY[t].  Rows are [batch][F][T] complex (fp32 pairs)."""
    A = torch.view_as_complex(a.contiguous()).to(torch.complex128)
    Bc = torch.view_as_complex(h_or_dy.contiguous()).to(torch.complex128)
    if mode in (0, 1):
        acc = torch.zeros_like(A)
        for n in range(Nf):
            Hn = Bc[..., n][..., None]
            acc += (Hn * _shift(A, pre - n)) if mode == 0 else (Hn.conj() * _shift(A, n - pre))
        res = acc
    else:
        res = torch.stack([(_shift(A, pre - n).conj() * Bc).sum(-1) for n in range(Nf)], -1)
    r = torch.view_as_real(res.to(torch.complex64))
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def _design(decays, weights, phases, tabs):
    """(decays, weights, phases) -> (A [B][F][Nf], H0 = A e^{j phase} with a zero frame on each side), following
    design_subband_filter / design_filter (subband_filtering.py:224-251) with the host-built tables."""
    B, F, Nf = phases.shape
    n = torch.arange(Nf, dtype=decays.dtype)
    D = weights[:, :, None] * torch.exp(decays)[:, :, None] ** (-n[None, None, :])            # [B, 25, Nf]
    D = torch.cat([torch.zeros(B, 1, Nf, dtype=D.dtype), D, torch.zeros(B, 1, Nf, dtype=D.dtype)], 1)  # 27 knots
    L = torch.log(D + 1e-6)
    k, fr = tabs["kidx"].long(), tabs["frac"].to(D.dtype)
    A = torch.exp(L[:, k] + fr[None, :, None] * (L[:, k + 1] - L[:, k])) + 1e-6                # [B, F, Nf]
    K = tabs["corr"].numel()
    A = torch.cat([A[..., :K] / tabs["corr"].to(D.dtype), A[..., K:]], -1) + tabs["dpmag"].to(D.dtype)
    H = torch.polar(A, phases.to(D.dtype))
    H0 = torch.nn.functional.pad(torch.view_as_real(H), (0, 0, 1, 1))
    return A, H0


def blind_design_fwd(decays, weights, phases, tabs, A, H0):
    a, h0 = _design(decays.double(), weights.double(), phases.double(), tabs)
    A.copy_(a.float())
    H0.copy_(h0.float())


def blind_design_bwd(decays, weights, phases, A, tabs, G, dphases, ddecays, dweights):
    """Gradients of <G, H0> w.r.t. the parameters (the analytic chain of blind.cu, here by autograd)."""
    with torch.enable_grad():
        d, w, p = (t.detach().double().requires_grad_(True) for t in (decays, weights, phases))
        _, h0 = _design(d, w, p, tabs)
        gd, gw, gp = torch.autograd.grad((h0 * G.double()).sum(), (d, w, p))
    ddecays.copy_(gd.float())
    dweights.copy_(gw.float())
    dphases.copy_(gp.float())


def adam_project(p, g, m, v, step, lr, beta1, beta2, eps, dmin, dmax, wmin, wmax):
    """blind.cu `adam_project_kernel`: torch.optim.Adam (bias-corrected) + project_params clamps on the first 25
    (decays) / next 25 (weights) entries of every row; NaN survives the clamp."""
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    m.copy_(m + (g - m) * (1.0 - beta1))
    v.copy_(v * beta2 + (1.0 - beta2) * g * g)
    new = p - (lr / bc1) * (m / (v.sqrt() / math.sqrt(bc2) + eps))
    lo = torch.full_like(new, -float("inf"))
    hi = torch.full_like(new, float("inf"))
    lo[:, :25], hi[:, :25] = dmin, dmax
    lo[:, 25:50], hi[:, 25:50] = wmin, wmax
    p.copy_(torch.where(torch.isnan(new), new, torch.minimum(torch.maximum(new, lo), hi)))


ALL = dict(pad_signal=pad_signal, reflect_fold=reflect_fold, dft_analysis=dft_analysis, dft_synthesis=dft_synthesis,
           fft_analysis=fft_analysis, fft_synthesis=fft_synthesis, ola_gather=ola_gather, lincomb3=lincomb3,
           row_stats=row_stats, comp_loss=comp_loss, fftconv=fftconv, fft_mixed=fft_mixed, minphase_pw=minphase_pw,
           subband_fir=subband_fir, blind_design_fwd=blind_design_fwd, blind_design_bwd=blind_design_bwd,
           adam_project=adam_project)
