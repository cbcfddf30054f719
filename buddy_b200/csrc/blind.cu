// Blind reverb-operator kernels (batched over utterances; every utterance owns its filter, parameters and Adam
// state — SURVEY.md App. C2):
//   * per-bin causal complex FIR  Y[f,t] = sum_n H[f,n] X[f,t+1-n]  and its gradients w.r.t. X and H
//     (SubbandFiltering.subband_filtering, testing/operators/subband_filtering.py:67-74)
//   * parametric filter design: 25 exponential-decay bands -> log -> piecewise-linear interpolation to 513 bins ->
//     exp -> OLA correction -> + direct path -> * exp(j phase)   and its backward
//     (BlindSubbandFiltering.design_subband_filter/correct_OLA/design_filter/update_H, :212-251,281-282)
//   * minimum-phase projection (utils/reverb_utils.py:3-23) pointwise stages, around a mixed-radix
//     (101 x 256 = 25 856-point) FFT written as a direct 101-point DFT + shared-memory radix-2 FFT
//   * Adam (torch.optim.Adam semantics) + parameter projection (project_params, :298-331)
#include <atomic>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {
extern std::atomic<long long> g_launches;
#define LAUNCH_END(name)                              \
  g_launches.fetch_add(1, std::memory_order_relaxed); \
  BUDDY_CHECK_LAUNCH(name);                           \
  return 0;
#define STREAM static_cast<cudaStream_t>(stream)

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // conj(a) * b
  return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

// ------------------------------------------------------------------------------------------------ FIR
// one block per (bin f, utterance b).  X: [B][F][Tx] complex, H: [B][F][Nf] complex (h_bs = 0: shared), Y: [B][F][Tx]
// mode 0: Y[t]  = sum_n H[n] X[t+pre-n]
// mode 1: dX[s] = sum_n conj(H[n]) dY[s-pre+n]
// mode 2: dH[n] (+)= sum_t conj(X[t+pre-n]) dY[t]
constexpr int kFirMaxT = 544, kFirMaxN = 128;
constexpr int kFirR = 3;   // outputs (modes 0, 1) / taps (mode 2) per thread or warp: every loaded value feeds kFirR MACs
// The signal row lives in shared memory with kFirMaxN zeros on both sides, so no tap needs a bounds check:
// sa[kFirMaxN + s] = a[s] for 0 <= s < Tx, 0 elsewhere.
__global__ void __launch_bounds__(192)
subband_fir_kernel(const float2* __restrict__ A, const float2* __restrict__ Hh, long long h_bs, float2* __restrict__ O,
                   int F, int Tx, int Nf, int pre, int mode, int accumulate) {
  __shared__ float2 sa[kFirMaxT + 2 * kFirMaxN + kFirR];
  __shared__ float2 sh[kFirMaxN + kFirR];
  __shared__ float2 sg[kFirMaxT];
  const int f = blockIdx.x, b = blockIdx.y;
  const long long row = (static_cast<long long>(b) * F + f);
  const float2* a = A + row * Tx;
  for (int i = threadIdx.x; i < kFirMaxT + 2 * kFirMaxN + kFirR; i += blockDim.x) {
    const int s0 = i - kFirMaxN;
    sa[i] = (s0 >= 0 && s0 < Tx) ? a[s0] : make_float2(0.f, 0.f);
  }
  if (mode != 2) {
    const float2* h = Hh + b * h_bs + static_cast<long long>(f) * Nf;
    for (int i = threadIdx.x; i < kFirMaxN + kFirR; i += blockDim.x) sh[i] = i < Nf ? h[i] : make_float2(0.f, 0.f);
  } else {
    const float2* g = Hh + row * Tx;  // in mode 2 the second operand is dY
    for (int i = threadIdx.x; i < Tx; i += blockDim.x) sg[i] = g[i];
  }
  __syncthreads();
  const float2* sa0 = sa + kFirMaxN;   // sa0[s] valid for -kFirMaxN <= s < Tx + kFirMaxN
  if (mode == 0) {
    // y[t] = sum_n h[n] x[t + pre - n]: thread = kFirR consecutive outputs, sliding window over x
    float2* o = O + row * Tx;
    for (int t0 = threadIdx.x * kFirR; t0 < Tx; t0 += blockDim.x * kFirR) {
      float2 acc[kFirR], win[kFirR];
#pragma unroll
      for (int r = 0; r < kFirR; ++r) {
        acc[r] = make_float2(0.f, 0.f);
        win[r] = sa0[t0 + pre + r];            // x[(t0 + r) + pre - 0]
      }
      for (int n = 0; n < Nf; ++n) {
        const float2 h = sh[n];
#pragma unroll
        for (int r = 0; r < kFirR; ++r) {
          acc[r].x = fmaf(h.x, win[r].x, fmaf(-h.y, win[r].y, acc[r].x));
          acc[r].y = fmaf(h.x, win[r].y, fmaf(h.y, win[r].x, acc[r].y));
        }
#pragma unroll
        for (int r = kFirR - 1; r > 0; --r) win[r] = win[r - 1];
        win[0] = sa0[t0 + pre - n - 1];        // the next tap reads one sample earlier
      }
#pragma unroll
      for (int r = 0; r < kFirR; ++r)
        if (t0 + r < Tx) o[t0 + r] = acc[r];
    }
  } else if (mode == 1) {
    // dx[s] = sum_n conj(h[n]) dy[s - pre + n]
    float2* o = O + row * Tx;
    for (int s0 = threadIdx.x * kFirR; s0 < Tx; s0 += blockDim.x * kFirR) {
      float2 acc[kFirR], win[kFirR];
#pragma unroll
      for (int r = 0; r < kFirR; ++r) {
        acc[r] = make_float2(0.f, 0.f);
        win[r] = sa0[s0 - pre + r];
      }
      for (int n = 0; n < Nf; ++n) {
        const float2 h = sh[n];
#pragma unroll
        for (int r = 0; r < kFirR; ++r) {       // conj(h) * v
          acc[r].x = fmaf(h.x, win[r].x, fmaf(h.y, win[r].y, acc[r].x));
          acc[r].y = fmaf(h.x, win[r].y, fmaf(-h.y, win[r].x, acc[r].y));
        }
#pragma unroll
        for (int r = 0; r < kFirR - 1; ++r) win[r] = win[r + 1];
        win[kFirR - 1] = sa0[s0 - pre + n + kFirR];
      }
#pragma unroll
      for (int r = 0; r < kFirR; ++r)
        if (s0 + r < Tx) o[s0 + r] = acc[r];
    }
  } else {
    // dH[n] = sum_t conj(x[t + pre - n]) dy[t]: warp = kFirR consecutive taps, lanes stride over t
    float2* o = O + row * Nf;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n0 = warp * kFirR; n0 < Nf; n0 += (blockDim.x >> 5) * kFirR) {
      float2 acc[kFirR];
#pragma unroll
      for (int r = 0; r < kFirR; ++r) acc[r] = make_float2(0.f, 0.f);
      for (int t = lane; t < Tx; t += 32) {
        const float2 g = sg[t];
#pragma unroll
        for (int r = 0; r < kFirR; ++r) {       // conj(x) * g
          const float2 x = sa0[t + pre - n0 - r];
          acc[r].x = fmaf(x.x, g.x, fmaf(x.y, g.y, acc[r].x));
          acc[r].y = fmaf(x.x, g.y, fmaf(-x.y, g.x, acc[r].y));
        }
      }
#pragma unroll
      for (int r = 0; r < kFirR; ++r) {
        float ax = warp_sum(acc[r].x), ay = warp_sum(acc[r].y);
        if (lane == 0 && n0 + r < Nf) {
          if (accumulate) {
            ax += o[n0 + r].x;
            ay += o[n0 + r].y;
          }
          o[n0 + r] = make_float2(ax, ay);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ filter design
struct DesignTabs {
  const int* kidx;     // [F] lower knot index (0..25)
  const float* frac;   // [F]
  const float* corr;   // [3] OLA correction divisors for the first 3 frames
  const float* dpmag;  // [F][Nf] direct-path magnitude correction
};
constexpr int kBands = 27;  // EQ knots incl. the two fixed extremes

// grid (Nf, B), block >= F threads (strided).  Writes A [B][F][Nf] and H0 [B][F][Nf+2] complex (zero frame each side)
__global__ void design_fwd_kernel(const float* __restrict__ decays, const float* __restrict__ weights,
                                  const float* __restrict__ phases, DesignTabs tb, int F, int Nf,
                                  float* __restrict__ A, float2* __restrict__ H0) {
  __shared__ float sL[kBands];
  const int n = blockIdx.x, b = blockIdx.y;
  if (threadIdx.x < kBands) {
    const int e = threadIdx.x;
    float D = 0.f;
    if (e >= 1 && e <= kBands - 2)
      D = weights[b * (kBands - 2) + e - 1] * powf(expf(decays[b * (kBands - 2) + e - 1]), -static_cast<float>(n));
    sL[e] = logf(D + 1e-6f);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const int k = tb.kidx[f];
    const float lf = sL[k] + tb.frac[f] * (sL[k + 1] - sL[k]);
    float a = expf(lf) + 1e-6f;
    if (n < 3) a /= tb.corr[n];
    a += tb.dpmag[f * Nf + n];
    const long long i = (static_cast<long long>(b) * F + f);
    A[i * Nf + n] = a;
    float sn, cs;
    sincosf(phases[i * Nf + n], &sn, &cs);
    H0[i * (Nf + 2) + n + 1] = make_float2(a * cs, a * sn);
    if (n == 0) {
      H0[i * (Nf + 2)] = make_float2(0.f, 0.f);
      H0[i * (Nf + 2) + Nf + 1] = make_float2(0.f, 0.f);
    }
  }
}
// backward: G = dL/dH0 [B][F][Nf+2] -> dphases [B][F][Nf] and per-tap partials of ddecays/dweights
// part[b][2][25][Nf]; design_bwd_reduce_kernel sums the taps in a fixed order (no atomics anywhere: the operator
// gradients feed Adam, whose sign-like first steps amplify any run-to-run difference).
constexpr int kMaxF = 1024;
__global__ void design_bwd_kernel(const float* __restrict__ decays, const float* __restrict__ weights,
                                  const float* __restrict__ phases, const float* __restrict__ A, DesignTabs tb,
                                  const float2* __restrict__ G, int F, int Nf, float* __restrict__ dphases,
                                  float* __restrict__ part) {
  __shared__ float sL[kBands], sD[kBands], sdL[kBands];
  __shared__ float s_dlf[kMaxF];
  const int n = blockIdx.x, b = blockIdx.y;
  if (threadIdx.x < kBands) {
    const int e = threadIdx.x;
    float D = 0.f;
    if (e >= 1 && e <= kBands - 2)
      D = weights[b * (kBands - 2) + e - 1] * powf(expf(decays[b * (kBands - 2) + e - 1]), -static_cast<float>(n));
    sD[e] = D;
    sL[e] = logf(D + 1e-6f);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const long long i = (static_cast<long long>(b) * F + f);
    const float2 g = G[i * (Nf + 2) + n + 1];
    float sn, cs;
    sincosf(phases[i * Nf + n], &sn, &cs);
    const float tr = cs * g.x + sn * g.y;   // Re(e^{-j phi} G)
    const float ti = cs * g.y - sn * g.x;   // Im(e^{-j phi} G)
    dphases[i * Nf + n] = A[i * Nf + n] * ti;
    float dA = tr;
    if (n < 3) dA /= tb.corr[n];
    const int k = tb.kidx[f];
    const float lf = sL[k] + tb.frac[f] * (sL[k + 1] - sL[k]);
    s_dlf[f] = dA * expf(lf);
  }
  __syncthreads();
  if (threadIdx.x < kBands) {
    // knot e collects (1-frac) of the bins in segment e and frac of the bins in segment e-1, in bin order
    const int e = threadIdx.x;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) {
      const int k = tb.kidx[f];
      if (k == e) acc += (1.f - tb.frac[f]) * s_dlf[f];
      else if (k + 1 == e) acc += tb.frac[f] * s_dlf[f];
    }
    sdL[e] = acc;
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x <= kBands - 2) {
    const int e = threadIdx.x;
    const float dD = sdL[e] / (sD[e] + 1e-6f);
    const float w = weights[b * (kBands - 2) + e - 1];
    const float ex = (w != 0.f) ? sD[e] / w : powf(expf(decays[b * (kBands - 2) + e - 1]), -static_cast<float>(n));
    float* pb = part + static_cast<long long>(b) * 2 * (kBands - 2) * Nf;
    pb[(e - 1) * Nf + n] = dD * ex;                                                   // d weights
    pb[((kBands - 2) + e - 1) * Nf + n] = dD * sD[e] * (-static_cast<float>(n));      // d decays
  }
}
// ddecays / dweights [B][25] = sum over the Nf taps of the partials, tap order
__global__ void design_bwd_reduce_kernel(const float* __restrict__ part, int Nf, float* __restrict__ ddecays,
                                         float* __restrict__ dweights) {
  const int b = blockIdx.x, e = threadIdx.x;
  if (e >= 2 * (kBands - 2)) return;
  const float* p = part + (static_cast<long long>(b) * 2 * (kBands - 2) + e) * Nf;
  float acc = 0.f;
  for (int n = 0; n < Nf; ++n) acc += p[n];
  if (e < kBands - 2) dweights[b * (kBands - 2) + e] = acc;
  else ddecays[b * (kBands - 2) + e - (kBands - 2)] = acc;
}

// ------------------------------------------------------------------------------------------------ mixed-radix FFT
// N = N1 * 256, N1 <= 128 (25 856 = 101 * 256).  sign = -1 forward, +1 inverse (unnormalised).
// stage A: for 8 adjacent columns n2: Y[k1][n2] = W_N^(sign n2 k1) * sum_n1 W_N1^(sign n1 k1) x[n1*256 + n2]
__global__ void __launch_bounds__(256)
fftmix_cols_kernel(const float* __restrict__ in, int in_real, int N1, float sign, float2* __restrict__ Y) {
  __shared__ float2 s[128 * 8];
  __shared__ float2 tw[128];
  const int b = blockIdx.y, c0 = blockIdx.x * 8;
  const long long N = static_cast<long long>(N1) * 256;
  for (int i = threadIdx.x; i < N1 * 8; i += blockDim.x) {
    const int n1 = i >> 3, c = i & 7;
    const long long idx = b * N + static_cast<long long>(n1) * 256 + c0 + c;
    s[i] = in_real ? make_float2(in[idx], 0.f) : reinterpret_cast<const float2*>(in)[idx];
  }
  for (int i = threadIdx.x; i < N1; i += blockDim.x) {
    float sn, cs;
    sincospif(sign * 2.f * static_cast<float>(i) / static_cast<float>(N1), &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < N1 * 8; o += blockDim.x) {
    const int k1 = o >> 3, c = o & 7;
    float2 acc = make_float2(0.f, 0.f);
    int idx = 0;  // (n1 * k1) mod N1
    for (int n1 = 0; n1 < N1; ++n1) {
      const float2 v = cmulf(s[n1 * 8 + c], tw[idx]);
      acc.x += v.x;
      acc.y += v.y;
      idx += k1;
      if (idx >= N1) idx -= N1;
    }
    float sn, cs;
    sincospif(sign * 2.f * static_cast<float>(k1 * (c0 + c)) / static_cast<float>(N), &sn, &cs);
    Y[b * N + static_cast<long long>(k1) * 256 + c0 + c] = cmulf(acc, make_float2(cs, sn));
  }
}
// stage B: per k1 row 256-point radix-2 DIF; out[k1 + N1*k2]
__global__ void __launch_bounds__(128)
fftmix_rows_kernel(const float2* __restrict__ Y, const float2* __restrict__ tw512, int N1, float sign,
                   float2* __restrict__ out) {
  __shared__ float2 s[256];
  const int b = blockIdx.y, k1 = blockIdx.x;
  const long long N = static_cast<long long>(N1) * 256;
  const float2* row = Y + b * N + static_cast<long long>(k1) * 256;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = row[i];
  __syncthreads();
  for (int lh = 7; lh >= 0; --lh) {
    const int h = 1 << lh;
    for (int j = threadIdx.x; j < 128; j += blockDim.x) {
      const int pos = j & (h - 1);
      const int i0 = ((j >> lh) << (lh + 1)) + pos;
      const float2 a = s[i0], bb = s[i0 + h];
      float2 w = tw512[(pos << (7 - lh)) * 2];  // W_256^k = exp(-2 pi i k/256)
      if (sign > 0.f) w.y = -w.y;
      s[i0] = make_float2(a.x + bb.x, a.y + bb.y);
      s[i0 + h] = cmulf(make_float2(a.x - bb.x, a.y - bb.y), w);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int k2 = __brev(static_cast<unsigned>(i)) >> 24;
    out[b * N + k1 + static_cast<long long>(N1) * k2] = s[i];
  }
}

// ------------------------------------------------------------------------------------------------ min-phase stages
// N-point vectors per utterance; `mode` selects the stage (see spectral.py / blind.py for the chain)
struct MpArgs {
  const float2* c0;  // complex input
  const float2* c1;  // second complex input (Hf) where needed
  const float* r0;   // real inputs
  const float* r1;
  float2* oc;        // complex output
  float* or0;        // real outputs
  float* or1;
  int N, T;          // FFT length, kept length
  float invN;
};
__global__ void minphase_pw_kernel(MpArgs a, int mode) {
  const int b = blockIdx.y;
  const long long base = static_cast<long long>(b) * a.N;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.N; k += gridDim.x * blockDim.x) {
    const long long i = base + k;
    switch (mode) {
      case 0: {  // after FFT#1: m = |Hf|, Lc = (log(m + 1e-8), 0)
        const float2 h = a.c0[i];
        const float m = sqrtf(h.x * h.x + h.y * h.y);
        a.or0[i] = m;
        a.oc[i] = make_float2(logf(m + 1e-8f), 0.f);
      } break;
      case 1: {  // after FFT#2: D = C * window (2 for k < N/2, else 0); also used for G_C = w * raw/N (invN)
        const float2 c = a.c0[i];
        const float w = (k < a.N / 2) ? 2.f * a.invN : 0.f;
        a.oc[i] = make_float2(c.x * w, c.y * w);
      } break;
      case 2: {  // after IFFT#3 (unnormalised): phi = -Im(c)/N ; E = m e^{j phi}
        const float phi = -a.c0[i].y * a.invN;
        float sn, cs;
        sincosf(phi, &sn, &cs);
        const float m = a.r0[i];
        a.or0[i] = phi;
        a.oc[i] = make_float2(m * cs, m * sn);
      } break;
      case 3: {  // after IFFT#4: hm = Re/N for k < T, h[0] = direct-path constant (r0[0] holds it) -> or0 [B][T]
        if (k < a.T) a.or0[static_cast<long long>(b) * a.T + k] = (k == 0) ? a.r0[0] : a.c0[i].x * a.invN;
      } break;
      case 4: {  // backward start: G_z real = dh2[k] for 1 <= k < T else 0   (r0: [B][T]) -> or0 [B][N]
        a.or0[i] = (k >= 1 && k < a.T) ? a.r0[static_cast<long long>(b) * a.T + k] : 0.f;
      } break;
      case 5: {  // G_E = raw/N ; t = e^{-j phi} G_E ; g_m1 = Re t ; G_c = (0, -m Im t)
        const float2 g = make_float2(a.c0[i].x * a.invN, a.c0[i].y * a.invN);
        float sn, cs;
        sincosf(a.r1[i], &sn, &cs);  // phi
        const float tr = cs * g.x + sn * g.y, ti = cs * g.y - sn * g.x;
        a.or0[i] = tr;
        a.oc[i] = make_float2(0.f, -a.r0[i] * ti);  // r0 = m
      } break;
      case 6: {  // G_L = Re(raw3); g_m = g_m1 + G_L/(m+1e-8); G_Hf = g_m * Hf / m
        const float m = a.r0[i];
        const float gm = a.r1[i] + a.c0[i].x / (m + 1e-8f);
        const float2 h = a.c1[i];
        a.oc[i] = (m > 0.f) ? make_float2(gm * h.x / m, gm * h.y / m) : make_float2(0.f, 0.f);
      } break;
      case 7: {  // g_u = Re(raw4) -> dh [B][T] (k < T)
        if (k < a.T) a.or0[static_cast<long long>(b) * a.T + k] = a.c0[i].x;
      } break;
    }
  }
}

// ------------------------------------------------------------------------------------------------ Adam + projection
// p, g, m, v: [B][n_per]; first 25 = decays, next 25 = weights, rest = phases (no projection)
__global__ void adam_project_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, long long n_total, int n_per, float step_size, float beta1,
                                    float beta2, float eps, float bc2_sqrt, float dmin, float dmax,
                                    float wmin, float wmax) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    float mi = m[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    float pi = p[i] - step_size * (mi / denom);
    const int j = static_cast<int>(i % n_per);
    // project_params (subband_filtering.py:298-331).  fminf/fmaxf would map a NaN parameter onto the bound; the
    // reference asserts on NaN (:330-331), so NaN must survive the projection and reach the sampler's finiteness check.
    if (pi == pi) {
      if (j < 25) pi = fminf(fmaxf(pi, dmin), dmax);
      else if (j < 50) pi = fminf(fmaxf(pi, wmin), wmax);
    }
    p[i] = pi;
  }
}
}  // namespace buddy

using namespace buddy;

extern "C" int buddy_subband_fir(const float* a, const float* h_or_dy, int64_t h_batch_stride, float* out, int batch,
                                 int F, int Tx, int Nf, int pre, int mode, int accumulate, void* stream) {
  if (Tx > kFirMaxT || Nf > kFirMaxN || mode < 0 || mode > 2) {
    set_last_error("buddy_subband_fir: unsupported size Tx=%d Nf=%d mode=%d", Tx, Nf, mode);
    return BUDDY_ERR_UNSUPPORTED;
  }
  subband_fir_kernel<<<dim3(F, batch), 192, 0, STREAM>>>(reinterpret_cast<const float2*>(a),
                                                         reinterpret_cast<const float2*>(h_or_dy), h_batch_stride / 2,
                                                         reinterpret_cast<float2*>(out), F, Tx, Nf, pre, mode,
                                                         accumulate);
  LAUNCH_END("subband_fir_kernel");
}
extern "C" int buddy_blind_design_fwd(const float* decays, const float* weights, const float* phases, const int* kidx,
                                      const float* frac, const float* corr, const float* dpmag, int batch, int F,
                                      int Nf, float* A, float* H0, void* stream) {
  DesignTabs tb{kidx, frac, corr, dpmag};
  design_fwd_kernel<<<dim3(Nf, batch), 256, 0, STREAM>>>(decays, weights, phases, tb, F, Nf, A,
                                                         reinterpret_cast<float2*>(H0));
  LAUNCH_END("design_fwd_kernel");
}
extern "C" int buddy_blind_design_bwd(const float* decays, const float* weights, const float* phases, const float* A,
                                      const int* kidx, const float* frac, const float* corr, const float* dpmag,
                                      const float* G, int batch, int F, int Nf, float* dphases, float* ddecays,
                                      float* dweights, float* scratch, void* stream) {
  DesignTabs tb{kidx, frac, corr, dpmag};
  if (F > kMaxF || !scratch) {
    set_last_error("buddy_blind_design_bwd: F must be <= %d and scratch [batch][50][Nf] must be given", kMaxF);
    return BUDDY_ERR_INVALID;
  }
  design_bwd_kernel<<<dim3(Nf, batch), 256, 0, STREAM>>>(decays, weights, phases, A, tb,
                                                         reinterpret_cast<const float2*>(G), F, Nf, dphases, scratch);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BUDDY_CHECK_LAUNCH("design_bwd_kernel");
  design_bwd_reduce_kernel<<<batch, 64, 0, STREAM>>>(scratch, Nf, ddecays, dweights);
  LAUNCH_END("design_bwd_reduce_kernel");
}
extern "C" int buddy_fft_mixed(const float* in, int in_real, float* work, float* out, int batch, int N1, int sign,
                               const float* tw512, void* stream) {
  if (N1 < 1 || N1 > 128) {
    set_last_error("buddy_fft_mixed: N1 must be in [1,128]");
    return BUDDY_ERR_UNSUPPORTED;
  }
  fftmix_cols_kernel<<<dim3(32, batch), 256, 0, STREAM>>>(in, in_real, N1, sign < 0 ? -1.f : 1.f,
                                                          reinterpret_cast<float2*>(work));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  fftmix_rows_kernel<<<dim3(N1, batch), 128, 0, STREAM>>>(reinterpret_cast<const float2*>(work),
                                                          reinterpret_cast<const float2*>(tw512), N1,
                                                          sign < 0 ? -1.f : 1.f, reinterpret_cast<float2*>(out));
  LAUNCH_END("fftmix kernels");
}
extern "C" int buddy_minphase_pw(int mode, const float* c0, const float* c1, const float* r0, const float* r1,
                                 float* oc, float* or0, float* or1, int batch, int N, int T, int scale_inv_n,
                                 void* stream) {
  MpArgs a{reinterpret_cast<const float2*>(c0), reinterpret_cast<const float2*>(c1), r0, r1,
           reinterpret_cast<float2*>(oc), or0, or1, N, T, 1.f / static_cast<float>(N)};
  if (mode == 1 && !scale_inv_n) a.invN = 1.f;  // stage 1 doubles as forward window (no 1/N) and backward (1/N)
  minphase_pw_kernel<<<dim3((N + 255) / 256, batch), 256, 0, STREAM>>>(a, mode);
  LAUNCH_END("minphase_pw_kernel");
}
extern "C" int buddy_adam_project(float* p, const float* g, float* m, float* v, int batch, int n_per, int step,
                                  float lr, float beta1, float beta2, float eps, float dmin, float dmax, float wmin,
                                  float wmax, void* stream) {
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const long long n = static_cast<long long>(batch) * n_per;
  long long gx = (n + 255) / 256;
  if (gx > 148 * 8) gx = 148 * 8;
  adam_project_kernel<<<static_cast<unsigned>(gx), 256, 0, STREAM>>>(p, g, m, v, n, n_per,
                                                                     static_cast<float>(static_cast<double>(lr) / bc1),
                                                                     beta1, beta2, eps,
                                                                     static_cast<float>(sqrt(bc2)), dmin, dmax, wmin,
                                                                     wmax);
  LAUNCH_END("adam_project_kernel");
}
