// Weighted prediction error (WPE) dereverberation of one-channel STFTs — the warm start of the blind sampler
// (`warm_initialization.mode: "wpe_scaled"`, conf/tester/blind_dereverberation_BUDDy.yaml:76-81).
//
// Replaces, on the GPU, the reference's call into the third-party numpy package nara_wpe
// (testing/EulerHeunSamplerDPS.py:32-54: `wpe(Y, taps=50, delay=2, iterations=5, statistics_mode='full')`):
// per frequency bin, `iterations` rounds of
//     lambda[t] = max(|x[t]|^2, 1e-10 * max_t |x[t]|^2)
//     R = sum_t ytilde[t] ytilde[t]^H / lambda[t],   P = sum_t ytilde[t] conj(y[t]) / lambda[t]
//     g = R^-1 P,   x[t] = y[t] - g^H ytilde[t],      ytilde[t] = (y[t-delay], ..., y[t-delay-taps+1]) (zeros before 0)
// (Nakatani et al., IEEE TASLP 18(7), 2010; Drude et al., ITG 2018).  The 513 / 257 bins are independent small
// Hermitian problems: one CTA per (bin, utterance), everything in shared memory, double precision like the numpy
// original (complex128): correlation matrix by one thread per matrix entry, in-place Cholesky, two triangular solves.
// A non-positive pivot (silent bin) zeroes that tap's coefficient — the minimum-norm solution nara_wpe's lstsq
// fallback returns for a singular system.
#include <atomic>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {
extern std::atomic<long long> g_launches;

constexpr int kWpeMaxTaps = 64;
constexpr int kWpeMaxT = 4096;   // frames per utterance (30 s at shift 128: 3754)

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

__global__ void __launch_bounds__(256, 1)
wpe_kernel(const float2* __restrict__ Y, int F, int T, int taps, int delay, int iters, float2* __restrict__ Z) {
  extern __shared__ __align__(16) double wpe_smem[];
  // (all double2 arrays first: 16-byte alignment whatever the parity of T)
  double2* y = reinterpret_cast<double2*>(wpe_smem);   // [T]
  double2* x = y + T;                                  // [T]
  double2* R = x + T;                                  // [taps][taps], lower triangle used; becomes L
  double2* P = R + taps * taps;                        // [taps] -> z -> g
  double* ip = reinterpret_cast<double*>(P + taps);    // [T]  1 / lambda
  double* red = ip + T;                                // [32]
  int* okp = reinterpret_cast<int*>(red + 32);         // [taps] pivot usable
  const int f = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float2* yin = Y + (static_cast<long long>(b) * F + f) * T;
  for (int t = tid; t < T; t += nt) {
    const float2 v = yin[t];
    y[t] = x[t] = make_double2(v.x, v.y);
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    // ---- inverse power with the relative floor
    double mx = 0.0;
    for (int t = tid; t < T; t += nt) {
      const double pw = x[t].x * x[t].x + x[t].y * x[t].y;
      ip[t] = pw;
      mx = fmax(mx, pw);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
      double m = 0.0;
      for (int w = 0; w < (nt >> 5); ++w) m = fmax(m, red[w]);
      red[0] = 1e-10 * m;
    }
    __syncthreads();
    const double eps = red[0];
    // (an all-zero bin has eps = 0: weights 0 -> R = 0 -> g = 0 -> x = y = 0, instead of the 0 * inf = NaN of the
    // numpy original)
    for (int t = tid; t < T; t += nt) ip[t] = eps > 0.0 ? 1.0 / fmax(ip[t], eps) : 0.0;
    __syncthreads();
    // ---- R (lower triangle, i >= j) and P
    const int npairs = taps * (taps + 1) / 2;
    for (int q = tid; q < npairs + taps; q += nt) {
      if (q < npairs) {
        int i = static_cast<int>((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > q) --i;
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        const int j = q - i * (i + 1) / 2;
        double2 acc = make_double2(0.0, 0.0);
        for (int t = delay + i; t < T; ++t) {
          const double2 a = y[t - delay - i], c = y[t - delay - j];
          const double w = ip[t];
          acc.x += w * (a.x * c.x + a.y * c.y);
          acc.y += w * (a.y * c.x - a.x * c.y);
        }
        R[i * taps + j] = acc;
      } else {
        const int i = q - npairs;
        double2 acc = make_double2(0.0, 0.0);
        for (int t = delay + i; t < T; ++t) {
          const double2 a = y[t - delay - i], c = y[t];
          const double w = ip[t];
          acc.x += w * (a.x * c.x + a.y * c.y);
          acc.y += w * (a.y * c.x - a.x * c.y);
        }
        P[i] = acc;
      }
    }
    __syncthreads();
    // ---- in-place Cholesky R = L L^H (right-looking)
    for (int k = 0; k < taps; ++k) {
      if (tid == 0) {
        const double d = R[k * taps + k].x;
        const int ok = (d > 0.0 && d < 1e300) ? 1 : 0;
        okp[k] = ok;
        R[k * taps + k] = make_double2(ok ? sqrt(d) : 1.0, 0.0);
      }
      __syncthreads();
      const double inv = okp[k] ? 1.0 / R[k * taps + k].x : 0.0;
      for (int i = k + 1 + tid; i < taps; i += nt) {
        R[i * taps + k].x *= inv;
        R[i * taps + k].y *= inv;
      }
      __syncthreads();
      const int n = taps - k - 1;
      for (int q = tid; q < n * n; q += nt) {
        const int i = k + 1 + q / n, j = k + 1 + q % n;
        if (j <= i) {
          const double2 u = cmul_conj(R[i * taps + k], R[j * taps + k]);
          R[i * taps + j].x -= u.x;
          R[i * taps + j].y -= u.y;
        }
      }
      __syncthreads();
    }
    // ---- L z = P, L^H g = z (one warp; rows are short)
    if (tid < 32) {
      for (int i = 0; i < taps; ++i) {
        double2 s = make_double2(0.0, 0.0);
        for (int j = tid; j < i; j += 32) {
          const double2 u = cmul(R[i * taps + j], P[j]);
          s.x += u.x;
          s.y += u.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
          s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        }
        if (tid == 0) {
          const double inv = okp[i] ? 1.0 / R[i * taps + i].x : 0.0;
          P[i] = make_double2((P[i].x - s.x) * inv, (P[i].y - s.y) * inv);
        }
        __syncwarp();
      }
      for (int i = taps - 1; i >= 0; --i) {
        double2 s = make_double2(0.0, 0.0);
        for (int j = i + 1 + tid; j < taps; j += 32) {
          const double2 l = R[j * taps + i];                       // (L^H)[i][j] = conj(L[j][i])
          const double2 u = cmul(make_double2(l.x, -l.y), P[j]);
          s.x += u.x;
          s.y += u.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
          s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        }
        if (tid == 0) {
          const double inv = okp[i] ? 1.0 / R[i * taps + i].x : 0.0;
          P[i] = make_double2((P[i].x - s.x) * inv, (P[i].y - s.y) * inv);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- x[t] = y[t] - sum_i conj(g[i]) y[t - delay - i]
    for (int t = tid; t < T; t += nt) {
      double2 acc = y[t];
      const int imax = min(taps, t - delay + 1);
      for (int i = 0; i < imax; ++i) {
        const double2 u = cmul_conj(y[t - delay - i], P[i]);     // y~ * conj(g)
        acc.x -= u.x;
        acc.y -= u.y;
      }
      x[t] = acc;
    }
    __syncthreads();
  }
  float2* zo = Z + (static_cast<long long>(b) * F + f) * T;
  for (int t = tid; t < T; t += nt) zo[t] = make_float2(static_cast<float>(x[t].x), static_cast<float>(x[t].y));
}
}  // namespace buddy

using namespace buddy;

extern "C" int buddy_wpe(const float* Y, int batch, int F, int T, int taps, int delay, int iterations, float* Z,
                         void* stream) {
  if (!Y || !Z || batch <= 0 || F <= 0 || T <= 0 || taps <= 0 || taps > kWpeMaxTaps || T > kWpeMaxT || delay < 0 ||
      iterations < 0) {
    set_last_error("buddy_wpe: unsupported size (batch %d F %d T %d taps %d delay %d iterations %d; taps <= %d, T <= %d)",
                   batch, F, T, taps, delay, iterations, kWpeMaxTaps, kWpeMaxT);
    return BUDDY_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(double) * (static_cast<size_t>(5) * T + 2 * static_cast<size_t>(taps) * taps + 2 * taps + 32) +
                      sizeof(int) * taps;
  if (smem > 227 * 1024) {
    set_last_error("buddy_wpe: %d frames x %d taps need %zu bytes of shared memory (> 227 KB)", T, taps, smem);
    return BUDDY_ERR_UNSUPPORTED;
  }
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    int e = check_cuda(cudaFuncSetAttribute(wpe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(wpe_kernel)");
    if (e) return e;
    attr_bytes = smem;
  }
  wpe_kernel<<<dim3(F, batch), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(Y), F, T, taps, delay, iterations, reinterpret_cast<float2*>(Z));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BUDDY_CHECK_LAUNCH("wpe_kernel");
  return 0;
}
