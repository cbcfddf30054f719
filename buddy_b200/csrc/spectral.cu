// Spectral kernels: STFT / iSTFT as small fp32 DFT-GEMMs with exact frame indexing, overlap-add gather,
// signal padding (+adjoints), the compressed-spectrum likelihood loss with its analytic gradient, and the
// 2^k-point FFT convolution of the informed reverb operator (four-step FFT in shared memory).
//
// Reference semantics
//   torch.stft / torch.istft as used by NCSNppTime.stft/istft (networks/ncsnpp.py:473-496) and by the
//   operators' apply_stft/apply_istft (testing/operators/subband_filtering.py:41-65,76-80; reverb.py:54-84);
//   l2_comp_stft_summean (utils/losses.py:59-64,74-76); fast_apply_RIR (utils/reverb_utils.py:25-60).
// The DFT matrices (window, 1/N, onesided weights, 1/sqrt(sum w^2) folded in) are built on the host in fp64.
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

// ------------------------------------------------------------------------------------------------
// analysis: out[b][f][t][c] = sum_{n<K} mat[2f+c][n] * sig[b][t*hop + n]      (t < frames; zero for t >= frames)
// 64 (m) x 64 (t) tiles, k-step 16, 256 threads x (4x4) outputs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dft_analysis_kernel(const float* __restrict__ sig, long long sig_ld, const float* __restrict__ mat, int M, int K,
                    int hop, int frames, int Tout, float* __restrict__ out) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * 64, t0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* sb = sig + b * sig_ld;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    {
      // A: 64 rows x 16 k  (each thread 4 consecutive k of one row)
      const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        As[kq + j][r] = (m < M && k < K) ? __ldg(mat + static_cast<long long>(m) * K + k) : 0.f;
      }
      // B: 64 frames x 16 k
      const int t = t0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        Bs[kq + j][r] = (t < frames && k < K) ? __ldg(sb + static_cast<long long>(t) * hop + k) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int F = M >> 1;
#pragma unroll
  for (int i = 0; i < 4; i += 2) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int f = m >> 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t < Tout)
        reinterpret_cast<float2*>(out)[(static_cast<long long>(b) * F + f) * Tout + t] =
            (t < frames) ? make_float2(acc[i][j], acc[i + 1][j]) : make_float2(0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// synthesis: fr[b][t][n] = sum_{m<M} S[b][m/2][t][m%2] * mat[m][n]     (t < frames, n < K)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dft_synthesis_kernel(const float* __restrict__ S, int Tin, const float* __restrict__ mat, int M, int K, int frames,
                     float* __restrict__ fr) {
  __shared__ float As[16][68];  // [m][t]
  __shared__ float Bs[16][68];  // [m][n]
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int F = M >> 1;
  float acc[4][4] = {};
  for (int m0 = 0; m0 < M; m0 += 16) {
    {
      // A: 8 bin-pairs x 64 t float2
      const int fp = threadIdx.x >> 5;          // 0..7
      const int tt = (threadIdx.x & 31) * 2;    // 0..62
      const int f = (m0 >> 1) + fp;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int t = t0 + tt + j;
        float2 v = make_float2(0.f, 0.f);
        if (f < F && t < frames) v = __ldg(reinterpret_cast<const float2*>(S) + (static_cast<long long>(b) * F + f) * Tin + t);
        As[fp * 2][tt + j] = v.x;
        As[fp * 2 + 1][tt + j] = v.y;
      }
      // B: 16 m x 64 n
      const int r = threadIdx.x >> 4, nq = (threadIdx.x & 15) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + nq + j;
        Bs[r][nq + j] = (m < M && n < K) ? __ldg(mat + static_cast<long long>(m) * K + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= frames) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < K) fr[(static_cast<long long>(b) * frames + t) * K + n] = acc[i][j];
    }
  }
}

// out[b][s] = tab[s + off] * scale_b[b] * sum_t fr[b][t][s + off - t*hop]      (0 <= s + off - t*hop < K)
__global__ void ola_gather_kernel(const float* __restrict__ fr, int frames, int K, int hop, int off, int n_out,
                                  const float* __restrict__ tab, const float* __restrict__ scale_b,
                                  float* __restrict__ out, long long out_ld) {
  const int b = blockIdx.y;
  const float sc = scale_b ? scale_b[b] : 1.f;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_out; s += gridDim.x * blockDim.x) {
    const int j = s + off;
    int t_hi = j / hop;
    if (t_hi > frames - 1) t_hi = frames - 1;
    int t_lo = (j - K + hop) / hop;  // smallest t with j - t*hop < K  (ceil((j-K+1)/hop))
    if (j - K + 1 <= 0) t_lo = 0;
    float a = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) a += __ldg(fr + (static_cast<long long>(b) * frames + t) * K + (j - t * hop));
    if (tab) a *= tab[j];
    out[b * out_ld + s] = a * sc;
  }
}

// padded[b][j] = tab[j] * scale_b[b] * x[b][src(j - left)]   mode 0: zero outside [0,N); mode 1: reflect
__global__ void pad_signal_kernel(const float* __restrict__ x, long long x_ld, int N, int left, int total, int mode,
                                  const float* __restrict__ tab, const float* __restrict__ scale_b,
                                  float* __restrict__ out) {
  const int b = blockIdx.y;
  const float sc = scale_b ? scale_b[b] : 1.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    int s = j - left;
    float v = 0.f;
    if (mode == 1) {
      if (s < 0) s = -s;
      if (s >= N) s = 2 * (N - 1) - s;
      if (s >= 0 && s < N) v = x[b * x_ld + s];
    } else if (s >= 0 && s < N) {
      v = x[b * x_ld + s];
    }
    if (tab) v *= tab[j];
    out[static_cast<long long>(b) * total + j] = v * sc;
  }
}

// adjoint of the reflect pad: dx[s] = (dxp[s+L] + [1<=s<=L] dxp[L-s] + [N-1-L<=s<=N-2] dxp[2N-2+L-s]) * scale_b
__global__ void reflect_fold_kernel(const float* __restrict__ dxp, int N, int L, const float* __restrict__ scale_b,
                                    float* __restrict__ dx, long long dx_ld) {
  const int b = blockIdx.y;
  const float sc = scale_b ? scale_b[b] : 1.f;
  const float* p = dxp + static_cast<long long>(b) * (N + 2 * L);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
    float v = p[s + L];
    if (s >= 1 && s <= L) v += p[L - s];
    if (s >= N - 1 - L && s <= N - 2) v += p[2 * N - 2 + L - s];
    dx[b * dx_ld + s] = v * sc;
  }
}

// ------------------------------------------------------------------------------------------------
// compressed-spectrum L2 ("l2_comp_stft_summean"): per utterance
//   loss_b = weight/frames * sum_{f,t} |Yc - Xc|^2,  Zc = (|Z|+1e-8)^c * exp(j angle Z)
//   grad[b][f][t] = dloss_b / dX (re, im)
// ------------------------------------------------------------------------------------------------
__global__ void comp_loss_kernel(const float2* __restrict__ Y, const float2* __restrict__ X, long long per_utt,
                                 float cexp, float wnorm, double* __restrict__ loss, float2* __restrict__ grad) {
  const int b = blockIdx.y;
  const float2* y = Y + b * per_utt;
  const float2* x = X + b * per_utt;
  float2* g = grad ? grad + b * per_utt : nullptr;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_utt;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 yv = y[i], xv = x[i];
    const float my = sqrtf(yv.x * yv.x + yv.y * yv.y), mx = sqrtf(xv.x * xv.x + xv.y * xv.y);
    // unit phasors; angle(0) = 0 in torch
    const float uyx = my > 0.f ? yv.x / my : 1.f, uyy = my > 0.f ? yv.y / my : 0.f;
    const float uxx = mx > 0.f ? xv.x / mx : 1.f, uxy = mx > 0.f ? xv.y / mx : 0.f;
    const float cy = powf(my + 1e-8f, cexp), cx = powf(mx + 1e-8f, cexp);
    const float dr = cx * uxx - cy * uyx, di = cx * uxy - cy * uyy;  // D = Xc - Yc
    acc += dr * dr + di * di;
    if (g) {
      float2 o = make_float2(0.f, 0.f);
      if (mx > 0.f) {
        // E = conj(u) D ; G = 2 u (alpha Re E + j beta Im E), alpha = c (m+eps)^(c-1), beta = (m+eps)^c / m
        const float er = uxx * dr + uxy * di, ei = uxx * di - uxy * dr;
        const float alpha = cexp * cx / (mx + 1e-8f), beta = cx / mx;
        const float pr = alpha * er, pi = beta * ei;
        o.x = 2.f * wnorm * (uxx * pr - uxy * pi);
        o.y = 2.f * wnorm * (uxx * pi + uxy * pr);
      }
      g[i] = o;
    }
  }
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss + b, static_cast<double>(v) * wnorm);
  }
}

// ------------------------------------------------------------------------------------------------
// row statistics: out[b] = (sum x, sum x^2) in fp64
// ------------------------------------------------------------------------------------------------
__global__ void row_stats_kernel(const float* __restrict__ x, long long ld, int n, double* __restrict__ out) {
  const int b = blockIdx.y;
  double s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = x[b * ld + i];
    s += v;
    q += v * v;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  __shared__ double red[2][32];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = (threadIdx.x < (blockDim.x >> 5)) ? red[0][threadIdx.x] : 0.0;
    double c = (threadIdx.x < (blockDim.x >> 5)) ? red[1][threadIdx.x] : 0.0;
    a = warp_sum(a);
    c = warp_sum(c);
    if (threadIdx.x == 0) {
      atomicAdd(out + 2 * b, a);
      atomicAdd(out + 2 * b + 1, c);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FFT convolution, L = 256 * N2 points (N2 = 512 -> 2^17), four-step decomposition n = n1*N2 + n2, k = k1 + 256*k2
//   cols_fwd : 256-pt DIF FFT over n1 for 8 adjacent columns, times W_L^(n2*k1)        -> Y[k1][n2]
//   rows     : per k1 row: N2-pt DIF FFT -> (store spectrum | times H (or conj H) -> N2-pt inverse DIT
//              -> times conj W_L^(n2*k1))                                             -> T[k1][n2]
//   cols_inv : 256-pt inverse DIT over k1, real part / L                              -> y[n1*N2+n2]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// in-place radix-2 DIF (natural in -> bit-reversed out), forward sign; tw = W_N^k (k < N/2), `tws` table stride
template <int LOGN>
__device__ __forceinline__ void fft_dif(float2* s, int sstride, const float2* __restrict__ tw, int tws, int tid,
                                        int nthreads) {
  constexpr int N = 1 << LOGN;
  for (int lh = LOGN - 1; lh >= 0; --lh) {
    const int h = 1 << lh;
    for (int j = tid; j < N / 2; j += nthreads) {
      const int pos = j & (h - 1);
      const int i0 = ((j >> lh) << (lh + 1)) + pos;
      const float2 a = s[i0 * sstride], b = s[(i0 + h) * sstride];
      const float2 w = tw[(pos << (LOGN - 1 - lh)) * tws];
      s[i0 * sstride] = make_float2(a.x + b.x, a.y + b.y);
      s[(i0 + h) * sstride] = cmul(make_float2(a.x - b.x, a.y - b.y), w);
    }
    __syncthreads();
  }
}
// in-place radix-2 DIT (bit-reversed in -> natural out), INVERSE sign (conj twiddles), unnormalised
template <int LOGN>
__device__ __forceinline__ void ifft_dit(float2* s, int sstride, const float2* __restrict__ tw, int tws, int tid,
                                         int nthreads) {
  constexpr int N = 1 << LOGN;
  for (int lh = 0; lh < LOGN; ++lh) {
    const int h = 1 << lh;
    for (int j = tid; j < N / 2; j += nthreads) {
      const int pos = j & (h - 1);
      const int i0 = ((j >> lh) << (lh + 1)) + pos;
      float2 w = tw[(pos << (LOGN - 1 - lh)) * tws];
      w.y = -w.y;
      const float2 a = s[i0 * sstride], t = cmul(s[(i0 + h) * sstride], w);
      s[i0 * sstride] = make_float2(a.x + t.x, a.y + t.y);
      s[(i0 + h) * sstride] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

constexpr int kN1 = 256, kLogN1 = 8, kColsPerBlock = 8;

// x real [b][x_ld] (zero beyond n_in) -> Y complex [b][256][N2]
template <int LOGN2>
__global__ void __launch_bounds__(256)
fftconv_cols_fwd_kernel(const float* __restrict__ x, long long x_ld, int n_in, const float2* __restrict__ tw512,
                        float2* __restrict__ Y) {
  constexpr int N2 = 1 << LOGN2;
  __shared__ float2 s[kN1 * kColsPerBlock];  // [n1][col]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kColsPerBlock;
  for (int i = threadIdx.x; i < kN1 * kColsPerBlock; i += blockDim.x) {
    const int n1 = i / kColsPerBlock, c = i % kColsPerBlock;
    const int n = n1 * N2 + c0 + c;
    s[i] = make_float2(n < n_in ? x[b * x_ld + n] : 0.f, 0.f);
  }
  __syncthreads();
  // 8 interleaved 256-pt FFTs: thread group (tid % 8) owns a column
  {
    const int col = threadIdx.x % kColsPerBlock;
    fft_dif<kLogN1>(s + col, kColsPerBlock, tw512, 2, threadIdx.x / kColsPerBlock, 256 / kColsPerBlock);
  }
  float2* Yb = Y + static_cast<long long>(b) * kN1 * N2;
  for (int i = threadIdx.x; i < kN1 * kColsPerBlock; i += blockDim.x) {
    const int r = i / kColsPerBlock, c = i % kColsPerBlock;
    const int k1 = bitrev(r, kLogN1);
    const int n2 = c0 + c;
    float sn, cs;
    sincospif(-2.f * static_cast<float>(k1 * n2) / static_cast<float>(kN1 * N2), &sn, &cs);
    Yb[k1 * N2 + n2] = cmul(s[i], make_float2(cs, sn));
  }
}

// mode 0: spectrum only (Z <- FFT rows, bit-reversed k2 order kept);  mode 1: times H;  mode 2: times conj(H)
template <int LOGN2>
__global__ void __launch_bounds__(256)
fftconv_rows_kernel(float2* __restrict__ Y, const float2* __restrict__ Hs, long long h_bs, const float2* __restrict__ tw,
                    int tws, int mode) {
  constexpr int N2 = 1 << LOGN2;
  __shared__ float2 s[N2];
  const int b = blockIdx.y, k1 = blockIdx.x;
  float2* row = Y + (static_cast<long long>(b) * kN1 + k1) * N2;
  for (int i = threadIdx.x; i < N2; i += blockDim.x) s[i] = row[i];
  __syncthreads();
  fft_dif<LOGN2>(s, 1, tw, tws, threadIdx.x, blockDim.x);
  if (mode == 0) {
    for (int i = threadIdx.x; i < N2; i += blockDim.x) row[i] = s[i];
    return;
  }
  const float2* hrow = Hs + b * h_bs + static_cast<long long>(k1) * N2;
  for (int i = threadIdx.x; i < N2; i += blockDim.x) {
    float2 h = hrow[i];
    if (mode == 2) h.y = -h.y;
    s[i] = cmul(s[i], h);
  }
  __syncthreads();
  ifft_dit<LOGN2>(s, 1, tw, tws, threadIdx.x, blockDim.x);
  for (int n2 = threadIdx.x; n2 < N2; n2 += blockDim.x) {
    float sn, cs;
    sincospif(2.f * static_cast<float>(k1 * n2) / static_cast<float>(kN1 * N2), &sn, &cs);
    row[n2] = cmul(s[n2], make_float2(cs, sn));
  }
}

template <int LOGN2>
__global__ void __launch_bounds__(256)
fftconv_cols_inv_kernel(const float2* __restrict__ T, const float2* __restrict__ tw512, int n_out,
                        float* __restrict__ y, long long y_ld) {
  constexpr int N2 = 1 << LOGN2;
  __shared__ float2 s[kN1 * kColsPerBlock];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kColsPerBlock;
  const float2* Tb = T + static_cast<long long>(b) * kN1 * N2;
  for (int i = threadIdx.x; i < kN1 * kColsPerBlock; i += blockDim.x) {
    const int r = i / kColsPerBlock, c = i % kColsPerBlock;
    s[i] = Tb[bitrev(r, kLogN1) * N2 + c0 + c];
  }
  __syncthreads();
  {
    const int col = threadIdx.x % kColsPerBlock;
    ifft_dit<kLogN1>(s + col, kColsPerBlock, tw512, 2, threadIdx.x / kColsPerBlock, 256 / kColsPerBlock);
  }
  const float inv = 1.f / static_cast<float>(kN1 * N2);
  for (int i = threadIdx.x; i < kN1 * kColsPerBlock; i += blockDim.x) {
    const int n1 = i / kColsPerBlock, c = i % kColsPerBlock;
    const int n = n1 * N2 + c0 + c;
    if (n < n_out) y[b * y_ld + n] = s[i].x * inv;
  }
}

static int grid1(long long items, int threads, int cap = 148 * 8) {
  long long g = (items + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}


// ------------------------------------------------------------------------------------------------
// FFT-based STFT analysis / synthesis for the 1024-point transforms of the likelihood and the blind operator
// (apply_stft / apply_istft, testing/operators/subband_filtering.py:41-65): the same linear maps as
// dft_analysis / dft_synthesis with a matrix  mat[2f+c][n] = a[f] * w[n] * (cos, -sin)(2 pi f n / 1024), but as
// shared-memory radix-4 / radix-16 FFTs (50 kFLOP per frame instead of 1 MFLOP): 8 frames per CTA.
//   analysis : out[b][f][t] = a[f] * FFT(w * frame_t)[f]
//   synthesis: fr[b][t][n]  = w[n] * Re( sum_f a[f] S[b][f][t] e^{+2 pi i f n / 1024} )
// ------------------------------------------------------------------------------------------------
constexpr int kFftN = 1024;
constexpr int kFftFrames = 8;   // 74 KB of shared memory per CTA: three CTAs per SM hide the stage barriers
// Shared-memory index of element k of frame fr: one padding slot after every 32 elements and an odd frame stride, so
// that neither the bit-reversed gathers (addresses 32 elements apart across a warp) nor the frame-fastest loops
// (addresses one frame apart) land in one bank (both were 8- to 16-way conflicts with the dense [frame][1024] layout).
constexpr int kFftStride = kFftN + kFftN / 32 + 1;   // 1057
__device__ __forceinline__ int fidx(int fr, int k) { return fr * kFftStride + k + (k >> 5); }
// position of output bin k after the in-place radix-4 transform below: base-4 digit reversal of k (= bit reversal of the
// 10 bits, then the two bits of every digit swapped back)
__device__ __forceinline__ int digitrev4_1024(int k) {
  const unsigned r = __brev(static_cast<unsigned>(k)) >> 22;
  return static_cast<int>(((r & 0x155u) << 1) | ((r >> 1) & 0x155u));
}
// per-stage twiddle tables of the radix-4 transform: the stage with butterfly span q = 4^s (s = 4..0) reads
// tws[(q - 1) + 3 * pos + (m - 1)] = exp(-2 pi i m pos / (4 q)), pos < q, m = 1..3 — 1023 entries in all; consecutive
// threads (consecutive pos) read consecutive triples.  tw_g[k] = exp(-2 pi i k / 1024), k < 512.
__device__ __forceinline__ void fft_twiddles_to_smem(float2* tws, const float2* __restrict__ tw_g) {
  for (int i = threadIdx.x; i < 1023; i += blockDim.x) {
    const int s2 = (31 - __clz(i + 1)) & ~1;          // 2 s: q = 4^s = largest power of four <= i + 1
    const int r = i - ((1 << s2) - 1);
    const int pos = r / 3, m = r - 3 * pos + 1;
    const int idx = (m * pos) << (8 - s2);            // m * pos * (256 / q) < 768
    float2 w = tw_g[idx & 511];
    if (idx >= 512) w = make_float2(-w.x, -w.y);
    tws[i] = w;
  }
}
// radix-4 decimation-in-frequency butterfly on (a0..a3) with output twiddles (w1, w2, w3); inv: conjugate transform
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3, bool inv) {
  const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y), d = make_float2(a1.x - a3.x, a1.y - a3.y);
  const float2 t3 = inv ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);   // +i d  /  -i d
  a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
  a1 = make_float2(t1.x + t3.x, t1.y + t3.y);
  a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
  a3 = make_float2(t1.x - t3.x, t1.y - t3.y);
}
__device__ __forceinline__ float2 cmul(float2 v, float2 w, bool conj_w) {
  const float wy = conj_w ? -w.y : w.y;
  return make_float2(v.x * w.x - v.y * wy, v.x * wy + v.y * w.x);
}
// in-place 1024-point transform over `kFftFrames` frames [frame][1024] (natural order in, base-4 digit-reversed out):
// three radix-4 passes through shared memory (spans 256, 64, 16: consecutive threads touch consecutive elements) and one
// radix-16 pass in registers (spans 4 and 1 on 16 consecutive elements) — 4 barriers and ~2100 instructions per thread
// where the radix-2 version took 10 and ~4800.  tw = per-stage tables (fft_twiddles_to_smem); conj_tw: inverse
// transform (unnormalised).
__device__ __forceinline__ void fft1024_dif(float2* s, const float2* tw, bool conj_tw) {
#pragma unroll 1
  for (int s2 = 8; s2 >= 4; s2 -= 2) {          // q = 256, 64, 16
    const int q = 1 << s2;
    for (int i = threadIdx.x; i < kFftFrames * 256; i += blockDim.x) {
      const int fr = i >> 8, j = i & 255;
      const int pos = j & (q - 1);
      const int k0 = ((j >> s2) << (s2 + 2)) + pos;
      const int i0 = fidx(fr, k0), i1 = fidx(fr, k0 + q), i2 = fidx(fr, k0 + 2 * q), i3 = fidx(fr, k0 + 3 * q);
      float2 a0 = s[i0], a1 = s[i1], a2 = s[i2], a3 = s[i3];
      const float2* w = tw + (q - 1) + 3 * pos;
      const float2 w1 = w[0], w2 = w[1], w3 = w[2];
      bfly4(a0, a1, a2, a3, conj_tw);
      s[i0] = a0;
      s[i1] = cmul(a1, w1, conj_tw);
      s[i2] = cmul(a2, w2, conj_tw);
      s[i3] = cmul(a3, w3, conj_tw);
    }
    __syncthreads();
  }
  // spans 4 and 1: sixteen consecutive elements per thread (they share one padding offset: 16 | 32)
  float2 w4[9];                                  // exp(-2 pi i m pos / 16), pos = 1..3, m = 1..3
#pragma unroll
  for (int k = 0; k < 9; ++k) w4[k] = tw[3 + 3 + k];
  for (int g = threadIdx.x; g < kFftFrames * 64; g += blockDim.x) {
    const int fr = g >> 6, base = fidx(fr, (g & 63) << 4);
    float2 v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = s[base + e];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      bfly4(v[pos], v[pos + 4], v[pos + 8], v[pos + 12], conj_tw);
      if (pos > 0) {
        v[pos + 4] = cmul(v[pos + 4], w4[3 * (pos - 1)], conj_tw);
        v[pos + 8] = cmul(v[pos + 8], w4[3 * (pos - 1) + 1], conj_tw);
        v[pos + 12] = cmul(v[pos + 12], w4[3 * (pos - 1) + 2], conj_tw);
      }
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) bfly4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3], conj_tw);
#pragma unroll
    for (int e = 0; e < 16; ++e) s[base + e] = v[e];
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256)
fft_analysis_kernel(const float* __restrict__ sig, long long sig_ld, const float* __restrict__ wv,
                    const float* __restrict__ av, const float2* __restrict__ tw_g, int bins, int K, int hop,
                    int frames, int Tout, float2* __restrict__ out) {
  extern __shared__ float2 fsm[];
  float2* s = fsm;                            // [kFftFrames] frames, padded (fidx)
  float2* tw = fsm + kFftFrames * kFftStride;  // [1023] per-stage tables
  const int b = blockIdx.y, t0 = blockIdx.x * kFftFrames;
  fft_twiddles_to_smem(tw, tw_g);
  const float* sb = sig + static_cast<long long>(b) * sig_ld;
  for (int i = threadIdx.x; i < kFftFrames * kFftN; i += blockDim.x) {
    const int fr = i >> 10, n = i & 1023;
    const int t = t0 + fr;
    float v = 0.f;
    if (n < K && t < frames) v = __ldg(wv + n) * __ldg(sb + static_cast<long long>(t) * hop + n);
    s[fidx(fr, n)] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft1024_dif(s, tw, false);
  float2* ob = out + static_cast<long long>(b) * bins * Tout;
  for (int i = threadIdx.x; i < bins * kFftFrames; i += blockDim.x) {
    const int f = i / kFftFrames, fr = i % kFftFrames;
    const int t = t0 + fr;
    if (t >= Tout) continue;
    float2 v = make_float2(0.f, 0.f);
    if (t < frames) {
      const float2 x = s[fidx(fr, digitrev4_1024(f))];
      const float a = __ldg(av + f);
      // DC and Nyquist of a real signal are real: exact zeros as in the matrix form (-sin rows vanish) — a rounding-
      // level residue here would be a spurious non-zero gradient for Adam, which normalises every element
      v = make_float2(a * x.x, (f == 0 || 2 * f == kFftN) ? 0.f : a * x.y);
    }
    ob[static_cast<long long>(f) * Tout + t] = v;
  }
}
__global__ void __launch_bounds__(256)
fft_synthesis_kernel(const float2* __restrict__ S, int Tin, const float* __restrict__ wv, const float* __restrict__ av,
                     const float2* __restrict__ tw_g, int bins, int K, int frames, float* __restrict__ fr_out) {
  extern __shared__ float2 fsm[];
  float2* s = fsm;
  float2* tw = fsm + kFftFrames * kFftStride;
  const int b = blockIdx.y, t0 = blockIdx.x * kFftFrames;
  fft_twiddles_to_smem(tw, tw_g);
  const float2* Sb = S + static_cast<long long>(b) * bins * Tin;
  for (int i = threadIdx.x; i < kFftFrames * kFftN; i += blockDim.x) {
    const int f = i / kFftFrames, fr = i % kFftFrames;     // frames fastest: contiguous global segments per bin
    const int t = t0 + fr;
    float2 v = make_float2(0.f, 0.f);
    if (f < bins && t < frames) {
      const float2 x = __ldg(Sb + static_cast<long long>(f) * Tin + t);
      const float a = __ldg(av + f);
      v = make_float2(a * x.x, (f == 0 || 2 * f == kFftN) ? 0.f : a * x.y);   // imaginary DC / Nyquist do not contribute
    }
    s[fidx(fr, f)] = v;
  }
  __syncthreads();
  fft1024_dif(s, tw, true);
  for (int i = threadIdx.x; i < kFftFrames * K; i += blockDim.x) {
    const int fr = i / K, n = i - fr * K;
    const int t = t0 + fr;
    if (t < frames)
      fr_out[(static_cast<long long>(b) * frames + t) * K + n] = __ldg(wv + n) * s[fidx(fr, digitrev4_1024(n))].x;
  }
}

}  // namespace buddy

using namespace buddy;

extern "C" int buddy_dft_analysis(const float* sig, int64_t sig_ld, int batch, const float* mat, int M, int K, int hop,
                                  int frames, int Tout, float* out, void* stream) {
  if (M <= 0 || (M & 1) || K <= 0 || frames <= 0 || Tout < frames) {
    set_last_error("buddy_dft_analysis: bad shape M=%d K=%d frames=%d Tout=%d", M, K, frames, Tout);
    return BUDDY_ERR_INVALID;
  }
  dim3 grid((Tout + 63) / 64, (M + 63) / 64, batch);
  dft_analysis_kernel<<<grid, 256, 0, STREAM>>>(sig, sig_ld, mat, M, K, hop, frames, Tout, out);
  LAUNCH_END("dft_analysis_kernel");
}
extern "C" int buddy_dft_synthesis(const float* S, int batch, int Tin, const float* mat, int M, int K, int frames,
                                   float* fr, void* stream) {
  if (M <= 0 || (M & 1) || K <= 0 || frames <= 0 || Tin < frames) {
    set_last_error("buddy_dft_synthesis: bad shape");
    return BUDDY_ERR_INVALID;
  }
  dim3 grid((K + 63) / 64, (frames + 63) / 64, batch);
  dft_synthesis_kernel<<<grid, 256, 0, STREAM>>>(S, Tin, mat, M, K, frames, fr);
  LAUNCH_END("dft_synthesis_kernel");
}

static int fft_smem_attr() {
  static bool done = false;
  if (!done) {
    const int bytes = (kFftFrames * kFftStride + 1024) * sizeof(float2);
    int e = check_cuda(cudaFuncSetAttribute(fft_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                       "cudaFuncSetAttribute(fft_analysis_kernel)");
    if (e) return e;
    e = check_cuda(cudaFuncSetAttribute(fft_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                   "cudaFuncSetAttribute(fft_synthesis_kernel)");
    if (e) return e;
    done = true;
  }
  return 0;
}
extern "C" int buddy_fft_analysis(const float* sig, int64_t sig_ld, int batch, const float* wv, const float* av,
                                  const float* tw1024, int bins, int K, int hop, int frames, int Tout, float* out,
                                  void* stream) {
  if (bins <= 0 || bins > kFftN / 2 + 1 || K <= 0 || K > kFftN || frames <= 0 || Tout < frames || hop <= 0) {
    set_last_error("buddy_fft_analysis: bad shape (1024-point transform: bins <= 513, K <= 1024)");
    return BUDDY_ERR_INVALID;
  }
  int e = fft_smem_attr();
  if (e) return e;
  const size_t smem = (kFftFrames * kFftStride + 1024) * sizeof(float2);
  dim3 grid((Tout + kFftFrames - 1) / kFftFrames, batch);
  fft_analysis_kernel<<<grid, 256, smem, STREAM>>>(sig, sig_ld, wv, av, reinterpret_cast<const float2*>(tw1024), bins,
                                                   K, hop, frames, Tout, reinterpret_cast<float2*>(out));
  LAUNCH_END("fft_analysis_kernel");
}
extern "C" int buddy_fft_synthesis(const float* S, int batch, int Tin, const float* wv, const float* av,
                                   const float* tw1024, int bins, int K, int frames, float* fr, void* stream) {
  if (bins <= 0 || bins > kFftN / 2 + 1 || K <= 0 || K > kFftN || frames <= 0 || Tin < frames) {
    set_last_error("buddy_fft_synthesis: bad shape (1024-point transform: bins <= 513, K <= 1024)");
    return BUDDY_ERR_INVALID;
  }
  int e = fft_smem_attr();
  if (e) return e;
  const size_t smem = (kFftFrames * kFftStride + 1024) * sizeof(float2);
  dim3 grid((frames + kFftFrames - 1) / kFftFrames, batch);
  fft_synthesis_kernel<<<grid, 256, smem, STREAM>>>(reinterpret_cast<const float2*>(S), Tin, wv, av,
                                                    reinterpret_cast<const float2*>(tw1024), bins, K, frames, fr);
  LAUNCH_END("fft_synthesis_kernel");
}
extern "C" int buddy_ola_gather(const float* fr, int batch, int frames, int K, int hop, int off, int n_out,
                                const float* tab, const float* scale_b, float* out, int64_t out_ld, void* stream) {
  ola_gather_kernel<<<dim3(grid1(n_out, 256, 64), batch), 256, 0, STREAM>>>(fr, frames, K, hop, off, n_out, tab,
                                                                            scale_b, out, out_ld);
  LAUNCH_END("ola_gather_kernel");
}
extern "C" int buddy_pad_signal(const float* x, int64_t x_ld, int batch, int N, int left, int total, int mode,
                                const float* tab, const float* scale_b, float* out, void* stream) {
  if (mode == 1 && left >= N) {
    set_last_error("buddy_pad_signal: reflect pad needs left < N");
    return BUDDY_ERR_INVALID;
  }
  pad_signal_kernel<<<dim3(grid1(total, 256, 64), batch), 256, 0, STREAM>>>(x, x_ld, N, left, total, mode, tab,
                                                                            scale_b, out);
  LAUNCH_END("pad_signal_kernel");
}
extern "C" int buddy_reflect_fold(const float* dxp, int batch, int N, int L, const float* scale_b, float* dx,
                                  int64_t dx_ld, void* stream) {
  reflect_fold_kernel<<<dim3(grid1(N, 256, 64), batch), 256, 0, STREAM>>>(dxp, N, L, scale_b, dx, dx_ld);
  LAUNCH_END("reflect_fold_kernel");
}
extern "C" int buddy_comp_loss(const float* Y, const float* X, int batch, int64_t bins_times_frames, int frames,
                               float compression, float weight, double* loss, float* grad, void* stream) {
  int e = check_cuda(cudaMemsetAsync(loss, 0, sizeof(double) * batch, STREAM), "memset loss");
  if (e) return e;
  comp_loss_kernel<<<dim3(grid1(bins_times_frames, 256, 64), batch), 256, 0, STREAM>>>(
      reinterpret_cast<const float2*>(Y), reinterpret_cast<const float2*>(X), bins_times_frames, compression,
      weight / static_cast<float>(frames), loss, reinterpret_cast<float2*>(grad));
  LAUNCH_END("comp_loss_kernel");
}
extern "C" int buddy_row_stats(const float* x, int64_t ld, int batch, int n, double* out, void* stream) {
  int e = check_cuda(cudaMemsetAsync(out, 0, sizeof(double) * 2 * batch, STREAM), "memset row_stats");
  if (e) return e;
  row_stats_kernel<<<dim3(grid1(n, 256, 32), batch), 256, 0, STREAM>>>(x, ld, n, out);
  LAUNCH_END("row_stats_kernel");
}

// L = 256 * 2^log2_n2.  work: complex scratch [batch][L].  mode 0: work <- spectrum of x (permuted order, reusable
// as `H`); mode 1: y = real(ifft(fft(x) * H))[:n_out]; mode 2: same with conj(H) (the adjoint).
extern "C" int buddy_fftconv(const float* x, int64_t x_ld, int batch, int n_in, int log2_n2, const float* tw512,
                             float* work, const float* Hs, int64_t h_batch_stride, int mode, float* y, int64_t y_ld,
                             int n_out, void* stream) {
  if (log2_n2 != 9 && log2_n2 != 8 && log2_n2 != 7) {
    set_last_error("buddy_fftconv: FFT length 256*2^%d unsupported (need 2^15..2^17)", log2_n2);
    return BUDDY_ERR_UNSUPPORTED;
  }
  const int N2 = 1 << log2_n2;
  if (n_in > kN1 * N2 || n_out > kN1 * N2) {
    set_last_error("buddy_fftconv: signal longer than the FFT");
    return BUDDY_ERR_INVALID;
  }
  const float2* tw = reinterpret_cast<const float2*>(tw512);
  float2* W = reinterpret_cast<float2*>(work);
  const float2* H = reinterpret_cast<const float2*>(Hs);
  dim3 gc(N2 / kColsPerBlock, batch), gr(kN1, batch);
  const int tws = 512 >> log2_n2;
#define DISPATCH(L2)                                                                                  \
  fftconv_cols_fwd_kernel<L2><<<gc, 256, 0, STREAM>>>(x, x_ld, n_in, tw, W);                          \
  g_launches.fetch_add(1, std::memory_order_relaxed);                                                 \
  fftconv_rows_kernel<L2><<<gr, 256, 0, STREAM>>>(W, H, h_batch_stride / 2, tw, tws, mode);           \
  g_launches.fetch_add(1, std::memory_order_relaxed);                                                 \
  if (mode != 0) {                                                                                    \
    fftconv_cols_inv_kernel<L2><<<gc, 256, 0, STREAM>>>(W, tw, n_out, y, y_ld);                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                               \
  }
  if (log2_n2 == 9) {
    DISPATCH(9)
  } else if (log2_n2 == 8) {
    DISPATCH(8)
  } else {
    DISPATCH(7)
  }
#undef DISPATCH
  BUDDY_CHECK_LAUNCH("fftconv kernels");
  return 0;
}
