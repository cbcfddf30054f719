// HBM-bound glue kernels of the score network (channels-last fp32 activations, fp16 tensor-core operands):
//   GroupNorm statistics / apply(+SiLU)(+nearest-x2 | 2x2-mean resample)(+virtual channel concat) and their
//   data-gradients, 2-channel im2col / col2im for the thin convolutions, input/output pyramid helpers,
//   softmax forward/backward and fp16 transposes for the bottleneck attention.
// Reference semantics: nn.GroupNorm(eps=1e-6) + SiLU (networks/ncsnpp_utils/layerspp.py:219-263),
// naive_up/downsample_2d (up_or_down_sampling.py:59-69), Combine (layerspp.py:52-59),
// AttnBlockpp softmax (layerspp.py:81-85), pyramid up/down (layerspp.py:117,156).
// All 128-bit loads/stores; one CTA never crosses an image (grid.y = batch index).
#include <cuda_fp8.h>

#include <atomic>
#include <cstdlib>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {
extern std::atomic<long long> g_launches;

#define LAUNCH_END(name)                                  \
  g_launches.fetch_add(1, std::memory_order_relaxed);     \
  BUDDY_CHECK_LAUNCH(name);                               \
  return 0;

// SiLU and its derivative with MUFU.EX2 + MUFU.RCP (no IEEE-division slow path): <= 2 ulp, far inside the fp32 budget
__device__ __forceinline__ float silu_f(float z) { return __fdividef(z, 1.f + __expf(-z)); }
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = __fdividef(1.f, 1.f + __expf(-z));
  return s * fmaf(z, 1.f - s, 1.f);
}
// p / d for p * d < 2^32 with magic = ceil(2^32 / d) (d >= 2; magic 0 encodes d == 1)
__device__ __forceinline__ uint32_t fast_div(uint32_t p, uint32_t magic) { return magic ? __umulhi(p, magic) : p; }

struct GnSrc {
  const float* xa;
  const float* xb;
  int Ca, Cb;
  const double* sa;  // bundle sums [B][Ca/4][2]
  const double* sb;  // bundle sums [B][Cb/4][2]
};

// per-group mean / rstd of image b from the per-4-channel-bundle (sum, sumsq) accumulators
__device__ __forceinline__ void group_stats_to_smem(float* s_mean, float* s_rstd, const GnSrc& s, int b, int G,
                                                     int cpg, double n, float eps) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double S = 0.0, Q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; c += 4) {
      const double* p = (c < s.Ca) ? s.sa + (static_cast<long long>(b) * (s.Ca >> 2) + (c >> 2)) * 2
                                   : s.sb + (static_cast<long long>(b) * (s.Cb >> 2) + ((c - s.Ca) >> 2)) * 2;
      S += p[0];
      Q += p[1];
    }
    const double m = S / n;
    double var = Q / n - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(m);
    s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

__device__ __forceinline__ void load8(const GnSrc& s, long long pix, int c, float (&v)[8]) {
  const float* p = (c < s.Ca) ? s.xa + pix * s.Ca + c : s.xb + pix * s.Cb + (c - s.Ca);
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8h(__half* p, const float (&v)[8]) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  uint4 q;
  q.x = *reinterpret_cast<uint32_t*>(&h0);
  q.y = *reinterpret_cast<uint32_t*>(&h1);
  q.z = *reinterpret_cast<uint32_t*>(&h2);
  q.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = q;
}

__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d);
// fp16 operand store; split: row = [hi (C) | lo (C)], lo = fp16(v - float(hi)); split 2: fp16 hi + e4m3 pair (base8)
__device__ __forceinline__ void store_op8(__half* base, long long pix, int C, int c, int split, const float (&v)[8],
                                          uint8_t* base8 = nullptr) {
  if (split == 2 && base8) {
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __half2float(__float2half_rn(v[j]));
      lo[j] = (v[j] - hi[j]) * 512.f;
    }
    store8h(base + pix * C + c, hi);
    uint8_t* r8 = base8 + pix * (2 * C);
    *reinterpret_cast<uint2*>(r8 + c) =
        make_uint2(pack4_e4m3(lo[0], lo[1], lo[2], lo[3]), pack4_e4m3(lo[4], lo[5], lo[6], lo[7]));
    *reinterpret_cast<uint2*>(r8 + C + c) =
        make_uint2(pack4_e4m3(hi[0], hi[1], hi[2], hi[3]), pack4_e4m3(hi[4], hi[5], hi[6], hi[7]));
    return;
  }
  if (!split || split == 2) {
    store8h(base + pix * C + c, v);
    return;
  }
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __half2float(__float2half_rn(v[j]));
    lo[j] = v[j] - hi[j];
  }
  store8h(base + pix * (2 * C) + c, hi);
  store8h(base + pix * (2 * C) + C + c, lo);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm bundle statistics of a plain fp32 [B][P][C] tensor (for tensors no conv epilogue produced)
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const int nb = C >> 2;  // bundles per pixel; blockDim.x is a multiple of nb
  const int bundle = threadIdx.x % nb;
  const long long ppb = blockDim.x / nb;  // pixels per block-iteration
  float s = 0.f, q = 0.f;
  const float* xb = x + static_cast<long long>(b) * P * C;
  for (long long pix = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / nb; pix < P;
       pix += static_cast<long long>(gridDim.x) * ppb) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb + pix * C) + bundle);
    s += v.x + v.y + v.z + v.w;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  extern __shared__ float sm[];  // [blockDim][2]
  sm[threadIdx.x * 2] = s;
  sm[threadIdx.x * 2 + 1] = q;
  __syncthreads();
  if (threadIdx.x < nb) {
    double S = 0.0, Q = 0.0;
    for (int t = threadIdx.x; t < blockDim.x; t += nb) {
      S += sm[t * 2];
      Q += sm[t * 2 + 1];
    }
    double* o = stats + (static_cast<long long>(b) * nb + bundle) * 2;
    atomicAdd(o, S);
    atomicAdd(o + 1, Q);
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU) (+resample) -> fp16 operand; optional raw fp16 copy of x for the 1x1 skip conv
// Thread = one fixed 4-channel bundle, looping over pixels (two per iteration, loads issued first): all per-channel
// constants (rstd*gamma, beta - mean*rstd*gamma) are hoisted into registers, the inner loop is FMA + SiLU + stores.
// ------------------------------------------------------------------------------------------------
struct GnApplyArgs {
  GnSrc s;
  const float* gamma;
  const float* beta;
  int H, W;  // resolution of x
  int G, cpg;
  float eps;
  int silu;
  int mode;  // 0 none, 1 nearest x2 up, 2 2x2-mean down
  __half* out;
  __half* out_raw;
  int split;
  uint8_t* out8;
  uint8_t* out_raw8;
  uint32_t magic;  // fast_div constant for the row length the pixel index is split by (modes 1, 2)
};

__device__ __forceinline__ float4 ld4(const GnSrc& s, long long pix, int c) {
  const float* p = (c < s.Ca) ? s.xa + pix * s.Ca + c : s.xb + pix * s.Cb + (c - s.Ca);
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ uint2 pack4h(float a, float b, float c, float d) {
  __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  uint2 q;
  q.x = *reinterpret_cast<uint32_t*>(&h0);
  q.y = *reinterpret_cast<uint32_t*>(&h1);
  return q;
}
__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}
// fp16 operand store of 4 channels.  split 1: row = [hi (C) | lo (C)] fp16.
// split 2: fp16 hi in `base` (row C) + e4m3 pair in `base8` (row 2C bytes) = [e4m3(lo * 2^9) | e4m3(hi)].
__device__ __forceinline__ void store_op4(__half* base, uint8_t* base8, size_t pix, int C, int c, int split,
                                          float4 v) {
  if (split == 2 && base8) {
    const float hx = __half2float(__float2half_rn(v.x)), hy = __half2float(__float2half_rn(v.y));
    const float hz = __half2float(__float2half_rn(v.z)), hw = __half2float(__float2half_rn(v.w));
    *reinterpret_cast<uint2*>(base + pix * C + c) = pack4h(hx, hy, hz, hw);
    uint8_t* r8 = base8 + pix * (2 * C);
    *reinterpret_cast<uint32_t*>(r8 + c) =
        pack4_e4m3((v.x - hx) * 512.f, (v.y - hy) * 512.f, (v.z - hz) * 512.f, (v.w - hw) * 512.f);
    *reinterpret_cast<uint32_t*>(r8 + C + c) = pack4_e4m3(hx, hy, hz, hw);
    return;
  }
  if (!split || split == 2) {   // split 2 without an fp8 buffer: the consumer runs a single fp16 pass
    *reinterpret_cast<uint2*>(base + pix * C + c) = pack4h(v.x, v.y, v.z, v.w);
    return;
  }
  const float hx = __half2float(__float2half_rn(v.x)), hy = __half2float(__float2half_rn(v.y));
  const float hz = __half2float(__float2half_rn(v.z)), hw = __half2float(__float2half_rn(v.w));
  *reinterpret_cast<uint2*>(base + pix * (2 * C) + c) = pack4h(hx, hy, hz, hw);
  *reinterpret_cast<uint2*>(base + pix * (2 * C) + C + c) = pack4h(v.x - hx, v.y - hy, v.z - hz, v.w - hw);
}
__device__ __forceinline__ float4 act4(float4 x, const float (&sc)[4], const float (&sh)[4], int silu) {
  float4 y = make_float4(fmaf(x.x, sc[0], sh[0]), fmaf(x.y, sc[1], sh[1]), fmaf(x.z, sc[2], sh[2]),
                         fmaf(x.w, sc[3], sh[3]));
  if (silu) {
    y.x = silu_f(y.x);
    y.y = silu_f(y.y);
    y.z = silu_f(y.z);
    y.w = silu_f(y.w);
  }
  return y;
}

__global__ void __launch_bounds__(256, 3) gn_apply_kernel(const GnApplyArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.y;
  const int C = a.s.Ca + a.s.Cb;
  group_stats_to_smem(s_mean, s_rstd, a.s, b, a.G, a.cpg, static_cast<double>(a.cpg) * a.H * a.W, a.eps);
  __syncthreads();
  const int c4n = C >> 2;  // blockDim.x is a multiple of c4n
  const int c = (threadIdx.x % c4n) * 4;
  const uint32_t lane_p = threadIdx.x / c4n;
  const uint32_t ppb = blockDim.x / c4n;
  float sc[4], sh[4];
  {
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + c));
    const float gv[4] = {g.x, g.y, g.z, g.w}, bv[4] = {be.x, be.y, be.z, be.w};
    const int grp = c / a.cpg;  // a 4-channel bundle never straddles a group (cpg % 4 == 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc[j] = s_rstd[grp] * gv[j];
      sh[j] = bv[j] - s_mean[grp] * sc[j];
    }
  }
  const uint32_t Ho = (a.mode == 1) ? a.H * 2 : (a.mode == 2 ? a.H / 2 : a.H);
  const uint32_t Wo = (a.mode == 1) ? a.W * 2 : (a.mode == 2 ? a.W / 2 : a.W);
  const uint32_t Pin = static_cast<uint32_t>(a.H) * a.W, Pout = Ho * Wo;
  const uint32_t Pwork = (a.mode == 2) ? Pout : Pin;
  // per-thread image bases (all further indexing is 32-bit pixel index * row length)
  const bool in_a = c < a.s.Ca;
  const int Cl = in_a ? a.s.Ca : a.s.Cb;
  const float* xin = (in_a ? a.s.xa + static_cast<size_t>(b) * Pin * Cl + c
                           : a.s.xb + static_cast<size_t>(b) * Pin * Cl + (c - a.s.Ca));
  const int row16 = (a.split == 1) ? 2 * C : C;
  __half* o16 = a.out + static_cast<size_t>(b) * Pout * row16;
  __half* r16 = a.out_raw ? a.out_raw + static_cast<size_t>(b) * Pout * row16 : nullptr;
  uint8_t* o8 = a.out8 ? a.out8 + static_cast<size_t>(b) * Pout * (2 * C) : nullptr;
  uint8_t* r8 = a.out_raw8 ? a.out_raw8 + static_cast<size_t>(b) * Pout * (2 * C) : nullptr;
  const uint32_t stride = gridDim.x * ppb;
  const uint32_t pstart = blockIdx.x * ppb + lane_p;
#define LDX(p) __ldg(reinterpret_cast<const float4*>(xin + static_cast<size_t>(p) * Cl))
  if (a.mode == 0) {
    for (uint32_t p0 = pstart; p0 < Pwork; p0 += 4 * stride) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t p = p0 + u * stride;
        x[u] = (p < Pwork) ? LDX(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t p = p0 + u * stride;
        if (p < Pwork) {
          store_op4(o16, o8, p, C, c, a.split, act4(x[u], sc, sh, a.silu));
          if (r16) store_op4(r16, r8, p, C, c, a.split, x[u]);
        }
      }
    }
  } else if (a.mode == 1) {
    for (uint32_t p0 = pstart; p0 < Pwork; p0 += 2 * stride) {
      const uint32_t p1 = p0 + stride;
      const bool has1 = p1 < Pwork;
      const float4 x0 = LDX(p0);
      const float4 x1 = has1 ? LDX(p1) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !has1) break;
        const uint32_t p = u ? p1 : p0;
        const float4 x = u ? x1 : x0;
        const float4 y = act4(x, sc, sh, a.silu);
        const uint32_t h = fast_div(p, a.magic), w = p - h * a.W;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const uint32_t po = (2 * h + (d >> 1)) * Wo + (2 * w + (d & 1));
          store_op4(o16, o8, po, C, c, a.split, y);
          if (r16) store_op4(r16, r8, po, C, c, a.split, x);
        }
      }
    }
  } else {
    for (uint32_t p0 = pstart; p0 < Pwork; p0 += 2 * stride) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t p = p0 + u * stride;
        if (p >= Pwork) break;
        const uint32_t ho = fast_div(p, a.magic), wo = p - ho * Wo;
        float4 xs[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) xs[d] = LDX((2 * ho + (d >> 1)) * a.W + (2 * wo + (d & 1)));
        float4 ya = make_float4(0.f, 0.f, 0.f, 0.f), xa = ya;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 y = act4(xs[d], sc, sh, a.silu);
          ya.x += y.x; ya.y += y.y; ya.z += y.z; ya.w += y.w;
          xa.x += xs[d].x; xa.y += xs[d].y; xa.z += xs[d].z; xa.w += xs[d].w;
        }
        ya = make_float4(ya.x * 0.25f, ya.y * 0.25f, ya.z * 0.25f, ya.w * 0.25f);
        xa = make_float4(xa.x * 0.25f, xa.y * 0.25f, xa.z * 0.25f, xa.w * 0.25f);
        store_op4(o16, o8, p, C, c, a.split, ya);
        if (r16) store_op4(r16, r8, p, C, c, a.split, xa);
      }
    }
  }
#undef LDX
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU)(+resample) backward w.r.t. the input.
//   da : fp32 gradient w.r.t. the activation output, at the conv's resolution, C = Ca + Cb channels
//   pass 1 (gn_bwd_stats): per (image, group) S1 = sum dxh, S2 = sum dxh*xh   (dxh = dz*gamma)
//   pass 2 (gn_bwd)      : dx = rstd*(dxh - S1/n - xh*S2/n) + R^T(dskip)*skip_scale + extra
// Thread = one fixed 4-channel bundle (constants hoisted), looping over pixels.
// ------------------------------------------------------------------------------------------------
struct GnBwdArgs {
  GnSrc s;
  const float* gamma;
  const float* beta;
  int H, W;  // resolution of x
  int G, cpg;
  float eps;
  int silu;
  int mode;
  const float* da;     // [B][Hc][Wc][C]
  const float* dskip;  // [B][Hc][Wc][C] or null
  float skip_scale;
  const float* extra_a;  // [B][H][W][Ca] or null
  const float* extra_b;  // [B][H][W][Cb] or null
  double* gsum;          // [B][G][2]
  float* dxa;            // [B][H][W][Ca] or null
  float* dxb;            // [B][H][W][Cb] or null
  __half* g16a;          // fp16(dx * g16_scale) or null
  __half* g16b;
  float g16_scale;
  int split;
  uint8_t* g8a;
  uint8_t* g8b;
  uint32_t magic;  // fast_div constant for W (modes 1, 2)
};

// gradient w.r.t. the activation output pulled back through the resample, for x-pixel p = h*W + w of one image
// (t points at the image's first pixel of the conv-resolution tensor, already offset to channel c; row length C)
__device__ __forceinline__ float4 pull_back4(const float* __restrict__ t, int mode, uint32_t p, uint32_t h, uint32_t w,
                                             uint32_t W, int C) {
  if (mode == 0) {
    return __ldg(reinterpret_cast<const float4*>(t + static_cast<size_t>(p) * C));
  } else if (mode == 1) {  // forward was nearest x2: sum the 4 children
    const uint32_t Wc = 2 * W;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          t + static_cast<size_t>((2 * h + (d >> 1)) * Wc + 2 * w + (d & 1)) * C));
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    return g;
  } else {  // forward was 2x2 mean: a quarter of the parent
    const uint32_t Wc = W / 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(t + static_cast<size_t>((h >> 1) * Wc + (w >> 1)) * C));
    return make_float4(v.x * 0.25f, v.y * 0.25f, v.z * 0.25f, v.w * 0.25f);
  }
}

template <bool kPass2>
__global__ void __launch_bounds__(256, kPass2 ? 3 : 4) gn_bwd_kernel(const GnBwdArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ float s_part[kPass2 ? 1 : 256][2];   // pass 0: per-thread partials, reduced in a fixed order
  const int b = blockIdx.y;
  const int C = a.s.Ca + a.s.Cb;
  const double n = static_cast<double>(a.cpg) * a.H * a.W;
  group_stats_to_smem(s_mean, s_rstd, a.s, b, a.G, a.cpg, n, a.eps);
  __syncthreads();
  const int c4n = C >> 2;  // blockDim.x is a multiple of c4n -> each thread keeps one channel bundle
  const int c = (threadIdx.x % c4n) * 4;
  const int grp = c / a.cpg;
  const uint32_t P = static_cast<uint32_t>(a.H) * a.W;
  const uint32_t ppb = blockDim.x / c4n;
  const float rstd = s_rstd[grp], nmr = -s_mean[grp] * rstd;  // xh = x*rstd + nmr
  float gam[4], bet[4];
  {
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta + c));
    gam[0] = g.x; gam[1] = g.y; gam[2] = g.z; gam[3] = g.w;
    bet[0] = be.x; bet[1] = be.y; bet[2] = be.z; bet[3] = be.w;
  }
  float m1 = 0.f, m2 = 0.f;
  if (kPass2) {
    m1 = static_cast<float>(a.gsum[(static_cast<long long>(b) * a.G + grp) * 2] / n);
    m2 = static_cast<float>(a.gsum[(static_cast<long long>(b) * a.G + grp) * 2 + 1] / n);
  }
  const bool in_a = c < a.s.Ca;
  const int cl = in_a ? c : c - a.s.Ca;
  const int Cl = in_a ? a.s.Ca : a.s.Cb;
  // per-thread image bases, pre-offset to this thread's channel bundle
  const size_t img_l = static_cast<size_t>(b) * P * Cl;
  const float* xin = (in_a ? a.s.xa : a.s.xb) + img_l + cl;
  const uint32_t Pc = (a.mode == 1) ? 4 * P : (a.mode == 2 ? P / 4 : P);  // pixels of the conv-resolution tensors
  const float* da = a.da + static_cast<size_t>(b) * Pc * C + c;
  const float* dsk = (kPass2 && a.dskip) ? a.dskip + static_cast<size_t>(b) * Pc * C + c : nullptr;
  const float* ex = in_a ? a.extra_a : a.extra_b;
  if (ex) ex += img_l + cl;
  float* o32 = in_a ? a.dxa : a.dxb;
  if (o32) o32 += img_l + cl;
  __half* o16 = in_a ? a.g16a : a.g16b;
  uint8_t* o8 = in_a ? a.g8a : a.g8b;
  if (o16) o16 += static_cast<size_t>(b) * P * (a.split == 1 ? 2 * Cl : Cl);
  if (o8) o8 += static_cast<size_t>(b) * P * (2 * Cl);
  const uint32_t stride = gridDim.x * ppb;
  float p1 = 0.f, p2 = 0.f;
  constexpr int U = 2;
  for (uint32_t p0 = blockIdx.x * ppb + threadIdx.x / c4n; p0 < P; p0 += U * stride) {
    float4 x4[U], g4[U], sk4[U], e4[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = p0 + u * stride;
      ok[u] = p < P;
      x4[u] = g4[u] = sk4[u] = e4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        uint32_t h = 0, w = 0;
        if (a.mode != 0) {
          h = fast_div(p, a.magic);
          w = p - h * a.W;
        }
        x4[u] = __ldg(reinterpret_cast<const float4*>(xin + static_cast<size_t>(p) * Cl));
        g4[u] = pull_back4(da, a.mode, p, h, w, a.W, C);
        if (kPass2) {
          if (dsk) sk4[u] = pull_back4(dsk, a.mode, p, h, w, a.W, C);
          if (ex) e4[u] = __ldg(reinterpret_cast<const float4*>(ex + static_cast<size_t>(p) * Cl));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const uint32_t p = p0 + u * stride;
      const float xv[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w}, gv[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
      const float sv[4] = {sk4[u].x, sk4[u].y, sk4[u].z, sk4[u].w}, ev[4] = {e4[u].x, e4[u].y, e4[u].z, e4[u].w};
      float dx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = fmaf(xv[j], rstd, nmr);
        float dz = gv[j];
        if (a.silu) dz *= dsilu_f(fmaf(xh, gam[j], bet[j]));
        const float dxh = dz * gam[j];
        if (!kPass2) {
          p1 += dxh;
          p2 = fmaf(dxh, xh, p2);
        } else {
          dx[j] = fmaf(rstd, dxh - m1 - xh * m2, fmaf(sv[j], a.skip_scale, ev[j]));
        }
      }
      if (kPass2) {
        if (o32) *reinterpret_cast<float4*>(o32 + static_cast<size_t>(p) * Cl) = make_float4(dx[0], dx[1], dx[2], dx[3]);
        if (o16)
          store_op4(o16, o8, p, Cl, cl, a.split,
                    make_float4(dx[0] * a.g16_scale, dx[1] * a.g16_scale, dx[2] * a.g16_scale, dx[3] * a.g16_scale));
      }
    }
  }
  if (!kPass2) {
    // CTA reduction in a fixed order (no floating-point atomics in shared memory): thread g sums the partials of
    // group g's bundles over the CTA's pixel lanes; the fp64 global atomics that follow add fp32 values almost
    // exactly, so the statistics — and with them every fp16 rounding decision downstream — repeat run to run.
    s_part[threadIdx.x][0] = p1;
    s_part[threadIdx.x][1] = p2;
    __syncthreads();
    const int bpg = a.cpg >> 2;  // bundles (threads) per group within one pixel lane
    for (int g = threadIdx.x; g < a.G; g += blockDim.x) {
      float s1 = 0.f, s2 = 0.f;
      for (uint32_t lp = 0; lp < ppb; ++lp)
        for (int j = 0; j < bpg; ++j) {
          const int t = lp * c4n + g * bpg + j;
          s1 += s_part[t][0];
          s2 += s_part[t][1];
        }
      atomicAdd(a.gsum + (static_cast<long long>(b) * a.G + g) * 2, static_cast<double>(s1));
      atomicAdd(a.gsum + (static_cast<long long>(b) * a.G + g) * 2 + 1, static_cast<double>(s2));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2-channel 3x3 im2col -> fp16 [B][H][W][64] (K index = tap*2 + ci, 18 used, rest zero); and its adjoint
// ------------------------------------------------------------------------------------------------
__global__ void im2col_c2_kernel(const float* __restrict__ x, int B, int H, int W, __half* __restrict__ col,
                                 int split, uint8_t* __restrict__ col8, float in_scale) {
  const long long P = static_cast<long long>(B) * H * W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P * 8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i >> 3;
    const int chunk = static_cast<int>(i & 7);  // 8 fp16 = 4 taps x 2 channels
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const long long b = p / (static_cast<long long>(W) * H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tap = chunk * 4 + j;
      float v0 = 0.f, v1 = 0.f;
      if (tap < 9) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          const float2 q = __ldg(reinterpret_cast<const float2*>(x) + (b * H + hh) * W + ww);
          v0 = q.x;
          v1 = q.y;
        }
      }
      v[2 * j] = v0 * in_scale;
      v[2 * j + 1] = v1 * in_scale;
    }
    store_op8(col, p, 64, chunk * 8, split, v, col8);
  }
}

// dx[b,h,w,ci] (+)= sum_tap dcol[b, h-dy, w-dx, tap*2+ci]   (dcol fp32, row stride ld)
__global__ void col2im_c2_kernel(const float* __restrict__ dcol, int ld, int B, int H, int W, float* __restrict__ dx,
                                 int accumulate) {
  const long long P = static_cast<long long>(B) * H * W;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const long long b = p / (static_cast<long long>(W) * H);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - (tap / 3 - 1), ww = w - (tap % 3 - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float2 q = __ldg(reinterpret_cast<const float2*>(dcol + ((b * H + hh) * W + ww) * ld + tap * 2));
        a0 += q.x;
        a1 += q.y;
      }
    }
    float2* o = reinterpret_cast<float2*>(dx) + p;
    if (accumulate) {
      const float2 old = *o;
      a0 += old.x;
      a1 += old.y;
    }
    *o = make_float2(a0, a1);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-channel helpers: 2x2 mean pool (input pyramid), nearest x2 (+add) (output pyramid) and adjoints
// ------------------------------------------------------------------------------------------------
// mode 0: out[Ho=H/2] = mean4(in)            mode 1: out[2H] = nearest(in) (+ add)
// mode 2: out[2H] (+)= in[parent]/4 (adjoint of mode 0)   mode 3: out[H/2] = sum4(in) (adjoint of mode 1)
__global__ void resample_c2_kernel(const float2* __restrict__ in, int B, int Hin, int Win, int mode,
                                   const float2* __restrict__ add, float2* __restrict__ out, int accumulate) {
  const bool shrink = (mode == 0 || mode == 3);
  const int Ho = shrink ? Hin / 2 : Hin * 2, Wo = shrink ? Win / 2 : Win * 2;
  const long long P = static_cast<long long>(B) * Ho * Wo;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(p % Wo);
    const int h = static_cast<int>((p / Wo) % Ho);
    const long long b = p / (static_cast<long long>(Wo) * Ho);
    float2 r;
    if (shrink) {
      float sx = 0.f, sy = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float2 q = __ldg(in + (b * Hin + 2 * h + (d >> 1)) * Win + 2 * w + (d & 1));
        sx += q.x;
        sy += q.y;
      }
      const float sc = (mode == 0) ? 0.25f : 1.f;
      r = make_float2(sx * sc, sy * sc);
    } else {
      const float2 q = __ldg(in + (b * Hin + (h >> 1)) * Win + (w >> 1));
      const float sc = (mode == 2) ? 0.25f : 1.f;
      r = make_float2(q.x * sc, q.y * sc);
    }
    if (add) {
      const float2 q = __ldg(add + p);
      r.x += q.x;
      r.y += q.y;
    }
    if (accumulate) {
      r.x += out[p].x;
      r.y += out[p].y;
    }
    out[p] = r;
  }
}

// Combine (layerspp.py:52-59, method 'sum'): out[p][c] = h[p][c] + w[c][0]*pyr[p][0] + w[c][1]*pyr[p][1] + bias[c]
__global__ void combine_fwd_kernel(const float* __restrict__ h, const float2* __restrict__ pyr,
                                   const float* __restrict__ w, const float* __restrict__ bias, long long P, int C,
                                   float* __restrict__ out) {
  const int c4n = C >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P * c4n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / c4n;
    const int c = static_cast<int>(i - p * c4n) * 4;
    const float2 q = __ldg(pyr + p);
    float4 v = __ldg(reinterpret_cast<const float4*>(h + p * C + c));
    const float4 w01 = __ldg(reinterpret_cast<const float4*>(w + 2 * c));      // w[c][0], w[c][1], w[c+1][0], ...
    const float4 w23 = __ldg(reinterpret_cast<const float4*>(w + 2 * c + 4));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    v.x += w01.x * q.x + w01.y * q.y + bb.x;
    v.y += w01.z * q.x + w01.w * q.y + bb.y;
    v.z += w23.x * q.x + w23.y * q.y + bb.z;
    v.w += w23.z * q.x + w23.w * q.y + bb.w;
    *reinterpret_cast<float4*>(out + p * C + c) = v;
  }
}
// adjoint w.r.t. the pyramid input: dpyr[p][i] = sum_c dout[p][c] * w[c][i]    (one warp per pixel)
__global__ void combine_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ w, long long P, int C,
                                   float2* __restrict__ dpyr) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long p = warp; p < P; p += nwarps) {
    float a0 = 0.f, a1 = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dout + p * C + c));
      const float4 w01 = __ldg(reinterpret_cast<const float4*>(w + 2 * c));
      const float4 w23 = __ldg(reinterpret_cast<const float4*>(w + 2 * c + 4));
      a0 += v.x * w01.x + v.y * w01.z + v.z * w23.x + v.w * w23.z;
      a1 += v.x * w01.y + v.y * w01.w + v.z * w23.y + v.w * w23.w;
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) dpyr[p] = make_float2(a0, a1);
  }
}

// y[p][:] = M (2x2) x[p][:] + bias   on 2-channel pixels (output_layer 1x1 conv, ncsnpp.py:113,445; and its adjoint)
__global__ void affine_c2_kernel(const float2* __restrict__ x, long long P, float m00, float m01, float m10, float m11,
                                 float b0, float b1, float2* __restrict__ y) {
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 q = __ldg(x + p);
    y[p] = make_float2(m00 * q.x + m01 * q.y + b0, m10 * q.x + m11 * q.y + b1);
  }
}

// ------------------------------------------------------------------------------------------------
// attention helpers
// ------------------------------------------------------------------------------------------------
// row softmax: logits fp32 [rows][n] -> P fp16 [rows][ldp]   (one block per row)
__global__ void softmax_fwd_kernel(const float* __restrict__ s, int n, __half* __restrict__ p, int ldp) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* sr = s + row * n;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, sr[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
    v = warp_max(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  m = red[0];
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += expf(sr[i] - m);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float inv = 1.f / red[0];
  __half* pr = p + row * ldp;
  for (int i = threadIdx.x; i < n; i += blockDim.x) pr[i] = __float2half_rn(expf(sr[i] - m) * inv);
}
// dS = P * (dP - sum_j dP*P) * scale   -> fp16 [rows][ld]
__global__ void softmax_bwd_kernel(const __half* __restrict__ p, int ldp, const float* __restrict__ dp, int n,
                                   float scale, __half* __restrict__ ds, int ldds) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const __half* pr = p + row * ldp;
  const float* dr = dp + row * n;
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dot += __half2float(pr[i]) * dr[i];
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  dot = red[0];
  __half* o = ds + row * ldds;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    o[i] = __float2half_rn(__half2float(pr[i]) * (dr[i] - dot) * scale);
}
// batched fp16 transpose: in [batch][R][ld_in] (first Cc columns) -> out [batch][Cc][ld_out] (first R columns)
__global__ void transpose_h_kernel(const __half* __restrict__ in, int R, int Cc, long long ld_in, long long bs_in,
                                   __half* __restrict__ out, long long ld_out, long long bs_out) {
  __shared__ __half tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const __half* ip = in + b * bs_in;
  __half* op = out + b * bs_out;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? ip[r * ld_in + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) op[c * ld_out + r] = tile[threadIdx.x][i];
  }
}
// fp32 -> fp16 with scale (contiguous)
__global__ void cast_scale_h_kernel(const float* __restrict__ x, long long n8, float scale, __half* __restrict__ y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= scale;
    store8h(y + i * 8, v);
  }
}

static int grid_for(long long items, int threads, int max_blocks = 148 * 16) {
  long long g = (items + threads - 1) / threads;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace buddy

using namespace buddy;

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int buddy_gn_stats(const float* x, int B, int64_t P, int C, double* stats, void* stream) {
  if (C % 4 || C <= 0 || C > 1024) {
    set_last_error("buddy_gn_stats: C must be a multiple of 4 (got %d)", C);
    return BUDDY_ERR_UNSUPPORTED;
  }
  const int nb = C / 4;
  int threads = nb * (256 / nb > 0 ? 256 / nb : 1);
  if (threads > 1024) {
    set_last_error("buddy_gn_stats: C too large");
    return BUDDY_ERR_UNSUPPORTED;
  }
  const long long ppb = threads / nb;
  long long gx = (P + ppb * 8 - 1) / (ppb * 8);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  gn_stats_kernel<<<dim3((unsigned)gx, B), threads, threads * 2 * sizeof(float), STREAM>>>(x, P, C, stats);
  LAUNCH_END("gn_stats_kernel");
}

// pixels per thread of the GroupNorm kernels (tuning knob: BUDDY_GN_PPT, default 32)
static int gn_ppt() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("BUDDY_GN_PPT");
    v = e ? atoi(e) : 32;
    if (v < 1 || v > 1024) v = 32;
  }
  return v;
}

// CTAs per image for the GroupNorm kernels: ~gn_ppt() pixels per thread on big images, but never fewer than ~2 x 148 CTAs
// per image — at the coarse levels (32 x 66 pixels) a 32-pixel-per-thread grid is 17 CTAs of serial, latency-bound
// loads (ncu, B = 1: 20-28 us per launch for 2 MB of data).  Depends on the image only, never on the batch, so the
// grouping of the partial sums — and with it every bit of the statistics — is the same for an utterance run alone or in
// a batch.
static long long gn_grid(long long P, long long ppb) {
  static const long long target = getenv("BUDDY_GN_CTAS") ? atoll(getenv("BUDDY_GN_CTAS")) : 296;
  long long ppt = P / (ppb * target);
  if (ppt < 2) ppt = 2;
  if (ppt > gn_ppt()) ppt = gn_ppt();
  long long gx = (P + ppb * ppt - 1) / (ppb * ppt);
  if (gx > 148 * 16) gx = 148 * 16;
  if (gx < 1) gx = 1;
  return gx;
}

static int check_gn(int Ca, int Cb, int G, const char* who) {
  const int C = Ca + Cb;
  if (Ca % 8 || Cb % 8 || C <= 0 || C > 1024 || G <= 0 || G > 32 || C % G || (C / G) % 4) {
    set_last_error("%s: unsupported channel/group configuration Ca=%d Cb=%d G=%d", who, Ca, Cb, G);
    return BUDDY_ERR_UNSUPPORTED;
  }
  return 0;
}

extern "C" int buddy_gn_apply(const buddy_gn_desc* d, void* stream) {
  int e = check_gn(d->Ca, d->Cb, d->groups, "buddy_gn_apply");
  if (e) return e;
  if (d->mode == 2 && ((d->H & 1) || (d->W & 1))) {
    set_last_error("buddy_gn_apply: 2x2-mean needs even H, W");
    return BUDDY_ERR_UNSUPPORTED;
  }
  GnApplyArgs a;
  a.s = {d->xa, d->xb, d->Ca, d->Cb, d->stats_a, d->stats_b};
  a.gamma = d->gamma;
  a.beta = d->beta;
  a.H = d->H;
  a.W = d->W;
  a.G = d->groups;
  a.cpg = (d->Ca + d->Cb) / d->groups;
  a.eps = d->eps;
  a.silu = d->silu;
  a.mode = d->mode;
  a.out = static_cast<__half*>(d->out);
  a.out_raw = static_cast<__half*>(d->out_raw);
  a.split = d->split;
  a.out8 = static_cast<uint8_t*>(d->out8);
  a.out_raw8 = static_cast<uint8_t*>(d->out_raw8);
  {
    const long long rowlen = d->mode == 2 ? d->W / 2 : d->W;
    if (static_cast<long long>(d->H) * d->W * (d->mode == 1 ? 4 : 1) >= (1LL << 31) ||
        static_cast<long long>(d->H) * d->W * rowlen >= (1LL << 32)) {
      set_last_error("buddy_gn_apply: image too large for 32-bit pixel indexing (H %d W %d)", d->H, d->W);
      return BUDDY_ERR_UNSUPPORTED;
    }
    a.magic = rowlen >= 2 ? static_cast<uint32_t>(((1ULL << 32) + rowlen - 1) / rowlen) : 0u;
  }
  const int c4n = (d->Ca + d->Cb) / 4;
  const int threads = c4n * (256 / c4n > 0 ? 256 / c4n : 1);
  const long long ppb = threads / c4n;
  const long long Pw = static_cast<long long>(d->mode == 2 ? (d->H / 2) * (d->W / 2) : d->H * d->W);
  const long long gx = gn_grid(Pw, ppb);
  gn_apply_kernel<<<dim3((unsigned)gx, d->batch), threads, 0, STREAM>>>(a);
  LAUNCH_END("gn_apply_kernel");
}

extern "C" int buddy_gn_bwd(const buddy_gn_desc* d, const buddy_gn_bwd_desc* g, void* stream) {
  int e = check_gn(d->Ca, d->Cb, d->groups, "buddy_gn_bwd");
  if (e) return e;
  const int C = d->Ca + d->Cb;
  GnBwdArgs a;
  a.s = {d->xa, d->xb, d->Ca, d->Cb, d->stats_a, d->stats_b};
  a.gamma = d->gamma;
  a.beta = d->beta;
  a.H = d->H;
  a.W = d->W;
  a.G = d->groups;
  a.cpg = C / d->groups;
  a.eps = d->eps;
  a.silu = d->silu;
  a.mode = d->mode;
  a.da = g->da;
  a.dskip = g->dskip;
  a.skip_scale = g->skip_scale;
  a.extra_a = g->extra_a;
  a.extra_b = g->extra_b;
  a.gsum = g->gsum;
  a.dxa = g->dxa;
  a.dxb = g->dxb;
  a.g16a = static_cast<__half*>(g->g16a);
  a.g16b = static_cast<__half*>(g->g16b);
  a.g16_scale = g->g16_scale;
  a.split = d->split;
  a.g8a = static_cast<uint8_t*>(g->g8a);
  a.g8b = static_cast<uint8_t*>(g->g8b);
  if (static_cast<long long>(d->H) * d->W * (d->mode == 1 ? 4 : 1) >= (1LL << 31) ||
      static_cast<long long>(d->H) * d->W * d->W >= (1LL << 32)) {
    set_last_error("buddy_gn_bwd: image too large for 32-bit pixel indexing (H %d W %d)", d->H, d->W);
    return BUDDY_ERR_UNSUPPORTED;
  }
  a.magic = d->W >= 2 ? static_cast<uint32_t>(((1ULL << 32) + d->W - 1) / d->W) : 0u;
  const int c4n = C / 4;
  const int threads = c4n * (256 / c4n > 0 ? 256 / c4n : 1);
  const long long ppb = threads / c4n;
  const long long P = static_cast<long long>(d->H) * d->W;
  const long long gx = gn_grid(P, ppb);
  if (!g->pass0_done) {
    e = check_cuda(cudaMemsetAsync(g->gsum, 0, sizeof(double) * 2 * d->groups * d->batch, STREAM), "memset gsum");
    if (e) return e;
    gn_bwd_kernel<false><<<dim3((unsigned)gx, d->batch), threads, 0, STREAM>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    BUDDY_CHECK_LAUNCH("gn_bwd_kernel<stats>");
  } else if (d->mode != 0 || d->xb) {
    set_last_error("buddy_gn_bwd: pass0_done needs mode 0 and a single input tensor");
    return BUDDY_ERR_INVALID;
  }
  gn_bwd_kernel<true><<<dim3((unsigned)gx, d->batch), threads, 0, STREAM>>>(a);
  LAUNCH_END("gn_bwd_kernel<apply>");
}

// GroupNorm (+SiLU) of one tensor written as fp32 (no resample, no operand packing): the `fir: True` variant puts a
// FIR resampler (upfirdn2d) between the activation and the convolution (layerspp.py:252-259), which needs the
// activation itself; the operand is made afterwards by cast_operand_kernel.
__global__ void __launch_bounds__(256) gn_act32_kernel(GnSrc s, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, long long P, int G, int cpg,
                                                       float eps, int silu, float* __restrict__ out) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.y;
  const int C = s.Ca;
  group_stats_to_smem(s_mean, s_rstd, s, b, G, cpg, static_cast<double>(cpg) * static_cast<double>(P), eps);
  __syncthreads();
  const int c4n = C >> 2;
  const long long total = P * c4n;
  const float* x = s.xa + static_cast<size_t>(b) * P * C;
  float* o = out + static_cast<size_t>(b) * P * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const long long p = i / c4n;
    const int grp = c / cpg;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float rs = s_rstd[grp], mu = s_mean[grp];
    const float sc[4] = {rs * g.x, rs * g.y, rs * g.z, rs * g.w};
    const float sh[4] = {be.x - mu * sc[0], be.y - mu * sc[1], be.z - mu * sc[2], be.w - mu * sc[3]};
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C + c));
    *reinterpret_cast<float4*>(o + p * C + c) = act4(v, sc, sh, silu);
  }
}

extern "C" int buddy_im2col_c2(const float* x, int B, int H, int W, void* col, int split, void* col8, float in_scale,
                               void* stream) {
  const long long items = static_cast<long long>(B) * H * W * 8;
  im2col_c2_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(x, B, H, W, static_cast<__half*>(col), split,
                                                              static_cast<uint8_t*>(col8), in_scale);
  LAUNCH_END("im2col_c2_kernel");
}
extern "C" int buddy_col2im_c2(const float* dcol, int ld, int B, int H, int W, float* dx, int accumulate,
                               void* stream) {
  const long long items = static_cast<long long>(B) * H * W;
  col2im_c2_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(dcol, ld, B, H, W, dx, accumulate);
  LAUNCH_END("col2im_c2_kernel");
}
extern "C" int buddy_resample_c2(const float* in, int B, int Hin, int Win, int mode, const float* add, float* out,
                                 int accumulate, void* stream) {
  if (mode < 0 || mode > 3 || ((mode == 0 || mode == 3) && ((Hin & 1) || (Win & 1)))) {
    set_last_error("buddy_resample_c2: bad mode/shape");
    return BUDDY_ERR_INVALID;
  }
  const bool shrink = (mode == 0 || mode == 3);
  const long long items = static_cast<long long>(B) * (shrink ? Hin / 2 : Hin * 2) * (shrink ? Win / 2 : Win * 2);
  resample_c2_kernel<<<grid_for(items, 256), 256, 0, STREAM>>>(reinterpret_cast<const float2*>(in), B, Hin, Win, mode,
                                                               reinterpret_cast<const float2*>(add),
                                                               reinterpret_cast<float2*>(out), accumulate);
  LAUNCH_END("resample_c2_kernel");
}
extern "C" int buddy_combine_fwd(const float* h, const float* pyr, const float* w, const float* bias, int64_t P, int C,
                                 float* out, void* stream) {
  if (C % 4) {
    set_last_error("buddy_combine_fwd: C %% 4 != 0");
    return BUDDY_ERR_UNSUPPORTED;
  }
  combine_fwd_kernel<<<grid_for(P * (C / 4), 256), 256, 0, STREAM>>>(h, reinterpret_cast<const float2*>(pyr), w, bias,
                                                                      P, C, out);
  LAUNCH_END("combine_fwd_kernel");
}
extern "C" int buddy_combine_bwd(const float* dout, const float* w, int64_t P, int C, float* dpyr, void* stream) {
  if (C % 128) {
    set_last_error("buddy_combine_bwd: C %% 128 != 0");
    return BUDDY_ERR_UNSUPPORTED;
  }
  combine_bwd_kernel<<<grid_for(P * 32, 256), 256, 0, STREAM>>>(dout, w, P, C, reinterpret_cast<float2*>(dpyr));
  LAUNCH_END("combine_bwd_kernel");
}
extern "C" int buddy_affine_c2(const float* x, int64_t P, const float* m_host, const float* b_host, float* y,
                               void* stream) {
  affine_c2_kernel<<<grid_for(P, 256), 256, 0, STREAM>>>(reinterpret_cast<const float2*>(x), P, m_host[0], m_host[1],
                                                         m_host[2], m_host[3], b_host[0], b_host[1],
                                                         reinterpret_cast<float2*>(y));
  LAUNCH_END("affine_c2_kernel");
}
extern "C" int buddy_softmax_fwd(const float* s, int64_t rows, int n, void* p, int ldp, void* stream) {
  softmax_fwd_kernel<<<(unsigned)rows, 256, 0, STREAM>>>(s, n, static_cast<__half*>(p), ldp);
  LAUNCH_END("softmax_fwd_kernel");
}
extern "C" int buddy_softmax_bwd(const void* p, int ldp, const float* dp, int64_t rows, int n, float scale, void* ds,
                                 int ldds, void* stream) {
  softmax_bwd_kernel<<<(unsigned)rows, 256, 0, STREAM>>>(static_cast<const __half*>(p), ldp, dp, n, scale,
                                                         static_cast<__half*>(ds), ldds);
  LAUNCH_END("softmax_bwd_kernel");
}
extern "C" int buddy_transpose_h(const void* in, int batch, int R, int Cc, int64_t ld_in, int64_t bs_in, void* out,
                                 int64_t ld_out, int64_t bs_out, void* stream) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, batch);
  transpose_h_kernel<<<grid, dim3(32, 8), 0, STREAM>>>(static_cast<const __half*>(in), R, Cc, ld_in, bs_in,
                                                       static_cast<__half*>(out), ld_out, bs_out);
  LAUNCH_END("transpose_h_kernel");
}
// fp32 [B][H][W][C] -> tensor-core operand of scale * x (fp16, + e4m3 pair when out8 is given), optionally through a
// nearest-neighbour x2 upsampling (up = 1: out is [B][2H][2W][C]).  For the convolutions that take a RAW tensor: the
// Downsample / Upsample modules of the `resblock_type: ddpm` variant (layerspp.py:93-160) and their data-gradients.
__global__ void cast_operand_kernel(const float* __restrict__ x, int B, int H, int W, int C, int up, float scale,
                                    __half* __restrict__ out, uint8_t* __restrict__ out8, int split) {
  const int c4n = C >> 2;
  const int Ho = up ? 2 * H : H, Wo = up ? 2 * W : W;
  const long long total = static_cast<long long>(B) * Ho * Wo * c4n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const long long po = i / c4n;                       // output pixel (b, ho, wo) linear
    long long pi = po;
    if (up) {
      const int wo = static_cast<int>(po % Wo);
      const long long r = po / Wo;
      const int ho = static_cast<int>(r % Ho);
      const long long b = r / Ho;
      pi = (b * H + (ho >> 1)) * W + (wo >> 1);
    }
    float4 v = __ldg(reinterpret_cast<const float4*>(x + pi * C + c));
    v = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    store_op4(out, out8, static_cast<size_t>(po), C, c, split, v);
  }
}

extern "C" int buddy_gn_act32(const float* x, const double* stats, const float* gamma, const float* beta, int batch,
                              int64_t pixels, int C, int groups, float eps, int silu, float* out, void* stream) {
  int e = check_gn(C, 0, groups, "buddy_gn_act32");
  if (e) return e;
  if (!x || !stats || !gamma || !beta || !out || batch <= 0 || pixels <= 0) {
    set_last_error("buddy_gn_act32: invalid argument");
    return BUDDY_ERR_INVALID;
  }
  GnSrc s = {x, nullptr, C, 0, stats, nullptr};
  const long long total = pixels * (C / 4);
  long long gx = (total + 256 * 8 - 1) / (256 * 8);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  gn_act32_kernel<<<dim3((unsigned)gx, batch), 256, 0, STREAM>>>(s, gamma, beta, pixels, groups, C / groups, eps, silu,
                                                                 out);
  LAUNCH_END("gn_act32_kernel");
}

extern "C" int buddy_cast_operand(const float* x, int batch, int H, int W, int C, int upsample, float scale, void* out16,
                                  void* out8, int split, void* stream) {
  if (!x || !out16 || C % 4 || batch <= 0 || H <= 0 || W <= 0) {
    set_last_error("buddy_cast_operand: invalid argument (C %% 4 == 0 required)");
    return BUDDY_ERR_INVALID;
  }
  const long long total = static_cast<long long>(batch) * H * W * (upsample ? 4 : 1) * (C / 4);
  cast_operand_kernel<<<grid_for(total, 256), 256, 0, STREAM>>>(x, batch, H, W, C, upsample, scale,
                                                                static_cast<__half*>(out16),
                                                                static_cast<uint8_t*>(out8), split);
  LAUNCH_END("cast_operand_kernel");
}

extern "C" int buddy_cast_scale_h(const float* x, int64_t n, float scale, void* y, void* stream) {
  if (n % 8) {
    set_last_error("buddy_cast_scale_h: n %% 8 != 0");
    return BUDDY_ERR_UNSUPPORTED;
  }
  cast_scale_h_kernel<<<grid_for(n / 8, 256), 256, 0, STREAM>>>(x, n / 8, scale, static_cast<__half*>(y));
  LAUNCH_END("cast_scale_h_kernel");
}
