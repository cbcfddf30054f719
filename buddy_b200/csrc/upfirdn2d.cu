// upfirdn2d for sm_100a: upsample by zero insertion -> pad / crop -> 2-D FIR -> downsample, fp32.
//
// The reference's only native operator (StyleGAN2's upfirdn2d: networks/ncsnpp_utils/op/upfirdn2d.cpp:12-23,
// op/upfirdn2d_kernel.cu:107-207, Python glue op/upfirdn2d.py:86-139), behind the `fir=True` up/down-sampling of
// NCSN++ (up_or_down_sampling.py:195-256).  The shipped configuration has `fir: False`, and in the reference as
// published the import of the operator is commented out (up_or_down_sampling.py:10), so `fir=True` raises NameError
// there; the operator itself is provided here with the same contract:
//     in  [major][in_h][in_w][minor],  kernel [kh][kw]
//     out [major][out_h][out_w][minor], out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1 (same for w)
//     out[m][oy][ox][c] = sum_{i,j} kernel[kh-1-i][kw-1-j] * U[oy*down_y + i][ox*down_x + j][c]
//     U = zero-inserted (up) input shifted by (pad_y0, pad_x0); negative pads crop.
// HBM-bound gather: one thread per output element in output order (coalesced stores; loads hit L1/L2 — every input
// element is used by ~kh*kw/(up_x*up_y*down_x*down_y) outputs); only the taps that land on a real sample are visited
// (stride up_x / up_y through the window), the FIR lives in shared memory.
#include <atomic>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {
extern std::atomic<long long> g_launches;

constexpr int kUpfirMaxTaps = 32;

struct UpfirArgs {
  const float* in;
  const float* kernel;
  float* out;
  int major, in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int out_h, out_w;
};

// kVec = 4: four consecutive minor elements (channels of a channels-last tensor) per thread, 128-bit loads / stores
template <int kVec>
__global__ void __launch_bounds__(256) upfirdn2d_kernel(const UpfirArgs a) {
  __shared__ float kflip[kUpfirMaxTaps * kUpfirMaxTaps];
  for (int i = threadIdx.x; i < a.kh * a.kw; i += blockDim.x) {
    const int ky = i / a.kw, kx = i - ky * a.kw;
    kflip[i] = a.kernel[(a.kh - 1 - ky) * a.kw + (a.kw - 1 - kx)];
  }
  __syncthreads();
  const uint32_t mv = static_cast<uint32_t>(a.minor / kVec);          // vectors per pixel
  const uint32_t per_img = static_cast<uint32_t>(a.out_h) * a.out_w * mv;   // host guarantees < 2^31
  const long long total = static_cast<long long>(a.major) * per_img;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx / per_img;
    uint32_t r = static_cast<uint32_t>(idx - m * per_img);
    const uint32_t c = (r % mv) * kVec;
    r /= mv;
    const int ox = static_cast<int>(r % a.out_w);
    const int oy = static_cast<int>(r / a.out_w);
    // window origin in zero-inserted input coordinates (may be negative: padding)
    const int y0 = oy * a.down_y - a.pad_y0, x0 = ox * a.down_x - a.pad_x0;
    // first tap row / column that lands on a real sample: y0 + i >= 0 and (y0 + i) % up_y == 0
    const int i0 = y0 < 0 ? -y0 : (a.up_y - (y0 % a.up_y)) % a.up_y;
    const int j0 = x0 < 0 ? -x0 : (a.up_x - (x0 % a.up_x)) % a.up_x;
    const float* base = a.in + m * a.in_h * a.in_w * a.minor + c;
    float acc[kVec];
#pragma unroll
    for (int v = 0; v < kVec; ++v) acc[v] = 0.f;
    for (int i = i0; i < a.kh; i += a.up_y) {
      const int iy = (y0 + i) / a.up_y;
      if (iy >= a.in_h) break;
      for (int j = j0; j < a.kw; j += a.up_x) {
        const int ix = (x0 + j) / a.up_x;
        if (ix >= a.in_w) break;
        const float kv = kflip[i * a.kw + j];
        const float* src = base + (static_cast<long long>(iy) * a.in_w + ix) * a.minor;
        if (kVec == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(src));
          acc[0] = fmaf(kv, q.x, acc[0]);
          acc[1] = fmaf(kv, q.y, acc[1]);
          acc[2] = fmaf(kv, q.z, acc[2]);
          acc[kVec - 1] = fmaf(kv, q.w, acc[kVec - 1]);
        } else {
          acc[0] = fmaf(kv, __ldg(src), acc[0]);
        }
      }
    }
    float* dst = a.out + ((m * a.out_h + oy) * a.out_w + ox) * a.minor + c;
    if (kVec == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[kVec - 1]);
    else dst[0] = acc[0];
  }
}
}  // namespace buddy

using namespace buddy;

extern "C" int buddy_upfirdn2d(const float* in, const float* kernel, int major, int in_h, int in_w, int minor, int kh,
                               int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                               int pad_y1, float* out, void* stream) {
  if (!in || !kernel || !out || major <= 0 || in_h <= 0 || in_w <= 0 || minor <= 0 || kh <= 0 || kw <= 0 ||
      kh > kUpfirMaxTaps || kw > kUpfirMaxTaps || up_x <= 0 || up_y <= 0 || down_x <= 0 || down_y <= 0) {
    set_last_error("buddy_upfirdn2d: invalid argument (kernel up to %dx%d taps)", kUpfirMaxTaps, kUpfirMaxTaps);
    return BUDDY_ERR_INVALID;
  }
  UpfirArgs a;
  a.in = in;
  a.kernel = kernel;
  a.out = out;
  a.major = major;
  a.in_h = in_h;
  a.in_w = in_w;
  a.minor = minor;
  a.kh = kh;
  a.kw = kw;
  a.up_x = up_x;
  a.up_y = up_y;
  a.down_x = down_x;
  a.down_y = down_y;
  a.pad_x0 = pad_x0;
  a.pad_y0 = pad_y0;
  const int ph = in_h * up_y + pad_y0 + pad_y1 - kh, pw = in_w * up_x + pad_x0 + pad_x1 - kw;
  if (ph < 0 || pw < 0) {
    set_last_error("buddy_upfirdn2d: kernel larger than the padded input");
    return BUDDY_ERR_INVALID;
  }
  a.out_h = ph / down_y + 1;
  a.out_w = pw / down_x + 1;
  if (static_cast<long long>(a.out_h) * a.out_w * minor >= (1LL << 31)) {
    set_last_error("buddy_upfirdn2d: one image of the output must have fewer than 2^31 elements");
    return BUDDY_ERR_UNSUPPORTED;
  }
  const bool vec = minor % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const long long total = static_cast<long long>(major) * a.out_h * a.out_w * (vec ? minor / 4 : minor);
  long long grid = (total + 255) / 256;
  if (grid > 148LL * 32) grid = 148LL * 32;
  if (vec) upfirdn2d_kernel<4><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else upfirdn2d_kernel<1><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BUDDY_CHECK_LAUNCH("upfirdn2d_kernel");
  return 0;
}
