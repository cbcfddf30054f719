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

__global__ void __launch_bounds__(256) upfirdn2d_kernel(const UpfirArgs a) {
  __shared__ float kflip[kUpfirMaxTaps * kUpfirMaxTaps];
  for (int i = threadIdx.x; i < a.kh * a.kw; i += blockDim.x) {
    const int ky = i / a.kw, kx = i - ky * a.kw;
    kflip[i] = a.kernel[(a.kh - 1 - ky) * a.kw + (a.kw - 1 - kx)];
  }
  __syncthreads();
  const long long total = static_cast<long long>(a.major) * a.out_h * a.out_w * a.minor;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % a.minor);
    long long r = idx / a.minor;
    const int ox = static_cast<int>(r % a.out_w);
    r /= a.out_w;
    const int oy = static_cast<int>(r % a.out_h);
    const long long m = r / a.out_h;
    // window origin in zero-inserted input coordinates (may be negative: padding)
    const int y0 = oy * a.down_y - a.pad_y0, x0 = ox * a.down_x - a.pad_x0;
    // first tap row / column that lands on a real sample: y0 + i >= 0 and (y0 + i) % up_y == 0
    const int i0 = y0 < 0 ? -y0 : (a.up_y - (y0 % a.up_y)) % a.up_y;
    const int j0 = x0 < 0 ? -x0 : (a.up_x - (x0 % a.up_x)) % a.up_x;
    const float* base = a.in + m * a.in_h * a.in_w * a.minor + c;
    float acc = 0.f;
    for (int i = i0; i < a.kh; i += a.up_y) {
      const int iy = (y0 + i) / a.up_y;
      if (iy >= a.in_h) break;
      for (int j = j0; j < a.kw; j += a.up_x) {
        const int ix = (x0 + j) / a.up_x;
        if (ix >= a.in_w) break;
        acc = fmaf(kflip[i * a.kw + j], __ldg(base + (static_cast<long long>(iy) * a.in_w + ix) * a.minor), acc);
      }
    }
    a.out[idx] = acc;
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
  const long long total = static_cast<long long>(major) * a.out_h * a.out_w * minor;
  long long grid = (total + 255) / 256;
  if (grid > 148LL * 32) grid = 148LL * 32;
  upfirdn2d_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BUDDY_CHECK_LAUNCH("upfirdn2d_kernel");
  return 0;
}
