// Dev probe (not product): can a K-major SWIZZLE_128B UMMA descriptor start at any 128-byte line of a TMA-written
// patch (i.e. is the XOR swizzle a function of the absolute shared-memory address), and does that still hold when the
// 8-row groups are 1280 bytes apart (SBO = 10 pixels: a (16+2) x (8+2)-pixel halo patch in the CURRENT channels-last
// layout)?  If yes, ONE patch per 64-channel chunk serves all nine taps of a 3x3 conv with no producer change.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_sw128_shift_probe umma_sw128_shift_probe.cu
// smem line L (128 B) holds 8 chunks of 8 fp16; chunk c is stored at position c ^ (L & 7) (what TMA SWIZZLE_128B writes
// into a 1024-byte aligned destination).  A[L][k] = (k / 8) * 256 + L.  B = 16x16 identity, so
// D[m][n] must be A[line(m)][16 * ksub + n] with line(m) = (m / 8) * (SBO / 128) + m % 8 + shift.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../buddy_b200/csrc/common.cuh"
namespace buddy { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace buddy;

__device__ __forceinline__ uint64_t make_noswizzle_kmajor_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

constexpr int kLines = 256;
__global__ void __launch_bounds__(128, 1) probe(int shift, int sbo, int ksub, int use_base_off, float* out) {
  __shared__ __align__(1024) __half sA[kLines * 64];
  __shared__ __align__(1024) __half sB[2 * 16 * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kLines * 64; i += blockDim.x) {
    const int L = i / 64, pos = (i / 8) % 8, e = i % 8;
    const int c = pos ^ (L & 7);   // logical chunk stored at this position
    (void)e;
    sA[i] = __float2half(float(c * 256 + L));
  }
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    const int kg = i / (16 * 8), n = (i / 8) % 16, kk = i % 8;
    sB[i] = __float2half((kg * 8 + kk) == n ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t d = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, 16);
    const uint32_t start = smem_u32(sA) + shift * 128;
    const uint64_t da = make_sw128_desc(start, sbo, use_base_off ? ((start >> 7) & 7) : 0) + 2 * ksub;
    const uint64_t db = make_noswizzle_kmajor_desc(smem_u32(sB), 16 * 16, 128);
    umma_f16(d, da, db, idesc, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(d + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(d, 32); }
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 16 * 4);
  float h[128 * 16];
  for (int bo = 0; bo < 2; ++bo)
    for (int sbo : {1024, 1280}) {
      int bad_cfg = 0;
      for (int shift : {0, 1, 2, 3, 7, 8, 10, 11, 12, 20, 21, 22}) {
        int bad = 0;
        int err = 0;
        for (int ksub = 0; ksub < 4; ++ksub) {
          cudaMemset(out, 0, sizeof(h));
          probe<<<1, 128>>>(shift, sbo, ksub, bo, out);
          err |= (int)cudaDeviceSynchronize();
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
              const int L = (m / 8) * (sbo / 128) + (m % 8) + shift;
              const float want = float((2 * ksub + n / 8) * 256 + L);
              if (h[m * 16 + n] != want) ++bad;
            }
        }
        printf("base_off %s SBO %4d shift %2d lines: %s (%d mismatches) err=%d\n", bo ? "set" : "0  ", sbo, shift,
               bad ? "MISMATCH" : "exact", bad, err);
        bad_cfg += bad;
      }
      printf("== base_off %s SBO %d: %s\n", bo ? "set" : "0", sbo, bad_cfg ? "NOT linear" : "ALL EXACT");
    }
  return 0;
}
